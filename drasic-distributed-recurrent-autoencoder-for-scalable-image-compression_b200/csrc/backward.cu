// HBM-bound kernels of backward-through-time for the DRASIC loop (what autograd executes for
// train_model.py:115 over models/dist_codec.py:36-65): LSTM cell pointwise backward, fp32 -> scaled
// fp16 operand conversion for the tensor-core dgrad / wgrad GEMMs, weight (un)packing for those GEMMs
// and the backward of the fused bottleneck and tail/head kernels (pointwise.cu).
#include "backward.h"
#include "pixel_stream.cuh"
#include "pointwise.h"

namespace drasic {

namespace {

__device__ __forceinline__ int logical_of_b(int p, int C, int order) {
  if (order == ORDER_QUAD) {
    const int cq = C >> 2;
    const int q = p / cq, c = p - q * cq;
    return c * 4 + q;
  }
  return p;
}
__device__ __forceinline__ int find_group_b(const int* ends, int n_groups, int idx) {
  int g = 0;
  while (g + 1 < n_groups && idx >= ends[g]) ++g;
  return g;
}

// ------------------------------------------------------------------ LSTM cell backward -> scaled fp16 planes
// forward (cell.py:133-139): i,f,o = sigmoid, g = tanh; c = f*c_prev + i*g; h = o*tanh(c)
__device__ __forceinline__ int shift_for(float amax, int target_exp) {
  if (!(amax > 0.0f) || !isfinite(amax)) return 0;
  int e;
  frexpf(amax, &e);   // amax * 2^(target_exp - e) in [2^(target_exp-1), 2^target_exp)
  return target_exp - e;
}

__global__ void __launch_bounds__(256, 3) lstm_bwd_fused_kernel(const LstmBwdFusedParams p) {
  ptx::griddep_launch_dependents();
  ptx::griddep_wait();
  // predicted scale: the previous iteration's amax lands in [2^11, 2^12) -- 16x head-room for growth
  const int shift_pred = shift_for(__uint_as_float(*p.prev_amax_bits), 12);
  int shift = shift_pred;
  if (p.verify) {
    const float a = __uint_as_float(*p.amax_bits);
    if (!(a > 0.0f)) return;                                 // all-zero tensor: any scale is exact
    const float scaled = ldexpf(a, shift_pred);
    if (isfinite(a) && scaled >= 16.0f && scaled <= 60000.0f) return;   // the prediction was good enough
    shift = shift_for(a, 14);                                // exact: [2^13, 2^14)
  }
  const float scale = ldexpf(1.0f, shift);
  if (blockIdx.x == 0 && threadIdx.x == 0) *p.inv_scale_out = ldexpf(1.0f, -shift);
  const int Co = p.Co;
  const size_t n_vec = p.n_pix * Co / 8;
  float amax = 0.0f;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t v = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; v < n_vec; v += stride) {
    const size_t e0 = v * 8;
    const size_t pix = e0 / Co;
    const int ch = static_cast<int>(e0 - pix * Co);
    auto ld8 = [](const float* src, float (&dst)[8]) {
      const float4 a = reinterpret_cast<const float4*>(src)[0], b = reinterpret_cast<const float4*>(src)[1];
      dst[0] = a.x; dst[1] = a.y; dst[2] = a.z; dst[3] = a.w; dst[4] = b.x; dst[5] = b.y; dst[6] = b.z; dst[7] = b.w;
    };
    float gi[8], gf[8], gg[8], go[8];
    if (p.gates_f32) {
      const float* gp = reinterpret_cast<const float*>(p.gates) + pix * (4 * Co) + ch;
      ld8(gp, gi); ld8(gp + Co, gf); ld8(gp + 2 * Co, gg); ld8(gp + 3 * Co, go);
    } else {
      const __half* gp = reinterpret_cast<const __half*>(p.gates) + pix * (4 * Co) + ch;
      auto ldh = [](const __half* src, float (&dst)[8]) {
        alignas(16) __half t[8];
        *reinterpret_cast<uint4*>(t) = *reinterpret_cast<const uint4*>(src);
        for (int i = 0; i < 8; ++i) dst[i] = __half2float(t[i]);
      };
      ldh(gp, gi); ldh(gp + Co, gf); ldh(gp + 2 * Co, gg); ldh(gp + 3 * Co, go);
    }
    float ct[8], cp[8], dh[8], dcv[8];
    ld8(p.c_t + e0, ct);
    if (p.c_prev) ld8(p.c_prev + e0, cp);
    else for (int i = 0; i < 8; ++i) cp[i] = 0.0f;
    ld8(p.dh_above + e0, dh);
    if (p.dh_rec) {
      float r[8];
      ld8(p.dh_rec + e0, r);
      for (int i = 0; i < 8; ++i) dh[i] += r[i];
    }
    if (p.dc_in) ld8(p.dc_in + e0, dcv);
    else for (int i = 0; i < 8; ++i) dcv[i] = 0.0f;
    float d4[4][8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float ig = gi[i], fg = gf[i], cg = gg[i], og = go[i];
      const float tc = tanhf(ct[i]);
      const float dct = fmaf(dh[i] * og, 1.0f - tc * tc, dcv[i]);
      d4[3][i] = dh[i] * tc * og * (1.0f - og);
      d4[0][i] = dct * cg * ig * (1.0f - ig);
      d4[1][i] = dct * cp[i] * fg * (1.0f - fg);
      d4[2][i] = dct * ig * (1.0f - cg * cg);
      dcv[i] = dct * fg;
      amax = fmaxf(amax, fmaxf(fmaxf(fabsf(d4[0][i]), fabsf(d4[1][i])), fmaxf(fabsf(d4[2][i]), fabsf(d4[3][i]))));
    }
    reinterpret_cast<float4*>(p.dc_out + e0)[0] = make_float4(dcv[0], dcv[1], dcv[2], dcv[3]);
    reinterpret_cast<float4*>(p.dc_out + e0)[1] = make_float4(dcv[4], dcv[5], dcv[6], dcv[7]);
    const size_t goff = pix * (4 * Co) + ch;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      alignas(16) __half h[8], l[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float s = d4[g][i] * scale;
        h[i] = __float2half_rn(s);
        l[i] = __float2half_rn(s - __half2float(h[i]));
      }
      *reinterpret_cast<uint4*>(p.hi + goff + g * Co) = *reinterpret_cast<const uint4*>(h);
      if (p.lo) *reinterpret_cast<uint4*>(p.lo + goff + g * Co) = *reinterpret_cast<const uint4*>(l);
    }
  }
  if (!p.verify) {
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    if ((threadIdx.x & 31) == 0 && amax > 0.0f) atomicMax(p.amax_bits, __float_as_uint(amax));
  }
}

// ------------------------------------------------------------------ dgrad weight packing
// rows n in [0, Cin+Co): physical input channel of x (n < Cin) or of h; K = tap' * 4Co + gate*Co + p_out
// with the tap flipped (conv transpose): value = W[gate*Co + co, channel(n), 8 - tap'].
__global__ void amax2_kernel(const __grid_constant__ PackBatch b, size_t na, size_t nb, unsigned int* __restrict__ out) {
  const float* __restrict__ wa = b.w_in[blockIdx.y];
  const float* __restrict__ wb = b.w_h[blockIdx.y];
  float m = 0.0f;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < na + nb; i += stride)
    m = fmaxf(m, fabsf(i < na ? wa[i] : wb[i - na]));
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out + b.slot[blockIdx.y], __float_as_uint(m));
}

// Block = 64 consecutive packed gate channels r (one gate, 64 physical output channels = one 64-element run of K per
// tap) x 16 consecutive output rows n (input channels): the 64 x 16 x 9 weights are staged in shared memory from runs of
// contiguous taps, then every (n, tap) writes its 64 K elements as one 128-byte run.
constexpr int PD_ROWS = 64, PD_CH = 16;
__global__ void __launch_bounds__(256) pack_dgrad_kernel(const __grid_constant__ PackBatch b, int Cin, int Co, int in_order,
                                                         int out_order, const unsigned int* __restrict__ amax_all,
                                                         __half* __restrict__ hi_all, __half* __restrict__ lo_all,
                                                         float* __restrict__ inv_all) {
  __shared__ float s_w[PD_ROWS][PD_CH * 9 + 1];   // (odd row stride: the 64 rows of one (n, tap) sit in distinct banks)
  const float* __restrict__ w_in = b.w_in[blockIdx.z];
  const float* __restrict__ w_h = b.w_h[blockIdx.z];
  const int slot = b.slot[blockIdx.z];
  const int K = 9 * 4 * Co;
  const int N = Cin + Co;
  // tile-contiguous layout [N tile of 256 rows][K block of 64][256 rows][64]; a group's plane holds whole tiles
  const size_t plane = static_cast<size_t>((N + 255) / 256 * 256) * K;
  __half* __restrict__ out_hi = hi_all + static_cast<size_t>(slot) * plane;
  __half* __restrict__ out_lo = lo_all + static_cast<size_t>(slot) * plane;
  const float amax = __uint_as_float(amax_all[slot]);
  int shift = 0;
  if (amax > 0.0f && isfinite(amax)) {
    int e;
    frexpf(amax, &e);
    shift = 11 - e;
  }
  const float scale = ldexpf(1.0f, shift);
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) inv_all[slot] = ldexpf(1.0f, -shift);
  const int r0 = blockIdx.x * PD_ROWS;     // packed gate channels r0 .. r0+63 (Co is a multiple of 64: one gate)
  const int n0 = blockIdx.y * PD_CH;       // output rows n0 .. n0+15 (Cin is a multiple of 64: all x or all h)
  const int gate = r0 / Co, pout0 = r0 - gate * Co;
  const bool is_x = n0 < Cin;
  for (int i = threadIdx.x; i < PD_ROWS * PD_CH * 9; i += blockDim.x) {
    const int row = i / (PD_CH * 9), rem = i - row * (PD_CH * 9);
    const int ch = rem / 9, tap = rem - ch * 9;
    const int orow = gate * Co + logical_of_b(pout0 + row, Co, out_order);
    float v;
    if (is_x) v = w_in[(static_cast<size_t>(orow) * Cin + logical_of_b(n0 + ch, Cin, in_order)) * 9 + tap];
    else v = w_h[(static_cast<size_t>(orow) * Co + logical_of_b(n0 + ch - Cin, Co, out_order)) * 9 + tap];
    s_w[row][ch * 9 + tap] = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < PD_CH * 9 * PD_ROWS; i += blockDim.x) {
    const int row = i & (PD_ROWS - 1);     // fastest: the 64 consecutive K elements of one (n, tap)
    const int ct = i >> 6;
    const int ch = ct / 9, tapf = ct - ch * 9;
    const int n = n0 + ch;
    const int k = tapf * 4 * Co + r0 + row;
    const float v = s_w[row][ch * 9 + 8 - tapf] * scale;   // conv transpose: tap flipped
    const __half hi = __float2half_rn(v);
    const size_t off = ((static_cast<size_t>(n >> 8) * (K / 64) + k / 64) * 256 + (n & 255)) * 64 + (k & 63);
    out_hi[off] = hi;
    out_lo[off] = __float2half_rn(v - __half2float(hi));
  }
}

// packed-order weight gradient [4Co (gate, p_out)][9Cin (tap, p_in) | 9Co (tap, p_h)] -> OIHW fp32
__global__ void unpack_wgrad_kernel(const float* __restrict__ gw, int Cin, int Co, int in_order,
                                    int out_order, float* __restrict__ dw_in, float* __restrict__ dw_h) {
  const int K = 9 * (Cin + Co);
  const size_t total = static_cast<size_t>(4 * Co) * K;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total; idx += stride) {
    const int r = static_cast<int>(idx / K);
    const int k = static_cast<int>(idx - static_cast<size_t>(r) * K);
    const int gate = r / Co, pout = r - gate * Co;
    const int orow = gate * Co + logical_of_b(pout, Co, out_order);
    const float v = gw[idx];
    if (k < 9 * Cin) {
      const int tap = k / Cin, pin = k - tap * Cin;
      dw_in[(static_cast<size_t>(orow) * Cin + logical_of_b(pin, Cin, in_order)) * 9 + tap] = v;
    } else {
      const int kk = k - 9 * Cin;
      const int tap = kk / Co, ph = kk - tap * Co;
      dw_h[(static_cast<size_t>(orow) * Co + logical_of_b(ph, Co, out_order)) * 9 + tap] = v;
    }
  }
}

// ------------------------------------------------------------------ bottleneck backward
// forward (pointwise.cu bottleneck_kernel): z = tanh(W4 h3); q = Q(z) (straight-through); d0 = tanh(W0 q)
// Weight-gradient accumulation across blocks.  Default: float atomics (the sum depends on the order in which the
// blocks arrive -- run-to-run differences at the 1e-6 level).  Deterministic mode (fix_scale > 0): the destination is
// an int64 buffer and every partial sum is added as round(v * fix_scale); integer addition is associative, so the
// result does not depend on the order (the host converts back, engine/codec.py).
__device__ __forceinline__ unsigned long long to_fix(float v, double scale) {
  return static_cast<unsigned long long>(__double2ll_rn(static_cast<double>(v) * scale));
}
__device__ __forceinline__ void accum_grad(float* dst, size_t i, float v, double fix_scale) {
  if (fix_scale > 0.0) atomicAdd(reinterpret_cast<unsigned long long*>(dst) + i, to_fix(v, fix_scale));
  else atomicAdd(dst + i, v);
}

__global__ void __launch_bounds__(128) bottleneck_bwd_kernel(const BottleneckBwdParams p) {
  ptx::griddep_launch_dependents();
  ptx::griddep_wait();
  extern __shared__ float sm[];
  const int HW = p.HW, CH = p.Ch, code = p.code;
  float* s_a = sm;                    // [HW][CH]   dd0 -> dpre0
  float* s_h = s_a + HW * CH;         // [HW][CH]   h3
  float* s_q = s_h + HW * CH;         // [HW][code] symbols
  float* s_dz = s_q + HW * code;      // [HW][code] dprez
  const int slot = blockIdx.x;
  const int orig = p.perm[slot];
  float* dh = p.dh_e3 + static_cast<size_t>(slot) * HW * CH;
  if (orig < 0) {  // padding slot: its gradients must be exactly zero
    for (int i = threadIdx.x; i < HW * CH; i += blockDim.x) dh[i] = 0.0f;
    return;
  }
  const int ge = p.n_enc_groups > 1 ? find_group_b(p.group_img_end, p.n_enc_groups, slot) : 0;
  const int gd = p.n_dec_groups > 1 ? find_group_b(p.group_img_end, p.n_dec_groups, slot) : 0;
  const float* w4 = p.w_e4 + static_cast<size_t>(ge) * code * CH;
  const float* w0 = p.w_d0 + static_cast<size_t>(gd) * CH * code;
  const int n_code = code * HW;
  for (int i = threadIdx.x; i < HW * CH; i += blockDim.x) {
    s_a[i] = p.dd0[static_cast<size_t>(slot) * HW * CH + i];
    s_h[i] = p.h_e3[static_cast<size_t>(slot) * HW * CH + i];
  }
  for (int o = threadIdx.x; o < n_code; o += blockDim.x) {
    const int k = o / HW, pix = o - k * HW;
    s_q[pix * code + k] = p.codes[static_cast<size_t>(orig) * n_code + o];
  }
  __syncthreads();
  // dpre0 = dd0 * (1 - d0^2), d0 recomputed from the symbols
  for (int o = threadIdx.x; o < HW * CH; o += blockDim.x) {
    const int pix = o / CH, m = o - pix * CH;
    float acc = 0.0f;
    for (int k = 0; k < code; ++k) acc = fmaf(__ldg(w0 + m * code + k), s_q[pix * code + k], acc);
    const float d0 = tanhf(acc);
    s_a[o] *= (1.0f - d0 * d0);
  }
  __syncthreads();
  // dW_d0[m,k] += sum_pix dpre0[pix,m] * q[pix,k]
  for (int o = threadIdx.x; o < CH * code; o += blockDim.x) {
    const int m = o / code, k = o - m * code;
    float acc = 0.0f;
    for (int pix = 0; pix < HW; ++pix) acc = fmaf(s_a[pix * CH + m], s_q[pix * code + k], acc);
    accum_grad(p.dw_d0, static_cast<size_t>(gd) * CH * code + o, acc, p.fix_scale);
  }
  // dq = W0^T dpre0 ; straight-through quantizer ; dprez = dq * (1 - z^2)
  for (int o = threadIdx.x; o < n_code; o += blockDim.x) {
    const int k = o / HW, pix = o - k * HW;
    float acc = 0.0f;
    for (int m = 0; m < CH; ++m) acc = fmaf(__ldg(w0 + m * code + k), s_a[pix * CH + m], acc);
    const float z = p.z[static_cast<size_t>(orig) * n_code + o];
    s_dz[pix * code + k] = acc * (1.0f - z * z);
  }
  __syncthreads();
  // dW_e4[k,c] += sum_pix dprez[pix,k] * h3[pix,c]
  for (int o = threadIdx.x; o < code * CH; o += blockDim.x) {
    const int k = o / CH, c = o - k * CH;
    float acc = 0.0f;
    for (int pix = 0; pix < HW; ++pix) acc = fmaf(s_dz[pix * code + k], s_h[pix * CH + c], acc);
    accum_grad(p.dw_e4, static_cast<size_t>(ge) * code * CH + o, acc, p.fix_scale);
  }
  // dh3[pix,c] = sum_k W4[k,c] * dprez[pix,k]
  for (int o = threadIdx.x; o < HW * CH; o += blockDim.x) {
    const int pix = o / CH, c = o - pix * CH;
    float acc = 0.0f;
    for (int k = 0; k < code; ++k) acc = fmaf(__ldg(w4 + k * CH + c), s_dz[pix * code + k], acc);
    dh[o] = acc;
  }
}

// ------------------------------------------------------------------ tail / head backward
// Same thread mapping as the forward tail/head kernel (pointwise.cu): four lanes per pixel, eight of the
// 32 feature channels per lane, pixels enumerated quad-major so the feature-side tensors (D3's h, dh_d3,
// du1) are contiguous per warp.  Weight gradients accumulate in registers over all passes of the block
// and are reduced once per block (shuffles over the lanes that own the same channel slice -> shared
// atomics -> one global atomic per weight).
constexpr int TB_THREADS = 256;
constexpr int TB_PIX = TB_THREADS / 4;
constexpr int MAXC = 4;
constexpr int FEATB = 32;

__device__ __forceinline__ int slice_channel(int sub, int j) { return j < 4 ? 4 * sub + j : 16 + 4 * sub + (j - 4); }

struct QuadMapB {  // quad-major pixel index -> raster pixel offset
  int W, qw, shift;
  __device__ __forceinline__ explicit QuadMapB(int W_) : W(W_), qw(W_ >> 1) {
    shift = (qw & (qw - 1)) == 0 ? 31 - __clz(qw) : -1;
  }
  __device__ __forceinline__ int raster(int i) const {
    const int quad = i >> 2, qd = i & 3;
    const int qy = shift >= 0 ? quad >> shift : quad / qw;
    const int qx = quad - qy * qw;
    return (2 * qy + (qd >> 1)) * W + 2 * qx + (qd & 1);
  }
};

// iteration t: G = (has_acc ? G : 0) + dLoss_t/dRecon_t ; d = tanh(W4 v4) ; backward through it
template <int NC>
__global__ void __launch_bounds__(TB_THREADS, 3) tail_bwd_kernel(const TailBwdParams p) {
  ptx::griddep_launch_dependents();
  ptx::griddep_wait();
  __shared__ float s_acc[MAXC * FEATB];
  __shared__ unsigned long long s_fix[MAXC * FEATB];   // deterministic mode: the block's partial sums in fixed point
  __shared__ alignas(16) float s_w4[MAXC][4][8];   // [c][sub][j]
  const bool fix = p.fix_scale > 0.0;
  const int HW = p.H * p.W;
  const int bpi = gridDim.x / p.B_pad;
  const int slot = blockIdx.x / bpi;
  const int part = blockIdx.x - slot * bpi;
  const int ppb = HW / bpi;
  const int orig = p.perm[slot];
  float* dh_img = p.dh_d3 + static_cast<size_t>(slot) * HW * FEATB;
  if (orig < 0) {  // padding slot: zero gradient
    float4* z = reinterpret_cast<float4*>(dh_img + static_cast<size_t>(part) * ppb * FEATB);
    for (int v = threadIdx.x; v < ppb * FEATB / 4; v += TB_THREADS) z[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  constexpr int C = NC;
  const int gd = p.n_dec_groups > 1 ? find_group_b(p.group_img_end, p.n_dec_groups, slot) : 0;
  const int sub = threadIdx.x & 3;
  const int lane = threadIdx.x & 31;
  const int lane_base = lane & ~3;
  const bool mine = sub < C;
  const QuadMapB qm(p.W);
  // heterogeneous rates: a frozen node's decoded refinement is the constant 0 -> no gradient into the decoder
  const bool frozen = p.group_iters != nullptr &&
                      p.iter >= p.group_iters[p.n_enc_groups > 1 ? find_group_b(p.group_img_end, p.n_enc_groups, slot) : 0];
  for (int t = threadIdx.x; t < MAXC * FEATB; t += TB_THREADS) {
    const int c = t >> 5, sb = (t >> 3) & 3, j = t & 7;
    s_w4[c][sb][j] = c < C ? p.w_d4[(static_cast<size_t>(gd) * C + c) * FEATB + slice_channel(sb, j)] : 0.0f;
    s_acc[t] = 0.0f;
    s_fix[t] = 0ull;
  }
  __syncthreads();
  float wacc[NC][8];
#pragma unroll
  for (int c = 0; c < NC; ++c)
#pragma unroll
    for (int j = 0; j < 8; ++j) wacc[c][j] = 0.0f;
  const float gain = p.first ? 0.5f : 1.0f;
  const size_t img_base = (static_cast<size_t>(orig) * C + (mine ? sub : 0)) * HW;
  __shared__ alignas(128) float4 s_feat[TB_THREADS / 32][4][64];
  __shared__ uint64_t s_bar[TB_THREADS / 32][4];
  const int warp = threadIdx.x >> 5;
  const int n_pass = ppb / TB_PIX;
  const int i_warp = part * ppb + warp * (ppb / (TB_THREADS / 32));
  WarpStream<1, 4> ws;
  ws.buf = &s_feat[warp][0][0];
  ws.bar = &s_bar[warp][0];
  ws.src = reinterpret_cast<const char*>(p.d3_next) + (static_cast<size_t>(slot) * HW + i_warp) * FEATB * sizeof(float);
  ws.n = n_pass;
  ws.start(lane);

  int i = i_warp + (lane >> 2);
  int rpix = qm.raster(i);
  float r = 0.0f, im = 0.0f, ga = 0.0f;
  if (mine) {
    r = p.recon_t[img_base + rpix];
    im = p.img[img_base + rpix];
    if (p.has_acc) ga = p.gacc[img_base + rpix];
  }
  for (int pass = 0; pass < n_pass; ++pass, i += 8) {
    float4 fa, fb;
    ws.wait(pass);
    ws.read(pass, 0, lane, fa, fb);
    ws.release(pass, lane);
    const float f[8] = {fa.x, fa.y, fa.z, fa.w, fb.x, fb.y, fb.z, fb.w};
    const float r_c = r, im_c = im, ga_c = ga;
    const size_t gi = img_base + rpix;
    if (pass + 1 < n_pass) {  // next pass's small loads (uniform branch)
      rpix = qm.raster(i + 8);
      if (mine) {
        r = p.recon_t[img_base + rpix];
        im = p.img[img_base + rpix];
        if (p.has_acc) ga = p.gacc[img_base + rpix];
      }
    }
    float acc[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      float a = 0.0f;
#pragma unroll
      for (int j = 0; j < 8; ++j) a = fmaf(s_w4[c][sub][j], f[j], a);
      a += __shfl_xor_sync(0xffffffffu, a, 1);
      a += __shfl_xor_sync(0xffffffffu, a, 2);
      acc[c] = a;
    }
    float dpre_mine = 0.0f;
    if (mine) {
      float direct;
      if (p.loss_kind == LOSS_MAE) direct = (r_c > im_c) ? 1.0f : ((r_c < im_c) ? -1.0f : 0.0f);
      else if (p.loss_kind == LOSS_MSE) direct = 2.0f * (r_c - im_c);
      // torch's binary_cross_entropy backward: (input - target) / max((1 - input) * input, 1e-12), also where the
      // forward's log clamp at -100 is active (the forward refuses inputs outside [0, 1], see tail_head_kernel)
      else direct = (r_c - im_c) / fmaxf(r_c * (1.0f - r_c), 1e-12f);
      const float G = ga_c + direct * p.loss_scale;
      p.gacc[gi] = G;
      float pre = acc[0];
#pragma unroll
      for (int c = 1; c < NC; ++c) pre = sub == c ? acc[c] : pre;
      const float d = tanhf(pre);
      dpre_mine = frozen ? 0.0f : G * gain * (1.0f - d * d);
    }
    float dpre[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) dpre[c] = __shfl_sync(0xffffffffu, dpre_mine, lane_base + c);
    float dv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float a = 0.0f;
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        a = fmaf(s_w4[c][sub][j], dpre[c], a);
        wacc[c][j] = fmaf(dpre[c], f[j], wacc[c][j]);
      }
      dv[j] = a;
    }
    float4* dst = reinterpret_cast<float4*>(dh_img + static_cast<size_t>(i) * FEATB);
    dst[sub] = make_float4(dv[0], dv[1], dv[2], dv[3]);
    dst[sub + 4] = make_float4(dv[4], dv[5], dv[6], dv[7]);
  }
  // dW_d4[c,k] += sum_pixels dpre[c] * v4[k]
#pragma unroll
  for (int c = 0; c < NC; ++c)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v = wacc[c][j];
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if (lane < 4) {
        if (fix) atomicAdd(&s_fix[c * FEATB + slice_channel(sub, j)], to_fix(v, p.fix_scale));
        else atomicAdd(&s_acc[c * FEATB + slice_channel(sub, j)], v);
      }
    }
  __syncthreads();
  if (threadIdx.x < C * FEATB) {
    const size_t o = static_cast<size_t>(gd) * C * FEATB + threadIdx.x;
    if (fix) atomicAdd(reinterpret_cast<unsigned long long*>(p.dw_d4) + o, s_fix[threadIdx.x]);
    else atomicAdd(p.dw_d4 + o, s_acc[threadIdx.x]);
  }
}

// iteration t: e0 = tanh(W0 x + b) recomputed; backward to W0, b and (dist codec, t > 0) to recon_{t-1}
template <int NC>
__global__ void __launch_bounds__(TB_THREADS, 3) head_bwd_kernel(const HeadBwdParams p) {
  ptx::griddep_launch_dependents();
  ptx::griddep_wait();
  __shared__ float s_acc[FEATB * MAXC + FEATB];
  __shared__ unsigned long long s_fix[FEATB * MAXC + FEATB];   // deterministic mode
  __shared__ alignas(16) float s_w0[4][8][MAXC];   // [sub][j][c]
  const bool fix = p.fix_scale > 0.0;
  __shared__ alignas(16) float s_b0[4][8];
  const int HW = p.H * p.W;
  const int bpi = gridDim.x / p.B_pad;
  const int slot = blockIdx.x / bpi;
  const int part = blockIdx.x - slot * bpi;
  const int ppb = HW / bpi;
  const int orig = p.perm[slot];
  if (orig < 0) return;
  constexpr int C = NC;
  const int ge = p.n_enc_groups > 1 ? find_group_b(p.group_img_end, p.n_enc_groups, slot) : 0;
  const int sub = threadIdx.x & 3;
  const int lane = threadIdx.x & 31;
  const int lane_base = lane & ~3;
  const bool mine = sub < C;
  const QuadMapB qm(p.W);
  for (int t = threadIdx.x; t < FEATB * MAXC + FEATB; t += TB_THREADS) { s_acc[t] = 0.0f; s_fix[t] = 0ull; }
  for (int t = threadIdx.x; t < MAXC * FEATB; t += TB_THREADS) {
    const int c = t >> 5, sb = (t >> 3) & 3, j = t & 7;
    const int k = slice_channel(sb, j);
    s_w0[sb][j][c] = c < C ? p.w_e0[(static_cast<size_t>(ge) * FEATB + k) * C + c] : 0.0f;
    if (c == 0) s_b0[sb][j] = p.b_e0 ? p.b_e0[static_cast<size_t>(ge) * FEATB + k] : 0.0f;
  }
  __syncthreads();
  float wacc[8][NC], bacc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    bacc[j] = 0.0f;
#pragma unroll
    for (int c = 0; c < NC; ++c) wacc[j][c] = 0.0f;
  }
  const size_t img_base = (static_cast<size_t>(orig) * C + (mine ? sub : 0)) * HW;
  __shared__ alignas(128) float4 s_feat[TB_THREADS / 32][4][64];
  __shared__ uint64_t s_bar[TB_THREADS / 32][4];
  const int warp = threadIdx.x >> 5;
  const int n_pass = ppb / TB_PIX;
  const int i_warp = part * ppb + warp * (ppb / (TB_THREADS / 32));
  WarpStream<1, 4> ws;
  ws.buf = &s_feat[warp][0][0];
  ws.bar = &s_bar[warp][0];
  ws.src = reinterpret_cast<const char*>(p.du1) + (static_cast<size_t>(slot) * HW + i_warp) * FEATB * sizeof(float);
  ws.n = n_pass;
  ws.start(lane);

  int i = i_warp + (lane >> 2);
  int rpix = qm.raster(i);
  float im = 0.0f, rp = 0.0f, ga = 0.0f;
  if (mine) {
    im = p.img[img_base + rpix];
    if (!p.first) rp = p.recon_prev[img_base + rpix];
    if (p.propagate) ga = p.gacc[img_base + rpix];
  }
  for (int pass = 0; pass < n_pass; ++pass, i += 8) {
    float4 da, db;
    ws.wait(pass);
    ws.read(pass, 0, lane, da, db);
    ws.release(pass, lane);
    const float du[8] = {da.x, da.y, da.z, da.w, db.x, db.y, db.z, db.w};
    const float im_c = im, rp_c = rp, ga_c = ga;
    const size_t gi = img_base + rpix;
    if (pass + 1 < n_pass) {  // next pass's small loads (uniform branch)
      rpix = qm.raster(i + 8);
      if (mine) {
        im = p.img[img_base + rpix];
        if (!p.first) rp = p.recon_prev[img_base + rpix];
        if (p.propagate) ga = p.gacc[img_base + rpix];
      }
    }
    float xv = 0.0f;
    if (mine) xv = p.first ? __fsub_rn(__fmul_rn(im_c, 2.0f), 1.0f) : __fsub_rn(im_c, rp_c);
    float x[NC], dx[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      x[c] = __shfl_sync(0xffffffffu, xv, lane_base + c);
      dx[c] = 0.0f;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float w[NC];
#pragma unroll
      for (int c = 0; c < NC; ++c) w[c] = s_w0[sub][j][c];
      float a = s_b0[sub][j];
#pragma unroll
      for (int c = 0; c < NC; ++c) a = fmaf(w[c], x[c], a);
      const float e = tanhf(a);
      const float dpre = du[j] * (1.0f - e * e);
      bacc[j] += dpre;
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        dx[c] = fmaf(w[c], dpre, dx[c]);
        wacc[j][c] = fmaf(dpre, x[c], wacc[j][c]);
      }
    }
    if (p.propagate) {  // uniform
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        dx[c] += __shfl_xor_sync(0xffffffffu, dx[c], 1);
        dx[c] += __shfl_xor_sync(0xffffffffu, dx[c], 2);
      }
      if (mine) {
        float dxm = dx[0];
#pragma unroll
        for (int c = 1; c < NC; ++c) dxm = sub == c ? dx[c] : dxm;
        p.gacc[gi] = ga_c - dxm;   // x_t = img - recon_{t-1}  (dist_codec.py:59)
      }
    }
  }
  // dW_e0[k,c] += sum_pixels dpre[k]*x[c], db[k] += sum dpre[k]
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = slice_channel(sub, j);
#pragma unroll
    for (int c = 0; c <= NC; ++c) {
      float v = c == NC ? bacc[j] : wacc[j][c < NC ? c : 0];
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if (lane < 4) {
        const int si = c == NC ? FEATB * MAXC + k : k * MAXC + c;
        if (fix) atomicAdd(&s_fix[si], to_fix(v, p.fix_scale)); else atomicAdd(&s_acc[si], v);
      }
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < FEATB * C; t += blockDim.x) {
    const int k = t / C, c = t - k * C;
    const size_t o = static_cast<size_t>(ge) * FEATB * C + t;
    if (fix) atomicAdd(reinterpret_cast<unsigned long long*>(p.dw_e0) + o, s_fix[k * MAXC + c]);
    else atomicAdd(p.dw_e0 + o, s_acc[k * MAXC + c]);
  }
  if (p.db_e0 && threadIdx.x < FEATB) {
    const size_t o = static_cast<size_t>(ge) * FEATB + threadIdx.x;
    if (fix) atomicAdd(reinterpret_cast<unsigned long long*>(p.db_e0) + o, s_fix[FEATB * MAXC + threadIdx.x]);
    else atomicAdd(p.db_e0 + o, s_acc[FEATB * MAXC + threadIdx.x]);
  }
}

inline int grid_for(size_t work_items, int threads) {
  const size_t blocks = (work_items + threads - 1) / threads;
  return static_cast<int>(blocks < 148 * 16 ? (blocks ? blocks : 1) : 148 * 16);
}

}  // namespace

int launch_lstm_bwd_fused(const LstmBwdFusedParams& p, cudaStream_t stream) {
  int grid = grid_for(p.n_pix * p.Co / 8, 256);
  // the verify pass almost always returns at once: one resident wave (3 blocks per SM) keeps its launch short and
  // still runs at full bandwidth (grid-stride loop) when the tensor does have to be redone
  if (p.verify && grid > 148 * 3) grid = 148 * 3;
  launch_pdl(lstm_bwd_fused_kernel, dim3(grid), dim3(256), 0, stream, p);
  return static_cast<int>(cudaGetLastError());
}
int launch_pack_dgrad_weights(const PackBatch& b, int n, int n_slots, int Cin, int Co, int in_order, int out_order,
                              __half* out_hi, __half* out_lo, float* inv_scale_out, unsigned int* scratch_amax,
                              cudaStream_t stream) {
  cudaError_t e = cudaMemsetAsync(scratch_amax, 0, sizeof(unsigned int) * n_slots, stream);
  if (e != cudaSuccess) return static_cast<int>(e);
  const size_t na = static_cast<size_t>(4 * Co) * Cin * 9, nb = static_cast<size_t>(4 * Co) * Co * 9;
  const int bx = n >= 4 ? 148 : 296;
  amax2_kernel<<<dim3(bx, n), 256, 0, stream>>>(b, na, nb, scratch_amax);
  pack_dgrad_kernel<<<dim3(4 * Co / PD_ROWS, (Cin + Co) / PD_CH, n), 256, 0, stream>>>(b, Cin, Co, in_order, out_order,
                                                                                       scratch_amax, out_hi, out_lo, inv_scale_out);
  return static_cast<int>(cudaGetLastError());
}
int launch_unpack_wgrad(const float* gw, int Cin, int Co, int in_order, int out_order, float* dw_in,
                        float* dw_h, cudaStream_t stream) {
  unpack_wgrad_kernel<<<1184, 256, 0, stream>>>(gw, Cin, Co, in_order, out_order, dw_in, dw_h);
  return static_cast<int>(cudaGetLastError());
}
int launch_bottleneck_bwd(const BottleneckBwdParams& p, cudaStream_t stream) {
  const size_t smem = (static_cast<size_t>(2) * p.HW * p.Ch + 2 * p.HW * p.code) * sizeof(float);
  launch_pdl(bottleneck_bwd_kernel, dim3(p.B_pad), dim3(128), smem, stream, p);
  return static_cast<int>(cudaGetLastError());
}
int launch_tail_bwd(const TailBwdParams& p, cudaStream_t stream) {
  const int HW = p.H * p.W;
  if (HW % TB_THREADS != 0 || p.C > MAXC || p.C < 1) return static_cast<int>(cudaErrorInvalidValue);
  const int grid = p.B_pad * pixel_blocks_per_image(p.B_pad, HW, TB_PIX, 148 * 6);
  switch (p.C) {
    case 1: launch_pdl(tail_bwd_kernel<1>, dim3(grid), dim3(TB_THREADS), 0, stream, p); break;
    case 2: launch_pdl(tail_bwd_kernel<2>, dim3(grid), dim3(TB_THREADS), 0, stream, p); break;
    case 3: launch_pdl(tail_bwd_kernel<3>, dim3(grid), dim3(TB_THREADS), 0, stream, p); break;
    default: launch_pdl(tail_bwd_kernel<4>, dim3(grid), dim3(TB_THREADS), 0, stream, p); break;
  }
  return static_cast<int>(cudaGetLastError());
}
int launch_head_bwd(const HeadBwdParams& p, cudaStream_t stream) {
  const int HW = p.H * p.W;
  if (HW % TB_THREADS != 0 || p.C > MAXC || p.C < 1) return static_cast<int>(cudaErrorInvalidValue);
  const int grid = p.B_pad * pixel_blocks_per_image(p.B_pad, HW, TB_PIX, 148 * 6);
  switch (p.C) {
    case 1: launch_pdl(head_bwd_kernel<1>, dim3(grid), dim3(TB_THREADS), 0, stream, p); break;
    case 2: launch_pdl(head_bwd_kernel<2>, dim3(grid), dim3(TB_THREADS), 0, stream, p); break;
    case 3: launch_pdl(head_bwd_kernel<3>, dim3(grid), dim3(TB_THREADS), 0, stream, p); break;
    default: launch_pdl(head_bwd_kernel<4>, dim3(grid), dim3(TB_THREADS), 0, stream, p); break;
  }
  return static_cast<int>(cudaGetLastError());
}

}  // namespace drasic
