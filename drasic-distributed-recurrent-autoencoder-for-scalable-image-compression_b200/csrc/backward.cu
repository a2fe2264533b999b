// HBM-bound kernels of backward-through-time for the DRASIC loop (what autograd executes for
// train_model.py:115 over models/dist_codec.py:36-65): LSTM cell pointwise backward, fp32 -> scaled
// fp16 operand conversion for the tensor-core dgrad / wgrad GEMMs, weight (un)packing for those GEMMs
// and the backward of the fused bottleneck and tail/head kernels (pointwise.cu).
#include "backward.h"
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

// ------------------------------------------------------------------ LSTM pointwise backward
// forward (cell.py:133-139): i,f,o = sigmoid, g = tanh; c = f*c_prev + i*g; h = o*tanh(c)
__global__ void __launch_bounds__(256) lstm_bwd_kernel(const LstmBwdParams p) {
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
    if (p.has_dc) ld8(p.dc + e0, dcv);
    else for (int i = 0; i < 8; ++i) dcv[i] = 0.0f;
    float di[8], df[8], dg[8], dob[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float ig = gi[i], fg = gf[i], cg = gg[i], og = go[i];
      const float tc = tanhf(ct[i]);
      const float dct = fmaf(dh[i] * og, 1.0f - tc * tc, dcv[i]);
      dob[i] = dh[i] * tc * og * (1.0f - og);
      di[i] = dct * cg * ig * (1.0f - ig);
      df[i] = dct * cp[i] * fg * (1.0f - fg);
      dg[i] = dct * ig * (1.0f - cg * cg);
      dcv[i] = dct * fg;
      amax = fmaxf(amax, fmaxf(fmaxf(fabsf(di[i]), fabsf(df[i])), fmaxf(fabsf(dg[i]), fabsf(dob[i]))));
    }
    auto st8 = [](float* dst, const float (&src)[8]) {
      reinterpret_cast<float4*>(dst)[0] = make_float4(src[0], src[1], src[2], src[3]);
      reinterpret_cast<float4*>(dst)[1] = make_float4(src[4], src[5], src[6], src[7]);
    };
    float* dgp = p.dgates + pix * (4 * Co) + ch;
    st8(dgp, di);
    st8(dgp + Co, df);
    st8(dgp + 2 * Co, dg);
    st8(dgp + 3 * Co, dob);
    st8(p.dc + e0, dcv);
  }
  for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  if ((threadIdx.x & 31) == 0 && amax > 0.0f) atomicMax(p.amax_bits, __float_as_uint(amax));
}

// ------------------------------------------------------------------ fp32 -> scaled fp16 planes
__global__ void __launch_bounds__(256) convert_scale_kernel(const ConvertParams p) {
  const float amax = __uint_as_float(*p.amax_bits);
  int shift = 0;
  if (amax > 0.0f && isfinite(amax)) {
    int e;
    frexpf(amax, &e);   // amax * 2^(14-e) in [2^13, 2^14): 4x head-room below the fp16 maximum
    shift = 14 - e;
  }
  const float scale = ldexpf(1.0f, shift);
  if (blockIdx.x == 0 && threadIdx.x == 0) *p.inv_scale_out = ldexpf(1.0f, -shift);
  const size_t n_vec = p.n / 8;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t v = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; v < n_vec; v += stride) {
    const float4 a = reinterpret_cast<const float4*>(p.src)[2 * v], b = reinterpret_cast<const float4*>(p.src)[2 * v + 1];
    const float x[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    alignas(16) __half h[8], l[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float s = x[i] * scale;
      h[i] = __float2half_rn(s);
      l[i] = __float2half_rn(s - __half2float(h[i]));
    }
    reinterpret_cast<uint4*>(p.hi)[v] = *reinterpret_cast<const uint4*>(h);
    if (p.lo) reinterpret_cast<uint4*>(p.lo)[v] = *reinterpret_cast<const uint4*>(l);
  }
}

// ------------------------------------------------------------------ dgrad weight packing
// rows n in [0, Cin+Co): physical input channel of x (n < Cin) or of h; K = tap' * 4Co + gate*Co + p_out
// with the tap flipped (conv transpose): value = W[gate*Co + co, channel(n), 8 - tap'].
__global__ void amax2_kernel(const float* __restrict__ a, size_t na, const float* __restrict__ b,
                             size_t nb, unsigned int* __restrict__ out) {
  float m = 0.0f;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < na + nb; i += stride)
    m = fmaxf(m, fabsf(i < na ? a[i] : b[i - na]));
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));
}

__global__ void pack_dgrad_kernel(const float* __restrict__ w_in, const float* __restrict__ w_h, int Cin,
                                  int Co, int in_order, int out_order,
                                  const unsigned int* __restrict__ amax_bits, __half* __restrict__ out_hi,
                                  __half* __restrict__ out_lo, float* __restrict__ inv_scale_out) {
  const int K = 9 * 4 * Co;
  const int N = Cin + Co;
  const size_t total = static_cast<size_t>(N) * K;
  const float amax = __uint_as_float(*amax_bits);
  int shift = 0;
  if (amax > 0.0f && isfinite(amax)) {
    int e;
    frexpf(amax, &e);
    shift = 11 - e;
  }
  const float scale = ldexpf(1.0f, shift);
  if (blockIdx.x == 0 && threadIdx.x == 0) *inv_scale_out = ldexpf(1.0f, -shift);
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total; idx += stride) {
    const int n = static_cast<int>(idx / K);
    const int k = static_cast<int>(idx - static_cast<size_t>(n) * K);
    const int tapf = k / (4 * Co);
    const int r = k - tapf * 4 * Co;
    const int gate = r / Co, pout = r - gate * Co;
    const int orow = gate * Co + logical_of_b(pout, Co, out_order);
    const int tap = 8 - tapf;
    float v;
    if (n < Cin) v = w_in[(static_cast<size_t>(orow) * Cin + logical_of_b(n, Cin, in_order)) * 9 + tap];
    else v = w_h[(static_cast<size_t>(orow) * Co + logical_of_b(n - Cin, Co, out_order)) * 9 + tap];
    v *= scale;
    const __half hi = __float2half_rn(v);
    out_hi[idx] = hi;
    out_lo[idx] = __float2half_rn(v - __half2float(hi));
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
__global__ void __launch_bounds__(128) bottleneck_bwd_kernel(const BottleneckBwdParams p) {
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
    atomicAdd(p.dw_d0 + static_cast<size_t>(gd) * CH * code + o, acc);
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
    atomicAdd(p.dw_e4 + static_cast<size_t>(ge) * code * CH + o, acc);
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
constexpr int TB_THREADS = 256;
constexpr int MAXC = 4;
constexpr int FEATB = 32;

// iteration t: G = (has_acc ? G : 0) + dLoss_t/dRecon_t ; d = tanh(W4 v4) ; backward through it
__global__ void __launch_bounds__(TB_THREADS) tail_bwd_kernel(const TailBwdParams p) {
  __shared__ float s_w4[MAXC * FEATB];
  __shared__ float s_acc[MAXC * FEATB];
  const int HW = p.H * p.W;
  const int blocks_per_img = HW / TB_THREADS;
  const int slot = blockIdx.x / blocks_per_img;
  const int pix = (blockIdx.x - slot * blocks_per_img) * TB_THREADS + threadIdx.x;
  const int orig = p.perm[slot];
  const int Y = pix / p.W, X = pix - Y * p.W;
  const int qd = (Y & 1) * 2 + (X & 1);
  const size_t ppix = (static_cast<size_t>(slot) * (p.H >> 1) + (Y >> 1)) * (p.W >> 1) + (X >> 1);
  float4* dst = reinterpret_cast<float4*>(p.dh_d3 + ppix * (4 * FEATB) + qd * FEATB);
  if (orig < 0) {  // padding slot: zero gradient
#pragma unroll
    for (int v = 0; v < FEATB / 4; ++v) dst[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  const int C = p.C;
  const int gd = p.n_dec_groups > 1 ? find_group_b(p.group_img_end, p.n_dec_groups, slot) : 0;
  if (threadIdx.x < C * FEATB) s_w4[threadIdx.x] = p.w_d4[static_cast<size_t>(gd) * C * FEATB + threadIdx.x];
  __syncthreads();
  float f[FEATB];
  const float4* src = reinterpret_cast<const float4*>(p.d3_next + (static_cast<size_t>(slot) * HW + pix) * FEATB);
#pragma unroll
  for (int v = 0; v < FEATB / 4; ++v) {
    const float4 t = src[v];
    f[4 * v + 0] = t.x; f[4 * v + 1] = t.y; f[4 * v + 2] = t.z; f[4 * v + 3] = t.w;
  }
  float dv[FEATB];
#pragma unroll
  for (int k = 0; k < FEATB; ++k) dv[k] = 0.0f;
  float dpre[MAXC] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    if (c < C) {
      const size_t gi = (static_cast<size_t>(orig) * C + c) * HW + pix;
      const float r = p.recon_t[gi], im = p.img[gi];
      float direct;
      if (p.loss_kind == LOSS_MAE) direct = (r > im) ? 1.0f : ((r < im) ? -1.0f : 0.0f);
      else if (p.loss_kind == LOSS_MSE) direct = 2.0f * (r - im);
      else direct = (r - im) / fmaxf(r * (1.0f - r), 1e-12f);
      const float G = (p.has_acc ? p.gacc[gi] : 0.0f) + direct * p.loss_scale;
      p.gacc[gi] = G;
      float acc = 0.0f;
#pragma unroll
      for (int k = 0; k < FEATB; ++k) acc = fmaf(s_w4[c * FEATB + k], f[k], acc);
      const float d = tanhf(acc);
      dpre[c] = G * (p.first ? 0.5f : 1.0f) * (1.0f - d * d);
#pragma unroll
      for (int k = 0; k < FEATB; ++k) dv[k] = fmaf(s_w4[c * FEATB + k], dpre[c], dv[k]);
    }
  }
#pragma unroll
  for (int v = 0; v < FEATB / 4; ++v) dst[v] = make_float4(dv[4 * v], dv[4 * v + 1], dv[4 * v + 2], dv[4 * v + 3]);
  // dW_d4[c,k] += sum_pixels dpre[c] * v4[k]: warp shuffle -> shared atomics -> one global atomic per block
  for (int i = threadIdx.x; i < MAXC * FEATB; i += blockDim.x) s_acc[i] = 0.0f;
  __syncthreads();
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    if (c < C) {
#pragma unroll
      for (int k = 0; k < FEATB; ++k) {
        float v = dpre[c] * f[k];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) atomicAdd(&s_acc[c * FEATB + k], v);
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < C * FEATB)
    atomicAdd(p.dw_d4 + static_cast<size_t>(gd) * C * FEATB + threadIdx.x, s_acc[threadIdx.x]);
}

// iteration t: e0 = tanh(W0 x + b) recomputed; backward to W0, b and (dist codec, t > 0) to recon_{t-1}
__global__ void __launch_bounds__(TB_THREADS) head_bwd_kernel(const HeadBwdParams p) {
  __shared__ float s_w0[FEATB * MAXC];
  __shared__ float s_b0[FEATB];
  __shared__ float s_acc[FEATB * MAXC + FEATB];
  const int HW = p.H * p.W;
  const int blocks_per_img = HW / TB_THREADS;
  const int slot = blockIdx.x / blocks_per_img;
  const int pix = (blockIdx.x - slot * blocks_per_img) * TB_THREADS + threadIdx.x;
  const int orig = p.perm[slot];
  if (orig < 0) return;
  const int C = p.C;
  const int ge = p.n_enc_groups > 1 ? find_group_b(p.group_img_end, p.n_enc_groups, slot) : 0;
  if (threadIdx.x < C * FEATB) s_w0[threadIdx.x] = p.w_e0[static_cast<size_t>(ge) * FEATB * C + threadIdx.x];
  if (threadIdx.x < FEATB) s_b0[threadIdx.x] = p.b_e0 ? p.b_e0[static_cast<size_t>(ge) * FEATB + threadIdx.x] : 0.0f;
  __syncthreads();
  const int Y = pix / p.W, X = pix - Y * p.W;
  const int qd = (Y & 1) * 2 + (X & 1);
  const size_t ppix = (static_cast<size_t>(slot) * (p.H >> 1) + (Y >> 1)) * (p.W >> 1) + (X >> 1);
  const float4* src = reinterpret_cast<const float4*>(p.du1 + ppix * (4 * FEATB) + qd * FEATB);
  float du[FEATB];
#pragma unroll
  for (int v = 0; v < FEATB / 4; ++v) {
    const float4 t = src[v];
    du[4 * v + 0] = t.x; du[4 * v + 1] = t.y; du[4 * v + 2] = t.z; du[4 * v + 3] = t.w;
  }
  float x[MAXC];
  for (int c = 0; c < C; ++c) {
    const size_t gi = (static_cast<size_t>(orig) * C + c) * HW + pix;
    const float im = p.img[gi];
    x[c] = p.first ? __fsub_rn(__fmul_rn(im, 2.0f), 1.0f) : __fsub_rn(im, p.recon_prev[gi]);
  }
  float dpre[FEATB];
  float dx[MAXC] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int k = 0; k < FEATB; ++k) {
    float a = s_b0[k];
    for (int c = 0; c < C; ++c) a = fmaf(s_w0[k * C + c], x[c], a);
    const float e = tanhf(a);
    dpre[k] = du[k] * (1.0f - e * e);
    for (int c = 0; c < C; ++c) dx[c] = fmaf(s_w0[k * C + c], dpre[k], dx[c]);
  }
  if (p.propagate) {
    for (int c = 0; c < C; ++c) {
      const size_t gi = (static_cast<size_t>(orig) * C + c) * HW + pix;
      p.gacc[gi] -= dx[c];   // x_t = img - recon_{t-1}  (dist_codec.py:59)
    }
  }
  // dW_e0[k,c] += sum_pixels dpre[k]*x[c], db[k] += sum dpre[k]
  for (int i = threadIdx.x; i < FEATB * MAXC + FEATB; i += blockDim.x) s_acc[i] = 0.0f;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < FEATB; ++k) {
#pragma unroll
    for (int c = 0; c <= MAXC; ++c) {
      if (c < C || c == MAXC) {
        float v = c == MAXC ? dpre[k] : dpre[k] * x[c];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) atomicAdd(&s_acc[c == MAXC ? FEATB * MAXC + k : k * MAXC + c], v);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < FEATB * C; i += blockDim.x) {
    const int k = i / C, c = i - k * C;
    atomicAdd(p.dw_e0 + static_cast<size_t>(ge) * FEATB * C + i, s_acc[k * MAXC + c]);
  }
  if (p.db_e0 && threadIdx.x < FEATB) atomicAdd(p.db_e0 + static_cast<size_t>(ge) * FEATB + threadIdx.x, s_acc[FEATB * MAXC + threadIdx.x]);
}

inline int grid_for(size_t work_items, int threads) {
  const size_t blocks = (work_items + threads - 1) / threads;
  return static_cast<int>(blocks < 148 * 16 ? (blocks ? blocks : 1) : 148 * 16);
}

}  // namespace

int launch_lstm_bwd(const LstmBwdParams& p, cudaStream_t stream) {
  lstm_bwd_kernel<<<grid_for(p.n_pix * p.Co / 8, 256), 256, 0, stream>>>(p);
  return static_cast<int>(cudaGetLastError());
}
int launch_convert_scale(const ConvertParams& p, cudaStream_t stream) {
  convert_scale_kernel<<<grid_for(p.n / 8, 256), 256, 0, stream>>>(p);
  return static_cast<int>(cudaGetLastError());
}
int launch_pack_dgrad_weights(const float* w_in, const float* w_h, int Cin, int Co, int in_order,
                              int out_order, __half* out_hi, __half* out_lo, float* inv_scale_out,
                              unsigned int* scratch_amax, cudaStream_t stream) {
  cudaError_t e = cudaMemsetAsync(scratch_amax, 0, sizeof(unsigned int), stream);
  if (e != cudaSuccess) return static_cast<int>(e);
  const size_t na = static_cast<size_t>(4 * Co) * Cin * 9, nb = static_cast<size_t>(4 * Co) * Co * 9;
  amax2_kernel<<<296, 256, 0, stream>>>(w_in, na, w_h, nb, scratch_amax);
  pack_dgrad_kernel<<<1184, 256, 0, stream>>>(w_in, w_h, Cin, Co, in_order, out_order, scratch_amax, out_hi,
                                              out_lo, inv_scale_out);
  return static_cast<int>(cudaGetLastError());
}
int launch_unpack_wgrad(const float* gw, int Cin, int Co, int in_order, int out_order, float* dw_in,
                        float* dw_h, cudaStream_t stream) {
  unpack_wgrad_kernel<<<1184, 256, 0, stream>>>(gw, Cin, Co, in_order, out_order, dw_in, dw_h);
  return static_cast<int>(cudaGetLastError());
}
int launch_bottleneck_bwd(const BottleneckBwdParams& p, cudaStream_t stream) {
  const size_t smem = (static_cast<size_t>(2) * p.HW * p.Ch + 2 * p.HW * p.code) * sizeof(float);
  bottleneck_bwd_kernel<<<p.B_pad, 128, smem, stream>>>(p);
  return static_cast<int>(cudaGetLastError());
}
int launch_tail_bwd(const TailBwdParams& p, cudaStream_t stream) {
  const int HW = p.H * p.W;
  if (HW % TB_THREADS != 0 || p.C > MAXC) return static_cast<int>(cudaErrorInvalidValue);
  tail_bwd_kernel<<<p.B_pad * (HW / TB_THREADS), TB_THREADS, 0, stream>>>(p);
  return static_cast<int>(cudaGetLastError());
}
int launch_head_bwd(const HeadBwdParams& p, cudaStream_t stream) {
  const int HW = p.H * p.W;
  if (HW % TB_THREADS != 0 || p.C > MAXC) return static_cast<int>(cudaErrorInvalidValue);
  head_bwd_kernel<<<p.B_pad * (HW / TB_THREADS), TB_THREADS, 0, stream>>>(p);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace drasic
