// HBM-bound fused kernels around the gate GEMMs of the DRASIC residual coding loop.
//
//  bottleneck_kernel  (K-B): encoder tail ConvCell 1x1 128->code + tanh (models/dist_codec.py:83-85)
//                            -> Quantize (functions/quantize.py:9-18) -> bit-pack (dist_codec.py:16-18)
//                            -> decoder head ConvCell 1x1 code->128 + tanh (dist_codec.py:89-91)
//  tail_head_kernel   (K-C): decoder tail ConvCell 1x1 32->C + tanh (dist_codec.py:101-103)
//                            -> (d+1)/2 at i==0, recon += d, loss, residual (dist_codec.py:56-59)
//                            -> encoder head ConvCell 1x1 C->32 (+bias) + tanh (dist_codec.py:71-73)
//                            -> pixel_unshuffle(2) store (dist_codec.py:74)
// The 1x1 convolutions here are a few hundred MACs per pixel: CUDA-core work. In precise mode their
// dot products accumulate in fp64 so this side adds at most one rounding to the reference result.
#include "kernels.h"
#include "pixel_stream.cuh"
#include "pointwise.h"

namespace drasic {

namespace {

__device__ __forceinline__ float tanh_fast_pw(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ int find_group(const int* ends, int n_groups, int idx) {
  int g = 0;
  while (g + 1 < n_groups && idx >= ends[g]) ++g;
  return g;
}

__device__ __forceinline__ void store_split(__half* hi, __half* lo, float v, bool with_lo) {
  v *= ACT_SCALE;
  const __half h = __float2half_rn(v);
  *hi = h;
  if (with_lo) *lo = __float2half_rn(v - __half2float(h));
}

// ------------------------------------------------------------------ K-B
// one CTA (128 threads) per tile of P = min(HW, 32) code pixels of one image slot of the node-sorted,
// padded batch (a 32x32 patch has HW = 16 code pixels: one CTA per image; an un-tiled 768x512 image has
// 6144: 192 CTAs per image)
__global__ void __launch_bounds__(128) bottleneck_kernel(const BottleneckParams p) {
  ptx::griddep_launch_dependents();
  ptx::griddep_wait();
  extern __shared__ float sm[];
  const int HW = p.HW, CH = p.Ch, code = p.code;
  const int P = HW < 32 ? HW : 32;
  const int tiles = HW / P;
  float* sh = sm;                       // [P][CH]
  float* scode = sm + P * CH;           // [P][code]
  const int slot = blockIdx.x / tiles;
  const int pix0 = (blockIdx.x - slot * tiles) * P;
  const int orig = p.perm[slot];
  if (orig < 0) return;  // padding slot
  const int ge = p.n_enc_groups > 1 ? find_group(p.group_img_end, p.n_enc_groups, slot) : 0;
  const int gd = p.n_dec_groups > 1 ? find_group(p.group_img_end, p.n_dec_groups, slot) : 0;
  const float* w4 = p.w_e4 + static_cast<size_t>(ge) * code * CH;
  const float* w0 = p.w_d0 + static_cast<size_t>(gd) * CH * code;
  const size_t img_codes = static_cast<size_t>(code) * HW;   // code elements per image, NCHW [code][HW]
  const int n_out = code * P;                                 // of this tile; o = k*P + local pixel

  if (p.bits_in) {
    // decode-only: the symbols of this tile come from the packed stream (MSB first over [code][HW])
    const uint8_t* bsrc = p.bits_in + static_cast<size_t>(orig) * (img_codes >> 3);
    for (int o = threadIdx.x; o < n_out; o += blockDim.x) {
      const int k = o / P, lp = o - k * P;
      const size_t e = static_cast<size_t>(k) * HW + pix0 + lp;
      const float q = ((bsrc[e >> 3] >> (7 - (e & 7))) & 1) ? 1.0f : -1.0f;
      p.code_out[static_cast<size_t>(orig) * img_codes + e] = q;
      scode[lp * code + k] = q;
    }
  } else {
  const float4* src = reinterpret_cast<const float4*>(p.h_e3 + (static_cast<size_t>(slot) * HW + pix0) * CH);
  for (int i = threadIdx.x; i < P * CH / 4; i += blockDim.x)
    reinterpret_cast<float4*>(sh)[i] = src[i];
  __syncthreads();

  // encoder tail + quantizer; o enumerates [code][P] of this tile, so a warp covers 32 consecutive
  // elements of the image's NCHW code tensor (P == 32), or two rows of 16 (P == HW == 16)
  for (int base = 0; base < n_out; base += blockDim.x) {
    const int o = base + threadIdx.x;
    float q = 0.0f;
    bool valid = o < n_out;
    const int k = o / P, lp = o - k * P;
    const size_t e = static_cast<size_t>(k) * HW + pix0 + lp;   // element of the image's code tensor
    if (valid) {
      const float* wr = w4 + k * CH;
      const float* hr = sh + lp * CH;
      float z;
      if (p.precise) {
        double acc = 0.0;
        for (int c = 0; c < CH; ++c) acc = fma(static_cast<double>(__ldg(wr + c)), static_cast<double>(hr[c]), acc);
        z = tanhf(static_cast<float>(acc));
      } else {
        float acc = 0.0f;
        for (int c = 0; c < CH; ++c) acc = fmaf(__ldg(wr + c), hr[c], acc);
        z = tanh_fast_pw(acc);
      }
      const size_t oidx = static_cast<size_t>(orig) * img_codes + e;
      if (p.z_out) p.z_out[oidx] = z;
      q = z;  // values outside [-1,1] and NaN pass through unchanged (quantize.py:16-17)
      if (p.noise) {
        // training: +1 with probability (z+1)/2 (quantize.py:11-14)
        const float u = p.noise[oidx];
        const float pr = (z + 1.0f) / 2.0f;
        if (pr >= u) q = 1.0f;
        else if (pr < u) q = -1.0f;
      } else {
        if (z >= 0.0f && z <= 1.0f) q = 1.0f;       // -0.0f >= 0 -> +1, as the reference
        else if (z >= -1.0f && z < 0.0f) q = -1.0f;
      }
      p.code_out[oidx] = q;
      scode[lp * code + k] = q;
    }
    // real bitstream: bit = (code > 0), numpy packbits order (MSB first) over the NCHW code tensor
    const unsigned int ballot = __ballot_sync(0xffffffffu, valid && q > 0.0f);
    if (p.bits_out && (threadIdx.x & 31) == 0 && valid) {
      const unsigned int rev = __brev(ballot);
      uint8_t* dst = p.bits_out + static_cast<size_t>(orig) * (img_codes >> 3);
      if (P == 32) {
        const size_t byte0 = e >> 3;                     // this warp: 32 pixels of code plane k
        for (int b = 0; b < 4; ++b) dst[byte0 + b] = static_cast<uint8_t>(rev >> (24 - 8 * b));
      } else {
        // P == HW (< 32): consecutive o are consecutive elements of the code tensor
        const int byte0 = o >> 3;
        const int nbytes = min(4, (n_out >> 3) - byte0);
        for (int b = 0; b < nbytes; ++b) dst[byte0 + b] = static_cast<uint8_t>(rev >> (24 - 8 * b));
      }
    }
  }
  }
  __syncthreads();

  // decoder head
  const size_t doff = (static_cast<size_t>(slot) * HW + pix0) * CH;
  __half* dhi = p.d1_in[0] + doff;
  __half* dlo = p.d1_in[1] + doff;
  for (int o = threadIdx.x; o < P * CH; o += blockDim.x) {
    const int lp = o / CH, m = o - lp * CH;
    const float* wr = w0 + m * code;
    const float* cr = scode + lp * code;
    float d;
    if (p.precise) {
      double acc = 0.0;
      for (int k = 0; k < code; ++k) acc = fma(static_cast<double>(__ldg(wr + k)), static_cast<double>(cr[k]), acc);
      d = tanhf(static_cast<float>(acc));
    } else {
      float acc = 0.0f;
      for (int k = 0; k < code; ++k) acc = fmaf(__ldg(wr + k), cr[k], acc);
      d = tanh_fast_pw(acc);
    }
    if (p.d0_out_f32) p.d0_out_f32[doff + o] = d;
    store_split(dhi + o, dlo + o, d, p.split != 0);
  }
}

// ------------------------------------------------------------------ K-C
// Four lanes per pixel, eight of the 32 feature channels per lane.  Pixels are enumerated quad-major
// (i = quad*4 + (dy*2+dx), quads row-major over H/2 x W/2), which is exactly the physical order of D3's
// hidden state (ORDER_QUAD) and of E1's input (pixel_unshuffle): both feature tensors are addressed as
// [(slot*HW + i)*32 + channel], so every warp-wide load / store of them is one contiguous run.
// The two 1x1 convolutions are fp32 FMA chains (8 terms per lane, then a two-level tree over the four
// lanes of the pixel): their rounding error is below that of the reference's own fp32 convolution.
// The next pass's global loads are issued before the current pass is computed (software prefetch).
constexpr int TAIL_THREADS = 256;
constexpr int TAIL_PIX = TAIL_THREADS / 4;  // pixels per pass
constexpr int MAX_C = 4;
constexpr int FEAT = 32;  // channels of the 32x32 feature map on both sides (dist_codec.py:71,101)

struct QuadMap {  // quad-major pixel index -> raster pixel offset
  int W, qw, shift;
  __device__ __forceinline__ explicit QuadMap(int W_) : W(W_), qw(W_ >> 1) {
    shift = (qw & (qw - 1)) == 0 ? 31 - __clz(qw) : -1;
  }
  __device__ __forceinline__ int raster(int i) const {
    const int quad = i >> 2, qd = i & 3;
    const int qy = shift >= 0 ? quad >> shift : quad / qw;
    const int qx = quad - qy * qw;
    return (2 * qy + (qd >> 1)) * W + 2 * qx + (qd & 1);
  }
};

constexpr int TAIL_U = 2;        // pixels per lane per pass (independent chains interleave)
constexpr int TAIL_STAGES = 2;   // chunks of 16 pixels (2 KB) in flight per warp

template <bool PRECISE, int NC>
__global__ void __launch_bounds__(TAIL_THREADS, 2) tail_head_kernel(const TailParams p) {
  ptx::griddep_launch_dependents();
  ptx::griddep_wait();
  constexpr int U = TAIL_U;
  using Stream = WarpStream<U, TAIL_STAGES>;
  __shared__ double s_red[2][TAIL_THREADS / 32];
  const int HW = p.H * p.W;
  const int bpi = gridDim.x / p.B_pad;  // blocks per image
  const int slot = blockIdx.x / bpi;
  const int part = blockIdx.x - slot * bpi;
  const int ppb = HW / bpi;             // pixels of this block (multiple of TAIL_PIX * U)
  const int orig = p.perm[slot];
  if (orig < 0) return;  // padding slot (uniform per block)
  constexpr int C = NC;
  const int ge = p.n_enc_groups > 1 ? find_group(p.group_img_end, p.n_enc_groups, slot) : 0;
  const int gd = p.n_dec_groups > 1 ? find_group(p.group_img_end, p.n_dec_groups, slot) : 0;
  const int sub = threadIdx.x & 3;                 // which channel slice of the pixel
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int lane_base = lane & ~3;
  const bool has_tail = p.mode != TAIL_INIT;
  const bool has_head = p.e1_in[0] != nullptr;
  const bool mine = sub < C;                       // this lane owns image channel `sub` of its pixel
  const bool step_mode = p.mode == TAIL_STEP;
  const bool has_img = p.img != nullptr;            // false = decode-only: no loss, no residual
  // heterogeneous rates: a node whose iteration budget is spent receives no further refinement
  const bool frozen = p.group_iters != nullptr && p.iter >= p.group_iters[ge];
  const QuadMap qm(p.W);

  // weights laid out per channel slice so a lane reads its eight channels with vector loads:
  // tail channels {4sub..4sub+3, 16+4sub..16+4sub+3}; head channels 8sub..8sub+7
  __shared__ alignas(16) float s_w4[MAX_C][4][8];   // [c][sub][j]
  __shared__ alignas(16) float s_w0[4][8][MAX_C];   // [sub][j][c]
  __shared__ alignas(16) float s_b0[4][8];
  for (int t = threadIdx.x; t < MAX_C * FEAT; t += TAIL_THREADS) {
    const int c = t >> 5, sb = (t >> 3) & 3, j = t & 7;
    const int kt = j < 4 ? 4 * sb + j : 16 + 4 * sb + (j - 4);
    s_w4[c][sb][j] = (has_tail && c < C) ? p.w_d4[(static_cast<size_t>(gd) * C + c) * FEAT + kt] : 0.0f;
    s_w0[sb][j][c] = (has_head && c < C) ? p.w_e0[(static_cast<size_t>(ge) * FEAT + 8 * sb + j) * C + c] : 0.0f;
    if (c == 0) s_b0[sb][j] = (has_head && p.b_e0) ? p.b_e0[static_cast<size_t>(ge) * FEAT + 8 * sb + j] : 0.0f;
  }
  __syncthreads();

  // each warp walks its own contiguous run of pixels, 8*U per pass; D3's h streams through shared memory
  __shared__ alignas(128) float4 s_feat[TAIL_THREADS / 32][TAIL_STAGES][Stream::CHUNK_F4];
  __shared__ uint64_t s_bar[TAIL_THREADS / 32][TAIL_STAGES];
  const int n_pass = ppb / (TAIL_PIX * U);
  const int i_warp = part * ppb + warp * (ppb / (TAIL_THREADS / 32));
  Stream ws;
  ws.buf = &s_feat[warp][0][0];
  ws.bar = &s_bar[warp][0];
  ws.src = reinterpret_cast<const char*>(p.d3_next) + (static_cast<size_t>(slot) * HW + i_warp) * FEAT * sizeof(float);
  ws.n = has_tail ? n_pass : 0;
  ws.start(lane);

  const size_t img_base = (static_cast<size_t>(orig) * C + (mine ? sub : 0)) * HW;
  // register prefetch of the small per-channel loads
  int i0 = i_warp + (lane >> 2);
  int rpix[U];
  float im[U], rp[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    rpix[u] = qm.raster(i0 + 8 * u);
    im[u] = rp[u] = 0.0f;
    if (mine) {
      if (has_img) im[u] = p.img[img_base + rpix[u]];
      if (step_mode) rp[u] = p.recon_prev[img_base + rpix[u]];
    }
  }

  double l_abs = 0.0, l_sq = 0.0;
  for (int pass = 0; pass < n_pass; ++pass, i0 += 8 * U) {
    float f[U][8];
    if (has_tail) {
      ws.wait(pass);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float4 fa, fb;
        ws.read(pass, u, lane, fa, fb);
        f[u][0] = fa.x; f[u][1] = fa.y; f[u][2] = fa.z; f[u][3] = fa.w;
        f[u][4] = fb.x; f[u][5] = fb.y; f[u][6] = fb.z; f[u][7] = fb.w;
      }
      ws.release(pass, lane);
    }
    float im_c[U], rp_c[U];
    size_t gi[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      im_c[u] = im[u]; rp_c[u] = rp[u];
      gi[u] = img_base + rpix[u];
    }
    if (pass + 1 < n_pass) {  // issue the next pass's loads now (uniform branch)
#pragma unroll
      for (int u = 0; u < U; ++u) {
        rpix[u] = qm.raster(i0 + 8 * U + 8 * u);
        if (mine) {
          if (has_img) im[u] = p.img[img_base + rpix[u]];
          if (step_mode) rp[u] = p.recon_prev[img_base + rpix[u]];
        }
      }
    }
    float xv[U];
    if (!has_tail) {
      // x = img * 2 - 1   (dist_codec.py:38)
#pragma unroll
      for (int u = 0; u < U; ++u) xv[u] = mine ? __fsub_rn(__fmul_rn(im_c[u], 2.0f), 1.0f) : 0.0f;
    } else {
      float pre[U];
      float pa = 0.0f, ps = 0.0f;   // this pass's loss terms (U elements), added to the fp64 sums once per pass
#pragma unroll
      for (int u = 0; u < U; ++u) pre[u] = 0.0f;
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        float w[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) w[j] = s_w4[c][sub][j];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          float a = 0.0f;
#pragma unroll
          for (int j = 0; j < 8; ++j) a = fmaf(w[j], f[u][j], a);
          a += __shfl_xor_sync(0xffffffffu, a, 1);
          a += __shfl_xor_sync(0xffffffffu, a, 2);
          pre[u] = sub == c ? a : pre[u];
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        xv[u] = 0.0f;
        if (mine) {
          float d = PRECISE ? tanhf(pre[u]) : tanh_fast_pw(pre[u]);
          if (p.dec_out) p.dec_out[gi[u]] = d;  // raw decoder output (saved for backward)
          float r;
          if (!step_mode) {
            d = __fadd_rn(d, 1.0f) / 2.0f;   // (decoded + 1) / 2   (dist_codec.py:56)
            if (frozen) d = 0.0f;
            r = __fadd_rn(0.0f, d);          // reconstructed = 0 + decoded (dist_codec.py:39,57)
          } else {
            if (frozen) d = 0.0f;
            r = __fadd_rn(rp_c[u], d);
          }
          p.recon_out[gi[u]] = r;
          if (!has_img) continue;
          const float e = __fsub_rn(r, im_c[u]);
          if (p.loss_kind == LOSS_MAE) pa += fabsf(e);
          else if (p.loss_kind == LOSS_MSE) pa = fmaf(e, e, pa);
          else if (r < 0.0f || r > 1.0f) {
            // F.binary_cross_entropy (dist_codec.py:26) raises for an input outside [0, 1]; the un-clamped cumulative
            // reconstruction leaves it after a few iterations.  The sum is poisoned here (NaN) and Model.forward raises.
            pa = __int_as_float(0x7fc00000);
          } else {  // bce with torch's log clamp at -100
            const float lr = fmaxf(logf(r), -100.0f), l1r = fmaxf(logf(1.0f - r), -100.0f);
            pa -= im_c[u] * lr + (1.0f - im_c[u]) * l1r;
          }
          ps = fmaf(e, e, ps);
          xv[u] = __fsub_rn(im_c[u], r);  // x = img - reconstructed   (dist_codec.py:59)
        }
      }
      l_abs += static_cast<double>(pa);
      l_sq += static_cast<double>(ps);
    }
    if (!has_head) continue;  // last iteration: no next encoder pass (uniform)
    // encoder head + pixel_unshuffle(2): E1 input [(slot*HW + i)*32 + k]
    float x[U][NC];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int c = 0; c < NC; ++c) x[u][c] = __shfl_sync(0xffffffffu, xv[u], lane_base + c);
    alignas(16) __half2 hi[U][4];
    alignas(16) __half2 lo[U][4];
    float ev[U][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float w[NC];
#pragma unroll
      for (int c = 0; c < NC; ++c) w[c] = s_w0[sub][j][c];
      const float b = s_b0[sub][j];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float a = b;
#pragma unroll
        for (int c = NC - 1; c >= 0; --c) a = fmaf(w[c], x[u][c], a);
        ev[u][j] = (PRECISE ? tanhf(a) : tanh_fast_pw(a)) * ACT_SCALE;
      }
    }
    // fp16 {hi, lo} split, two channels per conversion
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        hi[u][q] = __floats2half2_rn(ev[u][2 * q], ev[u][2 * q + 1]);
        const float2 hf = __half22float2(hi[u][q]);
        lo[u][q] = __floats2half2_rn(ev[u][2 * q] - hf.x, ev[u][2 * q + 1] - hf.y);
      }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t off = (static_cast<size_t>(slot) * HW + i0 + 8 * u) * FEAT + 8 * sub;
      *reinterpret_cast<uint4*>(p.e1_in[0] + off) = *reinterpret_cast<const uint4*>(hi[u]);
      if (p.split) *reinterpret_cast<uint4*>(p.e1_in[1] + off) = *reinterpret_cast<const uint4*>(lo[u]);
    }
  }

  if (has_tail && has_img) {
    // block reduction of the loss terms -> one fp64 atomic pair per block
    for (int o = 16; o > 0; o >>= 1) {
      l_abs += __shfl_xor_sync(0xffffffffu, l_abs, o);
      l_sq += __shfl_xor_sync(0xffffffffu, l_sq, o);
    }
    if ((threadIdx.x & 31) == 0) {
      s_red[0][threadIdx.x >> 5] = l_abs;
      s_red[1][threadIdx.x >> 5] = l_sq;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double a = 0.0, b = 0.0;
      for (int w = 0; w < TAIL_THREADS / 32; ++w) { a += s_red[0][w]; b += s_red[1][w]; }
      atomicAdd(&p.loss_acc[0], a);
      atomicAdd(&p.loss_acc[1], b);
    }
  }
}


// ------------------------------------------------------------------ stand-alone cells
// ConvCell with a 1x1 kernel called outside the fused loop (modules/cell.py:49-72): NCHW fp32 in/out.
__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case ACT_TANH: return tanhf(v);
    case ACT_RELU: return fmaxf(v, 0.0f);
    case ACT_SIGMOID: return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-v)));
    case ACT_HARDTANH: return fminf(fmaxf(v, -1.0f), 1.0f);
    case ACT_ELU: case ACT_CELU: return v > 0.0f ? v : expm1f(v);
    case ACT_SELU: return 1.0507009873554804934193349852946f * (v > 0.0f ? v : 1.6732632423543772848170429916717f * expm1f(v));
    default: return v;
  }
}

__global__ void conv1x1_kernel(const float* __restrict__ x, const float* __restrict__ w,
                               const float* __restrict__ b, float* __restrict__ y, int N, int Cin,
                               int Cout, int HW, int act) {
  const size_t total = static_cast<size_t>(N) * Cout * HW;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total; idx += stride) {
    const int pix = static_cast<int>(idx % HW);
    const int o = static_cast<int>((idx / HW) % Cout);
    const size_t n = idx / (static_cast<size_t>(HW) * Cout);
    const float* xr = x + n * Cin * HW + pix;
    const float* wr = w + static_cast<size_t>(o) * Cin;
    double acc = 0.0;
    for (int c = 0; c < Cin; ++c)
      acc = fma(static_cast<double>(__ldg(wr + c)), static_cast<double>(xr[static_cast<size_t>(c) * HW]), acc);
    if (b) acc += static_cast<double>(b[o]);
    y[idx] = apply_act(static_cast<float>(acc), act);
  }
}

// Quantize.forward (functions/quantize.py:9-18), element-wise; u == nullptr -> deterministic sign
__global__ void quantize_kernel(const float* __restrict__ x, const float* __restrict__ u,
                                float* __restrict__ y, size_t n) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += stride) {
    const float z = x[i];
    float q = z;
    if (u) {
      const float pr = (z + 1.0f) / 2.0f;
      if (pr >= u[i]) q = 1.0f;
      else if (pr < u[i]) q = -1.0f;
    } else {
      if (z >= 0.0f && z <= 1.0f) q = 1.0f;
      else if (z >= -1.0f && z < 0.0f) q = -1.0f;
    }
    y[i] = q;
  }
}

}  // namespace

int launch_bottleneck(const BottleneckParams& p, cudaStream_t stream) {
  const int P = p.HW < 32 ? p.HW : 32;
  if (p.HW % P || (p.code * P) % 32) return static_cast<int>(cudaErrorInvalidValue);
  const size_t smem = static_cast<size_t>(P) * (p.Ch + p.code) * sizeof(float);
  launch_pdl(bottleneck_kernel, dim3(p.B_pad * (p.HW / P)), dim3(128), smem, stream, p);
  return static_cast<int>(cudaGetLastError());
}

// blocks per image: enough blocks to fill the machine a few times over, each a whole number of passes
int pixel_blocks_per_image(int B_pad, int HW, int pix_per_pass, int min_blocks) {
  const int max_bpi = HW / pix_per_pass;
  int bpi = 1;
  while (bpi < max_bpi && static_cast<long long>(B_pad) * bpi < min_blocks) bpi *= 2;
  while (max_bpi % bpi) bpi /= 2;
  return bpi;
}

int launch_tail_head(const TailParams& p, cudaStream_t stream) {
  const int HW = p.H * p.W;
  if (HW % TAIL_THREADS != 0 || p.C > MAX_C || p.C < 1 || ((p.H | p.W) & 1)) return static_cast<int>(cudaErrorInvalidValue);
  const int grid = p.B_pad * pixel_blocks_per_image(p.B_pad, HW, TAIL_PIX * TAIL_U, 148 * 4);
#define DRASIC_TAIL_LAUNCH(NC)                                                            \
  if (p.precise) launch_pdl(tail_head_kernel<true, NC>, dim3(grid), dim3(TAIL_THREADS), 0, stream, p);        \
  else launch_pdl(tail_head_kernel<false, NC>, dim3(grid), dim3(TAIL_THREADS), 0, stream, p)
  switch (p.C) {
    case 1: DRASIC_TAIL_LAUNCH(1); break;
    case 2: DRASIC_TAIL_LAUNCH(2); break;
    case 3: DRASIC_TAIL_LAUNCH(3); break;
    default: DRASIC_TAIL_LAUNCH(4); break;
  }
#undef DRASIC_TAIL_LAUNCH
  return static_cast<int>(cudaGetLastError());
}

}  // namespace drasic

namespace drasic {
int launch_conv1x1(const float* x, const float* w, const float* b, float* y, int N, int Cin, int Cout,
                   int HW, int act, cudaStream_t stream) {
  const size_t total = static_cast<size_t>(N) * Cout * HW;
  const int blocks = static_cast<int>(total / 256 + 1 < 148 * 8 ? total / 256 + 1 : 148 * 8);
  conv1x1_kernel<<<blocks, 256, 0, stream>>>(x, w, b, y, N, Cin, Cout, HW, act);
  return static_cast<int>(cudaGetLastError());
}
int launch_quantize(const float* x, const float* u, float* y, size_t n, cudaStream_t stream) {
  const int blocks = static_cast<int>(n / 256 + 1 < 148 * 8 ? n / 256 + 1 : 148 * 8);
  quantize_kernel<<<blocks, 256, 0, stream>>>(x, u, y, n);
  return static_cast<int>(cudaGetLastError());
}
}  // namespace drasic
