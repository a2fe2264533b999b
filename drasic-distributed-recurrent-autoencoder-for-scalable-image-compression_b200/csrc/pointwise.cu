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
// one CTA (128 threads) per image slot of the node-sorted, padded batch
__global__ void __launch_bounds__(128) bottleneck_kernel(const BottleneckParams p) {
  extern __shared__ float sm[];
  const int HW = p.HW, CH = p.Ch, code = p.code;
  float* sh = sm;                       // [HW][CH]
  float* scode = sm + HW * CH;          // [HW][code]
  const int slot = blockIdx.x;
  const int orig = p.perm[slot];
  if (orig < 0) return;  // padding slot
  const int ge = p.n_enc_groups > 1 ? find_group(p.group_img_end, p.n_enc_groups, slot) : 0;
  const int gd = p.n_dec_groups > 1 ? find_group(p.group_img_end, p.n_dec_groups, slot) : 0;
  const float* w4 = p.w_e4 + static_cast<size_t>(ge) * code * CH;
  const float* w0 = p.w_d0 + static_cast<size_t>(gd) * CH * code;

  const float4* src = reinterpret_cast<const float4*>(p.h_e3 + static_cast<size_t>(slot) * HW * CH);
  for (int i = threadIdx.x; i < HW * CH / 4; i += blockDim.x)
    reinterpret_cast<float4*>(sh)[i] = src[i];
  __syncthreads();

  // encoder tail + quantizer; o enumerates the NCHW code tensor [code][HW] of this image
  const int n_out = code * HW;
  for (int base = 0; base < n_out; base += blockDim.x) {
    const int o = base + threadIdx.x;
    float q = 0.0f;
    bool valid = o < n_out;
    if (valid) {
      const int k = o / HW, pix = o - k * HW;
      const float* wr = w4 + k * CH;
      const float* hr = sh + pix * CH;
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
      const size_t oidx = static_cast<size_t>(orig) * n_out + o;
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
      scode[pix * code + k] = q;
    }
    // real bitstream: bit = (code > 0), numpy packbits order (MSB first) over the NCHW code tensor
    const unsigned int ballot = __ballot_sync(0xffffffffu, valid && q > 0.0f);
    if (p.bits_out && (threadIdx.x & 31) == 0 && base + (threadIdx.x & ~31) < n_out) {
      const unsigned int rev = __brev(ballot);
      const int byte0 = (base + threadIdx.x) >> 3;
      uint8_t* dst = p.bits_out + static_cast<size_t>(orig) * (n_out >> 3) + byte0;
      const int nbytes = min(4, (n_out >> 3) - byte0);
      for (int b = 0; b < nbytes; ++b) dst[b] = static_cast<uint8_t>(rev >> (24 - 8 * b));
    }
  }
  __syncthreads();

  // decoder head
  __half* dhi = p.d1_in[0] + static_cast<size_t>(slot) * HW * CH;
  __half* dlo = p.d1_in[1] + static_cast<size_t>(slot) * HW * CH;
  for (int o = threadIdx.x; o < HW * CH; o += blockDim.x) {
    const int pix = o / CH, m = o - pix * CH;
    const float* wr = w0 + m * code;
    const float* cr = scode + pix * code;
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
    if (p.d0_out_f32) p.d0_out_f32[static_cast<size_t>(slot) * HW * CH + o] = d;
    store_split(dhi + o, dlo + o, d, p.split != 0);
  }
}

// ------------------------------------------------------------------ K-C
constexpr int TAIL_THREADS = 256;
constexpr int MAX_C = 4;
constexpr int FEAT = 32;  // channels of the 32x32 feature map on both sides (dist_codec.py:71,101)

__global__ void __launch_bounds__(TAIL_THREADS) tail_head_kernel(const TailParams p) {
  __shared__ float s_w4[MAX_C * FEAT];
  __shared__ float s_w0[FEAT * MAX_C];
  __shared__ float s_b0[FEAT];
  __shared__ double s_red[2][TAIL_THREADS / 32];
  const int HW = p.H * p.W;
  const int blocks_per_img = HW / TAIL_THREADS;
  const int slot = blockIdx.x / blocks_per_img;
  const int pix = (blockIdx.x - slot * blocks_per_img) * TAIL_THREADS + threadIdx.x;
  const int orig = p.perm[slot];
  if (orig < 0) return;  // padding slot (uniform per block)
  const int C = p.C;
  const int ge = p.n_enc_groups > 1 ? find_group(p.group_img_end, p.n_enc_groups, slot) : 0;
  const int gd = p.n_dec_groups > 1 ? find_group(p.group_img_end, p.n_dec_groups, slot) : 0;
  if (threadIdx.x < C * FEAT) {
    if (p.mode != TAIL_INIT) s_w4[threadIdx.x] = p.w_d4[static_cast<size_t>(gd) * C * FEAT + threadIdx.x];
    s_w0[threadIdx.x] = p.w_e0[static_cast<size_t>(ge) * FEAT * C + threadIdx.x];
  }
  if (threadIdx.x < FEAT) s_b0[threadIdx.x] = p.b_e0 ? p.b_e0[static_cast<size_t>(ge) * FEAT + threadIdx.x] : 0.0f;
  __syncthreads();

  const int Y = pix / p.W, X = pix - Y * p.W;
  float x[MAX_C];
  double l_abs = 0.0, l_sq = 0.0;
  if (p.mode == TAIL_INIT) {
    // x = img * 2 - 1   (dist_codec.py:38)
    for (int c = 0; c < C; ++c) {
      const float im = p.img[(static_cast<size_t>(orig) * C + c) * HW + pix];
      x[c] = __fsub_rn(__fmul_rn(im, 2.0f), 1.0f);
    }
  } else {
    float f[FEAT];
    const float4* src = reinterpret_cast<const float4*>(p.d3_next + (static_cast<size_t>(slot) * HW + pix) * FEAT);
#pragma unroll
    for (int v = 0; v < FEAT / 4; ++v) {
      const float4 t = src[v];
      f[4 * v + 0] = t.x; f[4 * v + 1] = t.y; f[4 * v + 2] = t.z; f[4 * v + 3] = t.w;
    }
    for (int c = 0; c < C; ++c) {
      float d;
      if (p.precise) {
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < FEAT; ++k) acc = fma(static_cast<double>(s_w4[c * FEAT + k]), static_cast<double>(f[k]), acc);
        d = tanhf(static_cast<float>(acc));
      } else {
        float acc = 0.0f;
#pragma unroll
        for (int k = 0; k < FEAT; ++k) acc = fmaf(s_w4[c * FEAT + k], f[k], acc);
        d = tanh_fast_pw(acc);
      }
      const size_t gi = (static_cast<size_t>(orig) * C + c) * HW + pix;
      if (p.dec_out) p.dec_out[gi] = d;  // raw decoder output (saved for backward)
      float r;
      if (p.mode == TAIL_FIRST) {
        d = __fadd_rn(d, 1.0f) / 2.0f;   // (decoded + 1) / 2   (dist_codec.py:56)
        r = __fadd_rn(0.0f, d);          // reconstructed = 0 + decoded (dist_codec.py:39,57)
      } else {
        r = __fadd_rn(p.recon_prev[gi], d);
      }
      p.recon_out[gi] = r;
      const float im = p.img[gi];
      const float e = __fsub_rn(r, im);
      if (p.loss_kind == LOSS_MAE) l_abs += static_cast<double>(fabsf(e));
      else if (p.loss_kind == LOSS_MSE) l_abs += static_cast<double>(e) * e;
      else {  // bce with torch's log clamp at -100
        const float lr = fmaxf(logf(r), -100.0f), l1r = fmaxf(logf(1.0f - r), -100.0f);
        l_abs += -static_cast<double>(im * lr + (1.0f - im) * l1r);
      }
      l_sq += static_cast<double>(e) * e;
      x[c] = __fsub_rn(im, r);  // x = img - reconstructed   (dist_codec.py:59)
    }
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

  if (p.e1_in[0] == nullptr) return;  // last iteration: no next encoder pass
  // encoder head + pixel_unshuffle(2): E1 input [B,H/2,W/2,4*FEAT], physical channel q*FEAT + k
  alignas(16) __half hi[FEAT];
  alignas(16) __half lo[FEAT];
#pragma unroll
  for (int k = 0; k < FEAT; ++k) {
    float e;
    if (p.precise) {
      double acc = 0.0;
      for (int c = 0; c < C; ++c) acc = fma(static_cast<double>(s_w0[k * C + c]), static_cast<double>(x[c]), acc);
      e = tanhf(static_cast<float>(acc + static_cast<double>(s_b0[k])));
    } else {
      float acc = s_b0[k];
      for (int c = 0; c < C; ++c) acc = fmaf(s_w0[k * C + c], x[c], acc);
      e = tanh_fast_pw(acc);
    }
    e *= ACT_SCALE;
    hi[k] = __float2half_rn(e);
    lo[k] = __float2half_rn(e - __half2float(hi[k]));
  }
  const int qd = (Y & 1) * 2 + (X & 1);
  const size_t npix = (static_cast<size_t>(slot) * (p.H >> 1) + (Y >> 1)) * (p.W >> 1) + (X >> 1);
  const size_t off = npix * (4 * FEAT) + qd * FEAT;
  uint4* dh = reinterpret_cast<uint4*>(p.e1_in[0] + off);
#pragma unroll
  for (int v = 0; v < FEAT / 8; ++v) dh[v] = reinterpret_cast<const uint4*>(hi)[v];
  if (p.split) {
    uint4* dl = reinterpret_cast<uint4*>(p.e1_in[1] + off);
#pragma unroll
    for (int v = 0; v < FEAT / 8; ++v) dl[v] = reinterpret_cast<const uint4*>(lo)[v];
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
  const size_t smem = static_cast<size_t>(p.HW) * (p.Ch + p.code) * sizeof(float);
  bottleneck_kernel<<<p.B_pad, 128, smem, stream>>>(p);
  return static_cast<int>(cudaGetLastError());
}

int launch_tail_head(const TailParams& p, cudaStream_t stream) {
  const int HW = p.H * p.W;
  if (HW % TAIL_THREADS != 0 || p.C > MAX_C) return static_cast<int>(cudaErrorInvalidValue);
  tail_head_kernel<<<p.B_pad * (HW / TAIL_THREADS), TAIL_THREADS, 0, stream>>>(p);
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
