// Multi-tensor Adam: the optimizer step of the reference's training loop (train_model.py:116 `optimizer.step()`
// on the optim.Adam built at train_model.py:155-161: L2 weight decay added to the gradient, bias-corrected
// first / second moments, no amsgrad) for every parameter of the codec in a few launches, so the training
// step stays on the device from the loss to the updated fp32 master weights.
//
// Up to ADAM_MAX_TENSORS tensors ride in the kernel arguments (no device-side table, no host sync); the
// work is cut into chunks of ADAM_CHUNK elements, one block per chunk.  HBM-bound: 16 B read + 12 B written
// per parameter.
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>

#include "optim.h"

namespace drasic {

namespace {

struct AdamBatch {
  float* param[ADAM_MAX_TENSORS];
  const float* grad[ADAM_MAX_TENSORS];
  float* m[ADAM_MAX_TENSORS];
  float* v[ADAM_MAX_TENSORS];
  long long numel[ADAM_MAX_TENSORS];
  int chunk_start[ADAM_MAX_TENSORS + 1];   // first chunk of tensor i; chunk_start[n] = chunks of the launch
  int n;
};

struct AdamMath {
  float beta2, one_minus_beta1, one_minus_beta2, eps, weight_decay, step_size, bc2_sqrt;
  __device__ __forceinline__ void operator()(float& p, float g, float& m, float& v) const {
    if (weight_decay != 0.0f) g = __fmaf_rn(p, weight_decay, g);
    m = __fmaf_rn(__fsub_rn(g, m), one_minus_beta1, m);                 // lerp(m, g, 1 - beta1)
    v = __fmaf_rn(__fmul_rn(one_minus_beta2, g), g, __fmul_rn(beta2, v));
    const float denom = __fadd_rn(__fdiv_rn(sqrtf(v), bc2_sqrt), eps);
    p = __fsub_rn(p, __fdiv_rn(__fmul_rn(step_size, m), denom));
  }
};

__global__ void __launch_bounds__(ADAM_THREADS) adam_kernel(const __grid_constant__ AdamBatch b, const AdamMath f) {
  // tensor of this chunk: last i with chunk_start[i] <= blockIdx.x
  int lo = 0, hi = b.n;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (b.chunk_start[mid] <= static_cast<int>(blockIdx.x)) lo = mid; else hi = mid;
  }
  const long long off = static_cast<long long>(blockIdx.x - b.chunk_start[lo]) * ADAM_CHUNK;
  const long long left = b.numel[lo] - off;
  const int n = left < ADAM_CHUNK ? static_cast<int>(left) : ADAM_CHUNK;
  float* p = b.param[lo] + off;
  const float* g = b.grad[lo] + off;
  float* m = b.m[lo] + off;
  float* v = b.v[lo] + off;
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15u) == 0;
  int done = 0;
  if (vec) {
    const int n4 = n >> 2;
    for (int i = threadIdx.x; i < n4; i += ADAM_THREADS) {
      float4 p4 = reinterpret_cast<float4*>(p)[i];
      const float4 g4 = reinterpret_cast<const float4*>(g)[i];
      float4 m4 = reinterpret_cast<float4*>(m)[i];
      float4 v4 = reinterpret_cast<float4*>(v)[i];
      f(p4.x, g4.x, m4.x, v4.x);
      f(p4.y, g4.y, m4.y, v4.y);
      f(p4.z, g4.z, m4.z, v4.z);
      f(p4.w, g4.w, m4.w, v4.w);
      reinterpret_cast<float4*>(p)[i] = p4;
      reinterpret_cast<float4*>(m)[i] = m4;
      reinterpret_cast<float4*>(v)[i] = v4;
    }
    done = n4 << 2;
  }
  for (int i = done + threadIdx.x; i < n; i += ADAM_THREADS) {
    float pi = p[i], mi = m[i], vi = v[i];
    f(pi, g[i], mi, vi);
    p[i] = pi; m[i] = mi; v[i] = vi;
  }
}

}  // namespace

int launch_adam(const AdamTensor* tensors, int n_tensors, const AdamHyper& h, cudaStream_t stream) {
  AdamMath f;
  // scalars are formed in double and rounded once, as torch forms them in Python before its fp32 tensor ops
  f.beta2 = static_cast<float>(h.beta2);
  f.one_minus_beta1 = static_cast<float>(1.0 - h.beta1);
  f.one_minus_beta2 = static_cast<float>(1.0 - h.beta2);
  f.eps = static_cast<float>(h.eps);
  f.weight_decay = static_cast<float>(h.weight_decay);
  f.step_size = static_cast<float>(h.lr / h.bias_correction1);
  f.bc2_sqrt = static_cast<float>(sqrt(h.bias_correction2));
  for (int first = 0; first < n_tensors; first += ADAM_MAX_TENSORS) {
    AdamBatch b;
    b.n = 0;
    int chunks = 0;
    for (int i = first; i < n_tensors && b.n < ADAM_MAX_TENSORS; ++i) {
      const AdamTensor& t = tensors[i];
      if (t.numel <= 0) continue;
      b.param[b.n] = t.param; b.grad[b.n] = t.grad; b.m[b.n] = t.exp_avg; b.v[b.n] = t.exp_avg_sq;
      b.numel[b.n] = t.numel;
      b.chunk_start[b.n] = chunks;
      chunks += static_cast<int>((t.numel + ADAM_CHUNK - 1) / ADAM_CHUNK);
      ++b.n;
    }
    if (b.n == 0) continue;
    for (int i = b.n; i <= ADAM_MAX_TENSORS; ++i) b.chunk_start[i] = chunks;
    adam_kernel<<<chunks, ADAM_THREADS, 0, stream>>>(b, f);
    const int e = static_cast<int>(cudaGetLastError());
    if (e) return e;
  }
  return 0;
}

}  // namespace drasic
