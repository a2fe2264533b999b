// Parameter blocks of the multi-tensor Adam step (optim.cu).
#pragma once
#include <cuda_runtime.h>

namespace drasic {

constexpr int ADAM_MAX_TENSORS = 64;     // tensors per launch (their pointers travel as kernel arguments)
constexpr int ADAM_CHUNK = 16384;        // elements per block
constexpr int ADAM_THREADS = 256;

struct AdamTensor {            // same layout as drasic_adam_tensor (include/drasic_b200.h)
  float* param;
  const float* grad;
  float* exp_avg;
  float* exp_avg_sq;
  long long numel;
};

struct AdamHyper {
  double lr, beta1, beta2, eps, weight_decay;
  double bias_correction1, bias_correction2;   // 1 - beta^step, computed by the caller in double like torch
};

int launch_adam(const AdamTensor* tensors, int n_tensors, const AdamHyper& h, cudaStream_t stream);

}  // namespace drasic
