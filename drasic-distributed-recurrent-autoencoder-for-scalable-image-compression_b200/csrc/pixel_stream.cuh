// Per-warp streaming of a contiguous fp32 feature tensor through shared memory for the HBM-bound
// pixel kernels (pointwise.cu K-C, backward.cu tail/head backward).  Each warp owns a ring of
// chunks (8*U pixels x 32 channels, U KB) filled by cp.async.bulk (the TMA engine) and tracked by
// mbarriers, so several KB per warp are in flight without holding registers.
#pragma once
#include "ptx.cuh"

namespace drasic {

template <int U, int STAGES>
struct WarpStream {
  static constexpr int CHUNK_F4 = 64 * U;            // float4 per chunk = 8*U pixels x 32 channels
  static constexpr int CHUNK_BYTES = CHUNK_F4 * 16;
  float4* buf;       // this warp's [STAGES][CHUNK_F4]
  uint64_t* bar;     // this warp's [STAGES]
  const char* src;   // chunk 0 in global memory (16-byte aligned)
  int n;             // chunks to stream

  __device__ __forceinline__ void issue(int chunk) const {
    const int st = chunk % STAGES;
    ptx::mbar_expect_tx(&bar[st], CHUNK_BYTES);
    ptx::bulk_load_1d(buf + st * CHUNK_F4, src + static_cast<size_t>(chunk) * CHUNK_BYTES, CHUNK_BYTES, &bar[st]);
  }
  __device__ __forceinline__ void start(int lane) const {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) ptx::mbar_init(&bar[s], 1);
      ptx::fence_mbar_init();
      for (int c = 0; c < STAGES && c < n; ++c) issue(c);
    }
    __syncwarp();
  }
  __device__ __forceinline__ void wait(int chunk) const {
    ptx::mbar_wait(&bar[chunk % STAGES], (chunk / STAGES) & 1);
  }
  // the two float4 (channels {4sub..4sub+3} and {16+4sub..}) of pixel u*8 + (lane>>2) of `chunk`
  __device__ __forceinline__ void read(int chunk, int u, int lane, float4& a, float4& b) const {
    const float4* c = buf + (chunk % STAGES) * CHUNK_F4 + (u * 8 + (lane >> 2)) * 8 + (lane & 3);
    // odd pixels read their second half first: the eight lanes of a quarter warp then hit 32 distinct banks
    if ((lane >> 2) & 1) { b = c[4]; a = c[0]; } else { a = c[0]; b = c[4]; }
  }
  // all lanes have read `chunk`: refill its stage with chunk + STAGES
  __device__ __forceinline__ void release(int chunk, int lane) const {
    __syncwarp();
    if (lane == 0 && chunk + STAGES < n) {
      ptx::fence_proxy_async();
      issue(chunk + STAGES);
    }
  }
};

}  // namespace drasic
