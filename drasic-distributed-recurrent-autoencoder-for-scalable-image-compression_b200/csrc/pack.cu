// Weight packing for the gate kernel: fp32 OIHW conv weights of one ConvLSTMCell
// (modules/cell.py:82-100: cell[0]['in'] [4Co,Cin,3,3] and cell[0]['hidden'] [4Co,Co,3,3], no bias)
// -> K-major fp16 {hi, lo} matrix [4Co rows, 9*Cin + 9*Co], stored tile by tile ([N tile][K block of 64][block_n rows][64]),
//    in the tile order the kernel expects:
//   row  r = nt*BLOCK_N + gate*SLOTS + s     (SLOTS = BLOCK_N/4, physical channel p = nt*SLOTS + s)
//   col  k = [x part: tap*Cin + pin] [h part: 9*Cin + tap*Co + ph]
// Physical->logical channel maps follow ChanOrder (kernels.h). Values are multiplied by a per-tensor
// power of two so that max|w| lands in [1024, 2048) (exact; undone in the gate epilogue).
#include "kernels.h"

namespace drasic {

namespace {

__device__ __forceinline__ int logical_of(int p, int C, int order) {
  if (order == ORDER_QUAD) {
    const int cq = C >> 2;
    const int q = p / cq, c = p - q * cq;
    return c * 4 + q;
  }
  return p;
}

__global__ void amax_kernel(const __grid_constant__ PackBatch b, size_t na, size_t nb, unsigned int* __restrict__ out) {
  const float* __restrict__ wa = b.w_in[blockIdx.y];
  const float* __restrict__ wb = b.w_h[blockIdx.y];
  float m = 0.0f;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < na + nb;
       i += stride) {
    const float v = i < na ? wa[i] : wb[i - na];
    m = fmaxf(m, fabsf(v));
  }
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out + b.slot[blockIdx.y], __float_as_uint(m));  // non-negative floats order as uints
}

// One packed row (output channel) per block: the row's OIHW weights -- Cin*9 + Co*9 contiguous floats -- are staged in
// shared memory with coalesced reads, then written out in packed order, 64 consecutive K elements = one 128-byte run.
__global__ void __launch_bounds__(256) pack_kernel(const __grid_constant__ PackBatch b, int Cin, int Co, int in_order,
                                                   int out_order, int block_n, const unsigned int* __restrict__ amax_all,
                                                   __half* __restrict__ hi_all, __half* __restrict__ lo_all,
                                                   float* __restrict__ inv_all) {
  extern __shared__ float s_row[];   // [Cin*9 | Co*9]
  const int slot = b.slot[blockIdx.y];
  const int K = 9 * (Cin + Co);
  const size_t total = static_cast<size_t>(4 * Co) * K;
  __half* __restrict__ out_hi = hi_all + static_cast<size_t>(slot) * total;
  __half* __restrict__ out_lo = lo_all + static_cast<size_t>(slot) * total;
  const float amax = __uint_as_float(amax_all[slot]);
  int shift = 0;
  if (amax > 0.0f && isfinite(amax)) {
    int e;
    frexpf(amax, &e);  // amax = m * 2^e, m in [0.5,1)  ->  amax * 2^(11-e) in [1024, 2048)
    shift = 11 - e;
  }
  const float scale = ldexpf(1.0f, shift);
  if (blockIdx.x == 0 && threadIdx.x == 0) inv_all[slot] = ldexpf(1.0f, -(shift + ACT_SHIFT));
  const int slots = block_n >> 2;
  const int r = blockIdx.x;                       // packed row
  const int nt = r / block_n;
  const int rr = r - nt * block_n;
  const int gate = rr / slots;
  const int sl = rr - gate * slots;
  const int co = logical_of(nt * slots + sl, Co, out_order);
  const int orow = gate * Co + co;  // gates.chunk(4, 1): i, f, g, o blocks (cell.py:133)
  const float* __restrict__ src_in = b.w_in[blockIdx.y] + static_cast<size_t>(orow) * Cin * 9;
  const float* __restrict__ src_h = b.w_h[blockIdx.y] + static_cast<size_t>(orow) * Co * 9;
  for (int i = threadIdx.x; i < Cin * 9; i += blockDim.x) s_row[i] = src_in[i];
  for (int i = threadIdx.x; i < Co * 9; i += blockDim.x) s_row[Cin * 9 + i] = src_h[i];
  __syncthreads();
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    float v;
    if (k < 9 * Cin) {
      const int tap = k / Cin, pin = k - tap * Cin;
      v = s_row[logical_of(pin, Cin, in_order) * 9 + tap];
    } else {
      const int kk = k - 9 * Cin;
      const int tap = kk / Co, ph = kk - tap * Co;
      v = s_row[Cin * 9 + logical_of(ph, Co, out_order) * 9 + tap];
    }
    v *= scale;
    const __half hi = __float2half_rn(v);
    // tile-contiguous layout [N tile][K block][block_n rows][64]: the (N tile, K block) operand tile one TMA box brings
    // is one contiguous run of block_n * 128 bytes (streams from HBM at full rate when the weights do not fit the L2)
    const size_t off = ((static_cast<size_t>(nt) * (K / BLOCK_K) + k / BLOCK_K) * block_n + rr) * BLOCK_K + (k % BLOCK_K);
    out_hi[off] = hi;
    out_lo[off] = __float2half_rn(v - __half2float(hi));
  }
}

}  // namespace

// packs the n groups of `b` into their planes (slot) of the [groups][4Co][9(Cin+Co)] arrays in two launches
int launch_pack_lstm_weights(const PackBatch& b, int n, int n_slots, int Cin, int Co, int in_order, int out_order, int block_n,
                             __half* out_hi, __half* out_lo, float* inv_scale_out, unsigned int* scratch_amax,
                             cudaStream_t stream) {
  cudaError_t e = cudaMemsetAsync(scratch_amax, 0, sizeof(unsigned int) * n_slots, stream);
  if (e != cudaSuccess) return static_cast<int>(e);
  const size_t na = static_cast<size_t>(4 * Co) * Cin * 9, nb = static_cast<size_t>(4 * Co) * Co * 9;
  const int bx = n >= 4 ? 148 : 296;
  amax_kernel<<<dim3(bx, n), 256, 0, stream>>>(b, na, nb, scratch_amax);
  pack_kernel<<<dim3(4 * Co, n), 256, static_cast<size_t>(9) * (Cin + Co) * sizeof(float), stream>>>(
      b, Cin, Co, in_order, out_order, block_n, scratch_amax, out_hi, out_lo, inv_scale_out);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace drasic
