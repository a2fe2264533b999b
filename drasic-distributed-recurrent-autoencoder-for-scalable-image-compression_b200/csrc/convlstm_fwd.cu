// Implicit-GEMM 3x3 convolution kernel for sm_100a (tcgen05 tensor cores, TMA-fed, TMEM accumulators)
// with two fused epilogues:
//
//  EPI_GATES  ConvLSTM cell step: gates = W_in*x + W_h*h_prev, gate nonlinearities, cell / hidden
//             update and the pixel (un)shuffle of the following PixelShuffleCell.  Replaces, per cell
//             call, reference modules/cell.py:117 (input conv), :132 (hidden conv + add), :133-137
//             (chunk, sigmoid/tanh), :138-139 (c = f*c + i*g ; h = o*tanh(c)) and the PixelShuffleCell
//             after it in models/dist_codec.py:70-104 (functions/shuffle.py:4-25).
//  EPI_DGRAD  backward-data of the same cell (what autograd runs for train_model.py:115):
//             [dX | dH_prev] = conv3x3^T(dGates) with W_in and W_h concatenated along N, the input
//             gradient routed through the inverse (un)shuffle to the producing layer.
//
// GEMM view:  D[pixels, N] = A[pixels, 9*C] * Wp^T
//   A is never materialised: K block (src, tap, 64-channel chunk) of an M tile of 128 pixels is one
//   TMA box of an NHWC fp16 plane shifted by the tap offset (out-of-bounds = zero padding).
//   Wp is packed once per weight version (pack.cu), K-major.
// Numerics: "split" mode carries every operand as fp16 hi + fp16 lo (22 significand bits after
//   power-of-two pre-scaling) and issues lo*hi + hi*lo + hi*hi with fp32 accumulation (fp32-grade);
//   the non-split mode issues only hi*hi.
// Roles: warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc), warps 2..5 = epilogue.
// Persistent: each CTA walks tiles blockIdx.x, +gridDim.x, ...; two TMEM accumulators so the
//   epilogue of tile i overlaps the MMAs of tile i+1.
#include "kernels.h"
#include "ptx.cuh"

namespace drasic {

namespace {

constexpr int A_TILE_BYTES = TILE_M * BLOCK_K * 2;  // 16 KB
constexpr int SMEM_DATA_BUDGET = 196608;            // 192 KB of operand stages
constexpr int NUM_THREADS = 192;

// PAIR: two CTAs of a cluster (cta_group::2) share one 256-pixel x BLOCK_N tile: each CTA stages its own
// 128 pixels of A and half of the B rows, the leader issues M=256 MMAs that read both CTAs' shared memory.
// Per CTA and K block that is 32 KB instead of 48 KB of L2 -> SMEM traffic for the same MMA work.
template <bool SPLIT, int BLOCK_N, bool PAIR = false>
struct Cfg {
  static constexpr int B_ROWS = PAIR ? BLOCK_N / 2 : BLOCK_N;   // weight rows staged by one CTA
  static constexpr int B_TILE_BYTES = B_ROWS * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = (SPLIT ? 2 : 1) * (A_TILE_BYTES + B_TILE_BYTES);
  static constexpr int STAGES_RAW = SMEM_DATA_BUDGET / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr uint32_t TMEM_COLS = 2 * BLOCK_N;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

__device__ __forceinline__ float sigmoid_precise(float x) {
  // torch CPU: 1 / (1 + exp(-x)) with ~1 ulp exp and IEEE division
  return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x)));
}
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigmoid_fast(float x) {
  return fmaf(0.5f, tanh_fast(0.5f * x), 0.5f);
}

__device__ __forceinline__ void split_store16(__half* __restrict__ hi, __half* __restrict__ lo,
                                              const float (&v)[16], bool with_lo) {
  // v is already multiplied by ACT_SCALE
  alignas(16) __half h[16];
  alignas(16) __half l[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    h[i] = __float2half_rn(v[i]);
    l[i] = __float2half_rn(v[i] - __half2float(h[i]));
  }
  reinterpret_cast<uint4*>(hi)[0] = reinterpret_cast<const uint4*>(h)[0];
  reinterpret_cast<uint4*>(hi)[1] = reinterpret_cast<const uint4*>(h)[1];
  if (with_lo) {
    reinterpret_cast<uint4*>(lo)[0] = reinterpret_cast<const uint4*>(l)[0];
    reinterpret_cast<uint4*>(lo)[1] = reinterpret_cast<const uint4*>(l)[1];
  }
}
__device__ __forceinline__ void store_half16(__half* __restrict__ dst, const float (&v)[16]) {
  alignas(16) __half h[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) h[i] = __float2half_rn(v[i]);
  reinterpret_cast<uint4*>(dst)[0] = reinterpret_cast<const uint4*>(h)[0];
  reinterpret_cast<uint4*>(dst)[1] = reinterpret_cast<const uint4*>(h)[1];
}
__device__ __forceinline__ void store_f32x16(float* __restrict__ dst, const float (&v)[16]) {
#pragma unroll
  for (int q = 0; q < 4; ++q)
    reinterpret_cast<float4*>(dst)[q] = make_float4(v[4 * q + 0], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
}

// work unit index -> (N tile, M unit).  tile_order 0: N tile fastest (units running side by side share their
// pixels, i.e. the A operand); 1: M unit fastest (they share the weight tile, the B operand)
__device__ __forceinline__ void tile_split(const GateParams& p, int tile, int m_units, int& nt, int& mu) {
  if (p.tile_order == 0) { nt = tile % p.n_ntiles; mu = tile / p.n_ntiles; }
  else { mu = tile % m_units; nt = tile / m_units; }
}

// Work units of the M dimension.  A node's slots are a multiple of 4 images, so at the 4x4 layers (one M tile =
// 8 images) a tile can hold images of two nodes: every node walks ALL tiles that overlap its slot range
// (group_mtile_start / group_mtile_end, ranges of neighbouring nodes may share a tile) with its own weights
// and the epilogue stores only the rows of its own images (group_img_end).  Units are single tiles, or -- CTA
// pairs -- two consecutive tiles of the same node; group_unit_end[] holds the running unit counts.  When a
// node has an odd number of tiles the second CTA of its last pair has no tile of its own: it stages the first
// CTA's pixels again (the cta_group::2 MMA reads both CTAs' shared memory) and stores nothing.
struct MUnit {
  int g;                // node (weight group)
  int img_lo, img_hi;   // slots whose rows this unit may store
  bool own;             // false: filler CTA of an odd pair
  bool pos;             // position-major tile: row r = image n0 + r at pixel (y0, x0)
  int n0, y0, x0;       // first image / pixel of the tile this CTA stages
  uint32_t taps;        // bit t: tap t (dy = t/3 - 1, dx = t%3 - 1) is walked (all nine for pixel-major tiles)
};
// taps of the 3x3 stencil that are in bounds at pixel (y, x): bit t <-> (dy = t/3 - 1, dx = t%3 - 1)
__device__ __forceinline__ uint32_t pos_taps(const GateParams& p, int y, int x) {
  const uint32_t mx = (x > 0 ? 1u : 0u) | 2u | (x < p.W - 1 ? 4u : 0u);
  return (y > 0 ? mx : 0u) | (mx << 3) | (y < p.H - 1 ? mx << 6 : 0u);
}
template <bool PAIR>
__device__ __forceinline__ MUnit m_unit(const GateParams& p, int mu, int rank, int tiles_x, int tiles_per_img) {
  MUnit u;
  if (mu < p.pos_units) {
    // position-major: both CTAs of a pair sit on the same position (same taps), adjacent blocks of 128 images
    int c = 0;
    while (c + 1 < p.pos_classes && mu >= p.pos_class_unit0[c + 1]) ++c;
    const int l = mu - p.pos_class_unit0[c];
    const int it = l / p.pos_class_n[c];                     // block of images: slow inside the class
    const int e = p.pos_class_pos0[c] + (l - it * p.pos_class_n[c]);
    int pos;
    if (p.pos_pair_mode) {   // two positions, one per CTA of the pair, on the same 128 images
      const int pa = __ldg(&p.pos_order[2 * e]), pb = __ldg(&p.pos_order[2 * e + 1]);
      pos = rank ? pb : pa;
      u.taps = pos_taps(p, pa / p.W, pa % p.W) | pos_taps(p, pb / p.W, pb % p.W);
      u.n0 = it * TILE_M;
    } else {                 // one position; the CTAs of a pair take adjacent blocks of 128 images
      pos = __ldg(&p.pos_order[e]);
      u.taps = pos_taps(p, pos / p.W, pos % p.W);
      u.n0 = (PAIR ? 2 * it + rank : it) * TILE_M;
    }
    u.y0 = pos / p.W;
    u.x0 = pos - u.y0 * p.W;
    u.g = 0; u.img_lo = 0; u.img_hi = 0x7fffffff; u.own = true; u.pos = true;
    return u;
  }
  mu -= p.pos_units;
  int first, end;
  if (p.n_groups > 1) {
    int g = 0;
    while (g + 1 < p.n_groups && mu >= p.group_unit_end[g]) ++g;
    const int unit0 = g ? p.group_unit_end[g - 1] : 0;
    first = p.group_mtile_start[g] + (PAIR ? 2 : 1) * (mu - unit0);
    end = p.group_mtile_end[g];
    u.g = g;
    u.img_lo = g ? p.group_img_end[g - 1] : 0;
    u.img_hi = p.group_img_end[g];
  } else {
    first = p.pos_tile_base + (PAIR ? 2 : 1) * mu;
    end = p.n_mtiles;
    u.g = 0;
    u.img_lo = 0;
    u.img_hi = 0x7fffffff;
  }
  u.own = first + rank < end;
  const int mt = u.own ? first + rank : first;
  u.pos = false;
  u.taps = 0x1ffu;
  if (p.bn > 1) {
    u.n0 = mt * p.bn; u.y0 = 0; u.x0 = 0;
  } else {
    u.n0 = mt / tiles_per_img;
    const int rem = mt % tiles_per_img;
    u.y0 = (rem / tiles_x) * p.bh;
    u.x0 = (rem % tiles_x) * p.bw;
  }
  return u;
}

// ---------------------------------------------------------------- schedule
// One piece of work of a worker (CTA, or CTA pair): K blocks [kb_lo, kb_hi) of work item (mu, nt).  The static
// schedule deals whole items round robin; the stream-K schedule gives every worker one contiguous range of the
// global K-block sequence, so its first and last segments may be parts of items.
struct Seg {
  int nt, mu;
  int kb_lo, kb_hi, nkb;   // nkb: K blocks of the whole item
  int g0;                  // stream-K: global index of the item's first K block
};
struct Walk {
  int item, stride, total;   // static schedule
  int g, g_end;              // stream-K
  int cpt, n_workers;
  __device__ __forceinline__ static int range_begin(const GateParams& p, int worker, int n_workers) {
    return static_cast<int>(static_cast<long long>(worker) * p.sk_kb0[p.sk_classes] / n_workers);
  }
  __device__ __forceinline__ Walk(const GateParams& p, int worker, int n_workers_, int chunks_per_tap) {
    cpt = chunks_per_tap;
    n_workers = n_workers_;
    item = worker; stride = n_workers_; total = p.n_munits * p.n_ntiles;
    g = g_end = 0;
    if (p.sk_on) {
      total = p.sk_static;
      g = range_begin(p, worker, n_workers_);
      g_end = range_begin(p, worker + 1, n_workers_);
    }
  }
  __device__ __forceinline__ bool next(const GateParams& p, Seg& s) {
    int it;
    if (item >= total) {   // whole items dealt: the stream-K range, if any
      if (g >= g_end) return false;
      int k = 0;
      while (k + 1 < p.sk_classes && g >= p.sk_kb0[k + 1]) ++k;
      const int off = g - p.sk_kb0[k], n = p.sk_nkb[k];
      const int q = off / n;
      it = p.sk_item0[k] + q;
      s.kb_lo = off - q * n;
      s.nkb = n;
      s.kb_hi = min(n, s.kb_lo + (g_end - g));
      s.g0 = g - s.kb_lo;
      g += s.kb_hi - s.kb_lo;
      tile_split(p, it, p.n_munits, s.nt, s.mu);
      return true;
    }
    it = item;
    item += stride;
    tile_split(p, it, p.n_munits, s.nt, s.mu);
    int taps = 9;
    if (s.mu < p.pos_units) {
      int c = 0;
      while (c + 1 < p.pos_classes && s.mu >= p.pos_class_unit0[c + 1]) ++c;
      taps = p.pos_class_taps[c];
    }
    s.nkb = taps * cpt;
    s.kb_lo = 0; s.kb_hi = s.nkb; s.g0 = 0;
    return true;
  }
};

// stream-K partial accumulators: slot of CTA `rank` of `worker`, laid out [16-column chunk][row][16] so that a warp
// (32 consecutive rows) moves 2 KB contiguous per chunk
__device__ __forceinline__ float* sk_slot(const GateParams& p, int worker, int ranks, int rank) {
  return p.sk_scratch + (static_cast<size_t>(worker) * ranks + rank) * SK_SLOT_FLOATS;
}
__device__ __forceinline__ void sk_add16(float (&v)[16], const float* __restrict__ part, int n_peers, size_t peer_stride,
                                         int chunk, int row) {
  for (int pp = 0; pp < n_peers; ++pp) {
    const float4* src = reinterpret_cast<const float4*>(part + pp * peer_stride + (static_cast<size_t>(chunk) * TILE_M + row) * 16);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 t = __ldcg(src + q);
      v[4 * q + 0] += t.x; v[4 * q + 1] += t.y; v[4 * q + 2] += t.z; v[4 * q + 3] += t.w;
    }
  }
}
__device__ __forceinline__ uint32_t ld_acquire_u32(const unsigned int* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void epilogue_bar_sync() {   // the four epilogue warps only
  asm volatile("bar.sync 1, 128;" ::: "memory");
}

// ---------------------------------------------------------------- epilogues (one tile row per thread)
// ConvLSTM gates: the N tile holds i,f,g,o of SLOTS channels (columns gate*SLOTS + s)
template <bool SPLIT, int BLOCK_N>
__device__ __forceinline__ void epilogue_gates(const GateParams& p, uint32_t t_row, int nt, int g, int n,
                                               int y, int x, bool row_ok, const float* part, int n_peers,
                                               size_t peer_stride, int row) {
  constexpr int SLOTS = BLOCK_N / 4;  // channels per N tile
  const int H = p.H, W = p.W, Co = p.Co;
  const float inv = __ldg(&p.w_inv_scale[g]);
  const size_t pix = (static_cast<size_t>(n) * H + y) * W + x;
#pragma unroll 1
  for (int j = 0; j < SLOTS / 16; ++j) {
    float gi[16], gf[16], gg[16], go[16];
    ptx::tmem_ld16(t_row + 0 * SLOTS + j * 16, gi);
    ptx::tmem_ld16(t_row + 1 * SLOTS + j * 16, gf);
    ptx::tmem_ld16(t_row + 2 * SLOTS + j * 16, gg);
    ptx::tmem_ld16(t_row + 3 * SLOTS + j * 16, go);
    ptx::tmem_ld_wait();
    if (!row_ok) continue;   // (the TMEM loads above are warp-collective; everything below is per row)
    if (n_peers) {           // stream-K: the other workers' parts of this item's K range
      sk_add16(gi, part, n_peers, peer_stride, 0 * (SLOTS / 16) + j, row);
      sk_add16(gf, part, n_peers, peer_stride, 1 * (SLOTS / 16) + j, row);
      sk_add16(gg, part, n_peers, peer_stride, 2 * (SLOTS / 16) + j, row);
      sk_add16(go, part, n_peers, peer_stride, 3 * (SLOTS / 16) + j, row);
    }
    const int p0 = nt * SLOTS + j * 16;  // first physical channel of this chunk
    float cv[16];
    if (p.has_cell) {
      const float* cp = (p.c_prev ? p.c_prev : p.c_state) + pix * Co + p0;
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const float4 t = reinterpret_cast<const float4*>(cp)[v];
        cv[4 * v + 0] = t.x; cv[4 * v + 1] = t.y; cv[4 * v + 2] = t.z; cv[4 * v + 3] = t.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) cv[i] = 0.0f;
    }
    float hv[16];
    if (p.precise) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        gi[i] = sigmoid_precise(gi[i] * inv);
        gf[i] = sigmoid_precise(gf[i] * inv);
        gg[i] = tanhf(gg[i] * inv);
        go[i] = sigmoid_precise(go[i] * inv);
        // reference rounding: separate multiplies and add (no fma), cell.py:138-139
        const float cn = __fadd_rn(__fmul_rn(gf[i], cv[i]), __fmul_rn(gi[i], gg[i]));
        cv[i] = cn;
        hv[i] = __fmul_rn(go[i], tanhf(cn));
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        gi[i] = sigmoid_fast(gi[i] * inv);
        gf[i] = sigmoid_fast(gf[i] * inv);
        gg[i] = tanh_fast(gg[i] * inv);
        go[i] = sigmoid_fast(go[i] * inv);
        const float cn = fmaf(gf[i], cv[i], gi[i] * gg[i]);
        cv[i] = cn;
        hv[i] = go[i] * tanh_fast(cn);
      }
    }
    store_f32x16(p.c_state + pix * Co + p0, cv);
    if (p.gates_out) {  // training: keep the activated gates for backward, [pix][gate][Co]
      if (p.gates_f32) {
        float* gp = reinterpret_cast<float*>(p.gates_out) + pix * (4 * Co) + p0;
        store_f32x16(gp, gi);
        store_f32x16(gp + Co, gf);
        store_f32x16(gp + 2 * Co, gg);
        store_f32x16(gp + 3 * Co, go);
      } else {
        __half* gp = reinterpret_cast<__half*>(p.gates_out) + pix * (4 * Co) + p0;
        store_half16(gp, gi);
        store_half16(gp + Co, gf);
        store_half16(gp + 2 * Co, gg);
        store_half16(gp + 3 * Co, go);
      }
    }
    // consumer copy (fused PixelShuffleCell)
    size_t noff = 0;
    if (p.next_mode == NEXT_SAME) {
      noff = pix * Co + p0;
    } else if (p.next_mode == NEXT_DOWN) {
      const size_t npix = (static_cast<size_t>(n) * (H >> 1) + (y >> 1)) * (W >> 1) + (x >> 1);
      noff = npix * (4 * Co) + ((y & 1) * 2 + (x & 1)) * Co + p0;
    } else if (p.next_mode == NEXT_UP) {
      const int cq = Co >> 2;
      const int qq = p0 / cq, cc = p0 - qq * cq;
      const size_t npix =
          (static_cast<size_t>(n) * (2 * H) + (2 * y + (qq >> 1))) * (2 * W) + (2 * x + (qq & 1));
      noff = npix * cq + cc;
    }
    if (p.next_mode != NEXT_NONE && p.next_f32) store_f32x16(reinterpret_cast<float*>(p.next[0]) + noff, hv);
#pragma unroll
    for (int i = 0; i < 16; ++i) hv[i] *= ACT_SCALE;
    split_store16(p.h_out[0] + pix * Co + p0, p.h_out[1] + pix * Co + p0, hv, SPLIT);
    if (p.next_mode != NEXT_NONE && !p.next_f32)
      split_store16(reinterpret_cast<__half*>(p.next[0]) + noff, reinterpret_cast<__half*>(p.next[1]) + noff,
                    hv, SPLIT);
  }
}

// dgrad: columns are output channels [dX (n_x) | dH_prev (n_out - n_x)], fp32 stores
template <bool SPLIT, int BLOCK_N>
__device__ __forceinline__ void epilogue_dgrad(const GateParams& p, uint32_t t_row, int nt, int g, int n,
                                               int y, int x, bool row_ok, const float* part, int n_peers,
                                               size_t peer_stride, int row) {
  const int H = p.H, W = p.W;
  const float inv = __ldg(&p.w_inv_scale[g]) * __ldg(p.a_inv_scale);
  const size_t pix = (static_cast<size_t>(n) * H + y) * W + x;
  const int col0 = nt * BLOCK_N;
  const int width = min(BLOCK_N, p.n_out - col0);
#pragma unroll 1
  for (int j = 0; j < width / 16; ++j) {
    float v[16];
    ptx::tmem_ld16(t_row + j * 16, v);
    ptx::tmem_ld_wait();
    if (!row_ok) continue;
    if (n_peers) sk_add16(v, part, n_peers, peer_stride, j, row);
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] *= inv;
    const int c = col0 + j * 16;
    if (c < p.n_x) {
      size_t off;
      if (p.dx_mode == DX_SAME) {
        off = pix * p.n_x + c;
      } else if (p.dx_mode == DX_INV_DOWN) {
        const int cp = p.n_x >> 2;  // producer's channel count
        const int qq = c / cp, cc = c - qq * cp;
        const size_t ppix =
            (static_cast<size_t>(n) * (2 * H) + (2 * y + (qq >> 1))) * (2 * W) + (2 * x + (qq & 1));
        off = ppix * cp + cc;
      } else {  // DX_INV_UP
        const int qq = (y & 1) * 2 + (x & 1);
        const size_t ppix = (static_cast<size_t>(n) * (H >> 1) + (y >> 1)) * (W >> 1) + (x >> 1);
        off = ppix * (4 * p.n_x) + qq * p.n_x + c;
      }
      store_f32x16(p.dx_out + off, v);
    } else if (p.dhrec_out) {
      store_f32x16(p.dhrec_out + pix * (p.n_out - p.n_x) + (c - p.n_x), v);
    }
  }
}

template <bool SPLIT, int BLOCK_N, int EPI, bool PAIR>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_gemm_kernel(const __grid_constant__ GateParams p) {
  using C = Cfg<SPLIT, BLOCK_N, PAIR>;
  // (a stream-K launch keeps to plain stream order on both sides: its workers wait for each other's partial
  // accumulators, and letting other grids share the SMs with it while they do was observed to hang)
  if (!p.sk_on) ptx::griddep_launch_dependents();
  const uint32_t cta_rank = PAIR ? ptx::cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B operand tiles need 1024-byte alignment
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* full_bar = bars;                    // [STAGES]
  uint64_t* empty_bar = bars + C::STAGES;       // [STAGES]
  uint64_t* tmem_full = bars + 2 * C::STAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;         // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tmem_full[a], 1);
      ptx::mbar_init(&tmem_empty[a], PAIR ? 8 : 4);  // one arrive per epilogue warp (of both CTAs of a pair)
    }
    ptx::fence_mbar_init();
    ptx::prefetch_tmap(&p.tm_x[0]);
    ptx::prefetch_tmap(&p.tm_w[0]);
    if (SPLIT) {
      ptx::prefetch_tmap(&p.tm_x[1]);
      ptx::prefetch_tmap(&p.tm_w[1]);
    }
    if (p.has_hidden) {
      ptx::prefetch_tmap(&p.tm_h[0]);
      if (SPLIT) ptx::prefetch_tmap(&p.tm_h[1]);
    }
    if (p.pos_units) {
      ptx::prefetch_tmap(&p.tm_xp[0]);
      if (SPLIT) ptx::prefetch_tmap(&p.tm_xp[1]);
      if (p.has_hidden) {
        ptx::prefetch_tmap(&p.tm_hp[0]);
        if (SPLIT) ptx::prefetch_tmap(&p.tm_hp[1]);
      }
    }
  }
  if (warp == 1) {
    if (PAIR) {
      ptx::tmem_alloc_pair(tmem_slot, C::TMEM_COLS);
      ptx::tmem_relinquish_pair();
    } else {
      ptx::tmem_alloc(tmem_slot, C::TMEM_COLS);
      ptx::tmem_relinquish();
    }
  }
  ptx::tc_fence_before();
  if (PAIR) ptx::cluster_sync(); else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // barriers, tensor memory and descriptor prefetches are set up: from here on the kernel reads what its predecessor wrote
  ptx::griddep_wait();

  // tile walk: a work unit is one M tile (1 CTA) or one pair of adjacent M tiles (CTA pair) x one N tile
  const int n_workers = PAIR ? gridDim.x >> 1 : gridDim.x;
  const int worker = PAIR ? blockIdx.x >> 1 : blockIdx.x;
  const int cx_chunks = p.Cin / BLOCK_K;
  const int ch_chunks = p.Co / BLOCK_K;
  const int chunks_per_tap = cx_chunks + (p.has_hidden ? ch_chunks : 0);
  const int tiles_x = p.W / p.bw;
  const int tiles_per_img = (p.H / p.bh) * tiles_x;

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      // the leader's full barriers (shared::cluster addresses; the CTA windows map linearly)
      const uint32_t full_remote0 = PAIR ? ptx::mapa_u32(&full_bar[0], 0) : 0u;
      Walk walk(p, worker, n_workers, chunks_per_tap);
      Seg sg;
      while (walk.next(p, sg)) {
        const int nt = sg.nt;
        const MUnit un = m_unit<PAIR>(p, sg.mu, static_cast<int>(cta_rank), tiles_x, tiles_per_img);
        // weight tiles are stored [group][N tile][K block][BLOCK_N rows][64]: row of K block 0 of this item's N tile
        const int wtile = (un.g * p.w_group_tiles + nt) * p.w_nkb;
        int whalf = 0;
        if (PAIR) {  // this CTA stages its half of the (possibly narrower, dgrad) N tile
          int width = BLOCK_N;
          if (EPI == EPI_DGRAD) width = min(BLOCK_N, p.n_out - nt * BLOCK_N);
          whalf = static_cast<int>(cta_rank) * (width >> 1);
        }
        // K blocks walk (source x | h, tap, 64-channel chunk); taps that lie in the zero padding for every row of a
        // position-major tile are skipped (the MMA issuer counts the same blocks); kb numbers the blocks of the item,
        // of which this segment covers [kb_lo, kb_hi)
        int kb = 0;
        for (int src = 0; src < (p.has_hidden ? 2 : 1); ++src) {
          const CUtensorMap* ta = un.pos ? (src ? p.tm_hp : p.tm_xp) : (src ? p.tm_h : p.tm_x);
          const int chunks = src ? ch_chunks : cx_chunks;
          const int kbase = src ? 9 * cx_chunks : 0;   // K block of the weight matrix where this source starts
          for (int tap = 0; tap < 9; ++tap) {
            if (!((un.taps >> tap) & 1u)) continue;
            const int dy = tap / 3 - 1, dx = tap - (tap / 3) * 3 - 1;
            if (kb + chunks <= sg.kb_lo || kb >= sg.kb_hi) { kb += chunks; continue; }
            for (int cc = 0; cc < chunks; ++cc, ++kb) {
              if (kb < sg.kb_lo || kb >= sg.kb_hi) continue;
              const int wrow = (wtile + kbase + tap * chunks + cc) * BLOCK_N + whalf;
              ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
              uint8_t* st = smem + stage * C::STAGE_BYTES;
              uint8_t* sb = st + (SPLIT ? 2 : 1) * A_TILE_BYTES;
              if (PAIR) {
                // both CTAs' bytes are accounted on the leader's barrier (its MMA thread consumes both halves)
                if (leader) ptx::mbar_expect_tx(&full_bar[stage], 2 * C::STAGE_BYTES);
                const uint32_t fb = full_remote0 + 8u * stage;
                ptx::tma_load_4d_pair(st, &ta[0], fb, cc * BLOCK_K, un.x0 + dx, un.y0 + dy, un.n0);
                if (SPLIT)
                  ptx::tma_load_4d_pair(st + A_TILE_BYTES, &ta[1], fb, cc * BLOCK_K, un.x0 + dx, un.y0 + dy, un.n0);
                ptx::tma_load_2d_pair(sb, &p.tm_w[0], fb, 0, wrow);
                if (SPLIT) ptx::tma_load_2d_pair(sb + C::B_TILE_BYTES, &p.tm_w[1], fb, 0, wrow);
              } else {
                ptx::mbar_expect_tx(&full_bar[stage], C::STAGE_BYTES);
                ptx::tma_load_4d(st, &ta[0], &full_bar[stage], cc * BLOCK_K, un.x0 + dx, un.y0 + dy, un.n0);
                if (SPLIT)
                  ptx::tma_load_4d(st + A_TILE_BYTES, &ta[1], &full_bar[stage], cc * BLOCK_K, un.x0 + dx,
                                   un.y0 + dy, un.n0);
                ptx::tma_load_2d(sb, &p.tm_w[0], &full_bar[stage], 0, wrow);
                if (SPLIT)
                  ptx::tma_load_2d(sb + C::B_TILE_BYTES, &p.tm_w[1], &full_bar[stage], 0, wrow);
              }
              if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    if (lane == 0 && leader) {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      const uint32_t smem0 = ptx::smem_u32(smem);
      const uint64_t desc_a_hi = ptx::umma_desc_sw128(smem0);
      const uint64_t desc_a_lo = ptx::umma_desc_sw128(smem0 + A_TILE_BYTES);
      const uint64_t desc_b_hi = ptx::umma_desc_sw128(smem0 + (SPLIT ? 2 : 1) * A_TILE_BYTES);
      const uint64_t desc_b_lo = ptx::umma_desc_sw128(smem0 + (SPLIT ? 2 : 1) * A_TILE_BYTES + C::B_TILE_BYTES);
      Walk walk(p, worker, n_workers, chunks_per_tap);
      Seg sg;
      for (; walk.next(p, sg); ++it) {
        // dgrad: the last N tile of a row may be narrower than BLOCK_N (UMMA N is a runtime field)
        int width = BLOCK_N;
        if (EPI == EPI_DGRAD) width = min(BLOCK_N, p.n_out - sg.nt * BLOCK_N);
        // K blocks of this segment: the taps the producer walks x the channel chunks of both sources
        const int nkb = sg.kb_hi - sg.kb_lo;
        const uint32_t idesc = ptx::umma_idesc_f16(PAIR ? 2 * TILE_M : TILE_M, width);
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < nkb; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          // descriptors differ from the stage-0 / k-0 ones only in the 14-bit start-address field
          const uint32_t st_off = static_cast<uint32_t>(stage) * (C::STAGE_BYTES >> 4);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint32_t off = st_off + k * ((UMMA_K * 2) >> 4);  // bytes inside the 128-byte swizzle row
            const uint64_t da_hi = desc_a_hi + off;
            const uint64_t db_hi = desc_b_hi + off;
            if (SPLIT) {
              const uint64_t da_lo = desc_a_lo + off;
              const uint64_t db_lo = desc_b_lo + off;
              // small terms first, then the dominant one
              if (PAIR) {
                ptx::umma_f16_pair(d_tmem, da_lo, db_hi, idesc, (kb | k) != 0);
                ptx::umma_f16_pair(d_tmem, da_hi, db_lo, idesc, 1);
                ptx::umma_f16_pair(d_tmem, da_hi, db_hi, idesc, 1);
              } else {
                ptx::umma_f16(d_tmem, da_lo, db_hi, idesc, (kb | k) != 0);
                ptx::umma_f16(d_tmem, da_hi, db_lo, idesc, 1);
                ptx::umma_f16(d_tmem, da_hi, db_hi, idesc, 1);
              }
            } else if (PAIR) {
              ptx::umma_f16_pair(d_tmem, da_hi, db_hi, idesc, (kb | k) != 0);
            } else {
              ptx::umma_f16(d_tmem, da_hi, db_hi, idesc, (kb | k) != 0);
            }
          }
          // smem slot reusable (in both CTAs of a pair) once these MMAs retire
          if (PAIR) ptx::umma_commit_pair(&empty_bar[stage]); else ptx::umma_commit(&empty_bar[stage]);
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        // accumulator complete (each CTA of a pair holds its own 128 rows)
        if (PAIR) ptx::umma_commit_pair(&tmem_full[acc]); else ptx::umma_commit(&tmem_full[acc]);
      }
    }
  } else {
    // ===================================================== epilogue (warps 2..5)
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    const int row = q * 32 + lane;
    const int xl = row % p.bw;
    const int yl = (row / p.bw) % p.bh;
    const int nl = row / (p.bw * p.bh);
    int it = 0;
    uint32_t tmem_empty_leader[2];
    if (PAIR)
      for (int a = 0; a < 2; ++a) tmem_empty_leader[a] = ptx::mapa_u32(&tmem_empty[a], 0);
    constexpr int RANKS = PAIR ? 2 : 1;
    Walk walk(p, worker, n_workers, chunks_per_tap);
    Seg sg;
    for (; walk.next(p, sg); ++it) {
      const int nt = sg.nt;
      const MUnit un = m_unit<PAIR>(p, sg.mu, static_cast<int>(cta_rank), tiles_x, tiles_per_img);
      const int g = un.g;
      // this thread's row: pixel (yl, xl) of image nl of the box, or image `row` of a position-major tile
      const int rn = un.n0 + (un.pos ? row : nl), ry = un.y0 + (un.pos ? 0 : yl), rx = un.x0 + (un.pos ? 0 : xl);
      // rows of another node's images (a shared 4x4 tile) and the filler CTA of an odd pair store nothing
      const bool row_ok = un.own && rn >= un.img_lo && rn < un.img_hi;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      ptx::mbar_wait(&tmem_full[acc], acc_phase);
      ptx::tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BLOCK_N;
      if (sg.kb_lo > 0) {
        // stream-K, not the item's first part: publish the raw accumulator for the worker that has the first part
        int width = BLOCK_N;
        if (EPI == EPI_DGRAD) width = min(BLOCK_N, p.n_out - nt * BLOCK_N);
        float* dst = sk_slot(p, worker, RANKS, static_cast<int>(cta_rank));
        for (int j = 0; j < width / 16; ++j) {
          float v[16];
          ptx::tmem_ld16(t_row + j * 16, v);
          ptx::tmem_ld_wait();
          store_f32x16(dst + (static_cast<size_t>(j) * TILE_M + row) * 16, v);
        }
        __threadfence();
        epilogue_bar_sync();
        if (threadIdx.x == 64) atomicExch(&p.sk_flags[worker * RANKS + cta_rank], 1u);
      } else {
        // stream-K, first part of an item that other workers finish: their partials were published at the start of
        // their ranges (or, for a worker whose whole range lies inside this item, at its end)
        int n_peers = 0;
        if (sg.kb_hi < sg.nkb) {
          while (worker + 1 + n_peers < n_workers && Walk::range_begin(p, worker + 1 + n_peers, n_workers) < sg.g0 + sg.nkb)
            ++n_peers;
          if (lane == 0) {
            for (int pp = 0; pp < n_peers; ++pp) {
              const unsigned int* f = &p.sk_flags[(worker + 1 + pp) * RANKS + cta_rank];
              uint64_t t0 = 0;
              for (uint32_t spin = 0; ld_acquire_u32(f) != 1u; ++spin) {
                __nanosleep(64);
                if ((spin & 0x3FFu) == 0x3FFu) {
                  const uint64_t now = ptx::globaltimer_ns();
                  if (t0 == 0) t0 = now;
                  else if (now - t0 > 4000000000ull) __trap();
                }
              }
            }
          }
          __syncwarp();
        }
        const float* part = sk_slot(p, worker + 1, RANKS, static_cast<int>(cta_rank));
        if (EPI == EPI_GATES)
          epilogue_gates<SPLIT, BLOCK_N>(p, t_row, nt, g, rn, ry, rx, row_ok, part, n_peers, RANKS * SK_SLOT_FLOATS, row);
        else
          epilogue_dgrad<SPLIT, BLOCK_N>(p, t_row, nt, g, rn, ry, rx, row_ok, part, n_peers, RANKS * SK_SLOT_FLOATS, row);
        if (n_peers) {   // every epilogue warp has read the partials: hand the slots back
          epilogue_bar_sync();
          if (threadIdx.x == 64)
            for (int pp = 0; pp < n_peers; ++pp) atomicExch(&p.sk_flags[(worker + 1 + pp) * RANKS + cta_rank], 0u);
        }
      }
      // release the accumulator to the MMA warp
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) ptx::mbar_arrive_cluster(tmem_empty_leader[acc]); else ptx::mbar_arrive(&tmem_empty[acc]);
      }
    }
  }

  ptx::tc_fence_before();
  if (PAIR) ptx::cluster_sync(); else __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    if (PAIR) ptx::tmem_dealloc_pair(tmem_base, C::TMEM_COLS); else ptx::tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

template <bool SPLIT, int BLOCK_N, int EPI, bool PAIR>
int configure_t() {
  using C = Cfg<SPLIT, BLOCK_N, PAIR>;
  auto kern = conv_gemm_kernel<SPLIT, BLOCK_N, EPI, PAIR>;
  static bool configured[64] = {};  // per instantiation and device (the attribute is per device)
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return static_cast<int>(cudaErrorInvalidDevice);
  if (!configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) return static_cast<int>(e);
    configured[dev] = true;
  }
  return 0;
}

template <bool SPLIT, int BLOCK_N, int EPI, bool PAIR>
void pair_launch_config(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr, int ctas, cudaStream_t stream) {
  using C = Cfg<SPLIT, BLOCK_N, PAIR>;
  cfg = cudaLaunchConfig_t{};
  cfg.gridDim = dim3(ctas);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = stream;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
}

// workers (CTAs, or CTA pairs) that are resident at the same time: the stream-K schedule lets a worker wait for the
// partial accumulators of the workers behind it, so it must never launch more than this
template <bool SPLIT, int BLOCK_N, int EPI, bool PAIR>
int max_workers_t(int num_sms) {
  if (configure_t<SPLIT, BLOCK_N, EPI, PAIR>()) return 0;
  const int cap = PAIR ? SK_MAX_WORKERS / 2 : SK_MAX_WORKERS;
  if (!PAIR) return num_sms < cap ? num_sms : cap;   // one CTA per SM (shared memory)
  static int cached[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
  if (!cached[dev]) {
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[2];
    pair_launch_config<SPLIT, BLOCK_N, EPI, PAIR>(cfg, attr, 2 * (num_sms / 2), nullptr);
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, conv_gemm_kernel<SPLIT, BLOCK_N, EPI, PAIR>, &cfg) != cudaSuccess) n = 0;
    (void)cudaGetLastError();
    cached[dev] = n > 0 ? n : -1;
  }
  const int n = cached[dev] > 0 ? cached[dev] : 0;
  return n < cap ? (n < num_sms / 2 ? n : num_sms / 2) : cap;
}

template <bool SPLIT, int BLOCK_N, int EPI, bool PAIR>
int launch_t(const GateParams& p, int num_sms, cudaStream_t stream) {
  using C = Cfg<SPLIT, BLOCK_N, PAIR>;
  auto kern = conv_gemm_kernel<SPLIT, BLOCK_N, EPI, PAIR>;
  if (int e = configure_t<SPLIT, BLOCK_N, EPI, PAIR>()) return e;
  const int total = p.n_munits * p.n_ntiles;
  int workers = PAIR ? num_sms / 2 : num_sms;
  if (total < workers) workers = total;
  if (p.sk_on) {
    workers = p.sk_workers;
    if (workers < 1 || workers > max_workers_t<SPLIT, BLOCK_N, EPI, PAIR>(num_sms)) return static_cast<int>(cudaErrorInvalidValue);
  }
  if (!PAIR) {
    if (!p.sk_on) return static_cast<int>(launch_pdl(kern, dim3(workers), dim3(NUM_THREADS), C::SMEM_BYTES, stream, p));
    kern<<<workers, NUM_THREADS, C::SMEM_BYTES, stream>>>(p);
    return static_cast<int>(cudaGetLastError());
  }
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[2];
  pair_launch_config<SPLIT, BLOCK_N, EPI, PAIR>(cfg, attr, 2 * workers, stream);
  if (pdl_enabled() && !p.sk_on) {
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.numAttrs = 2;
  }
  return static_cast<int>(cudaLaunchKernelEx(&cfg, kern, p));
}

}  // namespace

size_t convlstm_gates_smem_bytes(bool split, int block_n) {
  if (split) {
    if (block_n == 256) return Cfg<true, 256>::SMEM_BYTES;
    if (block_n == 128) return Cfg<true, 128>::SMEM_BYTES;
    return Cfg<true, 64>::SMEM_BYTES;
  }
  if (block_n == 256) return Cfg<false, 256>::SMEM_BYTES;
  if (block_n == 128) return Cfg<false, 128>::SMEM_BYTES;
  return Cfg<false, 64>::SMEM_BYTES;
}

int launch_convlstm_gates(const GateParams& p, bool split, int block_n, bool pair, int num_sms,
                          cudaStream_t stream) {
  if (pair) {  // CTA pairs: 256-pixel x 256-column tiles only
    if (block_n != 256 || p.n_munits <= 0) return static_cast<int>(cudaErrorInvalidValue);
    return split ? launch_t<true, 256, EPI_GATES, true>(p, num_sms, stream)
                 : launch_t<false, 256, EPI_GATES, true>(p, num_sms, stream);
  }
  if (split) {
    switch (block_n) {
      case 256: return launch_t<true, 256, EPI_GATES, false>(p, num_sms, stream);
      case 128: return launch_t<true, 128, EPI_GATES, false>(p, num_sms, stream);
      case 64: return launch_t<true, 64, EPI_GATES, false>(p, num_sms, stream);
    }
  } else {
    switch (block_n) {
      case 256: return launch_t<false, 256, EPI_GATES, false>(p, num_sms, stream);
      case 128: return launch_t<false, 128, EPI_GATES, false>(p, num_sms, stream);
      case 64: return launch_t<false, 64, EPI_GATES, false>(p, num_sms, stream);
    }
  }
  return static_cast<int>(cudaErrorInvalidValue);
}

int launch_convlstm_dgrad(const GateParams& p, bool split, bool pair, int num_sms, cudaStream_t stream) {
  if (pair) {
    if (p.n_munits <= 0) return static_cast<int>(cudaErrorInvalidValue);
    return split ? launch_t<true, 256, EPI_DGRAD, true>(p, num_sms, stream)
                 : launch_t<false, 256, EPI_DGRAD, true>(p, num_sms, stream);
  }
  return split ? launch_t<true, 256, EPI_DGRAD, false>(p, num_sms, stream)
               : launch_t<false, 256, EPI_DGRAD, false>(p, num_sms, stream);
}

// off by default: with it on, the B=1024 training step (GEMMs of the main stream next to the weight-gradient GEMMs of
// the side stream) was observed to hang in two of three runs on B200; +1.5 % when it runs (profiles/README.md)
namespace { bool g_pdl = false; }
bool pdl_enabled() { return g_pdl; }
void set_pdl(bool on) { g_pdl = on; }

int conv_gemm_max_workers(bool split, int block_n, bool dgrad, bool pair, int num_sms) {
  if (dgrad) {
    if (pair) return split ? max_workers_t<true, 256, EPI_DGRAD, true>(num_sms) : max_workers_t<false, 256, EPI_DGRAD, true>(num_sms);
    return split ? max_workers_t<true, 256, EPI_DGRAD, false>(num_sms) : max_workers_t<false, 256, EPI_DGRAD, false>(num_sms);
  }
  if (pair) return split ? max_workers_t<true, 256, EPI_GATES, true>(num_sms) : max_workers_t<false, 256, EPI_GATES, true>(num_sms);
  if (split) {
    switch (block_n) {
      case 256: return max_workers_t<true, 256, EPI_GATES, false>(num_sms);
      case 128: return max_workers_t<true, 128, EPI_GATES, false>(num_sms);
      case 64: return max_workers_t<true, 64, EPI_GATES, false>(num_sms);
    }
  } else {
    switch (block_n) {
      case 256: return max_workers_t<false, 256, EPI_GATES, false>(num_sms);
      case 128: return max_workers_t<false, 128, EPI_GATES, false>(num_sms);
      case 64: return max_workers_t<false, 64, EPI_GATES, false>(num_sms);
    }
  }
  return 0;
}

}  // namespace drasic
