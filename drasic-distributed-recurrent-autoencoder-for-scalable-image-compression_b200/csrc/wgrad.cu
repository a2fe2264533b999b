// Weight-gradient GEMM of a ConvLSTM cell for sm_100a (tcgen05, TMA, TMEM) -- the wgrad half of what
// autograd runs for `loss.backward()` (reference train_model.py:115) through modules/cell.py:117,132:
//
//   dW_in[gch, ci, tap] = sum_pixels dG[pix, gch] * X[pix + tap, ci]
//   dW_h [gch, ch, tap] = sum_pixels dG[pix, gch] * H_prev[pix + tap, ch]
//
// GEMM view: D[M = 128 gate channels, N = up to 256 (tap, channel) columns] += A^T B over K = pixels.
// Both operands are NHWC fp16 planes, i.e. MN-major with respect to K = pixels: a K block of 64
// pixels is loaded as TMA boxes of (64 channels x 64 pixels) -- 2 boxes of dG for A, up to 4 boxes of
// the tap-shifted activation plane for B (out-of-bounds pixels read as zero = conv padding) -- and fed
// to tcgen05.mma through MN-major SWIZZLE_128B descriptors.  The tensor maps are 5-D with the 64-channel
// chunk index as the OUTERMOST dimension, so one TMA instruction brings two stacked boxes (3 TMA
// instructions per K block; with one instruction per box the single producer thread could not keep up).  One launch covers one (layer, iteration);
// tiles are (node, m-tile, n-tile, k-split) and accumulate into the fp32 packed-order gradient with
// vector atomics, so iterations, k-splits and launches compose by addition.
// Roles / pipeline are those of the gate kernel: warp 0 TMA, warp 1 MMA, warps 2..5 epilogue.
#include "backward.h"
#include "ptx.cuh"

// Diagnostic switches (skip the TMA loads / the MMAs ...; results are then wrong by construction) exist only in
// -DDRASIC_DIAG builds; the shipped library compiles them out.
#ifdef DRASIC_DIAG
#define WG_DBG(p) ((p).dbg)
#else
#define WG_DBG(p) 0
#endif

namespace drasic {

namespace {

constexpr int WG_THREADS = 192;
constexpr int BOX_BYTES = 64 * 64 * 2;       // one (64 ch x 64 pixel) fp16 box = 8 KB
constexpr int WG_BLOCK_N = 256;
constexpr int WG_DATA_BUDGET = 196608;

// PAIR: two CTAs of a cluster (cta_group::2) own 256 gate channels x up to 256 columns: each stages its own
// 128 gate channels of dG and half of the activation columns (32 KB instead of 48 KB per K block and CTA).
// WIDE (pairs only): the pair owns 256 gate channels x up to 512 columns -- two N = 256 MMAs per K step that
// share the dG operand, the whole TMEM as one accumulator (no epilogue overlap; wgrad tiles run for
// hundreds of K blocks).  L2 -> SMEM bytes per unit of MMA work drop from 32 KB to 24 KB per CTA: the kernel
// is bound by that fill rate (~43 B/clk/SM, the L2 throughput cap), not by the tensor pipe.
template <bool SPLIT, bool PAIR = false, bool WIDE = false>
struct WCfg {
  static constexpr int UNITS = WIDE ? 8 : 4;                    // 64-column units per tile
  static constexpr int A_BYTES = 2 * BOX_BYTES;                 // 128 gate channels
  static constexpr int B_BYTES = (PAIR ? UNITS / 2 : UNITS) * BOX_BYTES;    // columns staged by one CTA
  static constexpr int STAGE_BYTES = (SPLIT ? 2 : 1) * (A_BYTES + B_BYTES);
  static constexpr int STAGES = WG_DATA_BUDGET / STAGE_BYTES;
  static constexpr uint32_t TMEM_COLS = 2 * WG_BLOCK_N;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
};

// MN-major SWIZZLE_128B operand: 64 contiguous MN elements (128 B) per K row, 8 K rows per 1024-byte
// swizzle atom, atoms stacked along K (SBO = 1024), next 64 MN elements LBO = 8192 bytes further.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(BOX_BYTES >> 4) << 16;  // LBO
  d |= static_cast<uint64_t>(1024u >> 4) << 32;      // SBO
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;               // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ uint32_t idesc_mn(int M, int N) {
  return (1u << 4) | (1u << 15) | (1u << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}

struct WTile {
  int g, mt, is_h, u0, nunits, kb_lo, kb_hi;
};

__device__ __forceinline__ WTile decode_tile(const WgradParams& p, int tile, int m_units, int upt) {
  const int nt_total = p.n_tiles_x + (p.has_h ? p.n_tiles_h : 0);
  WTile t;
  // order: N tile fastest, then M unit, then K split, then node -- the tiles that run concurrently on the
  // persistent grid share their K range, so its dG / activation pixels come from DRAM once and are then
  // served from L2 to the other (M, N) tiles (with the K split fastest every wave re-read the whole tensor)
  int r = tile;
  const int nt = r % nt_total;
  r /= nt_total;
  t.mt = r % m_units;      // M tile (single CTA) or pair of M tiles (CTA pair)
  r /= m_units;
  const int ks = r % p.k_splits;
  t.g = r / p.k_splits;
  t.is_h = nt >= p.n_tiles_x;
  const int ntl = t.is_h ? nt - p.n_tiles_x : nt;
  const int units = 9 * ((t.is_h ? p.Co : p.Cin) / 64);
  t.u0 = ntl * upt;
  t.nunits = min(upt, units - t.u0);
  const int g_lo = (t.g == 0 || !p.group_kb_end) ? 0 : p.group_kb_end[t.g - 1];
  const int g_hi = p.group_kb_end ? p.group_kb_end[t.g] : p.n_kb;
  const int len = g_hi - g_lo;
  const int per = (len + p.k_splits - 1) / p.k_splits;
  t.kb_lo = min(g_hi, g_lo + ks * per);
  t.kb_hi = min(g_hi, t.kb_lo + per);
  return t;
}

template <bool SPLIT, bool PAIR, bool WIDE>
__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_kernel(const __grid_constant__ WgradParams p) {
  using C = WCfg<SPLIT, PAIR, WIDE>;
  static_assert(!WIDE || PAIR, "wide tiles are a CTA-pair mode");
  const uint32_t cta_rank = PAIR ? ptx::cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + C::STAGES;
  uint64_t* tmem_full = bars + 2 * C::STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
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
      ptx::mbar_init(&tmem_empty[a], PAIR ? 8 : 4);
    }
    ptx::fence_mbar_init();
    ptx::prefetch_tmap(&p.tm_g[0]);
    ptx::prefetch_tmap(&p.tm_x[0]);
    if (p.has_h) ptx::prefetch_tmap(&p.tm_h[0]);
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

  const int nt_total = p.n_tiles_x + (p.has_h ? p.n_tiles_h : 0);
  const int m_units = PAIR ? p.m_tiles >> 1 : p.m_tiles;
  const int total_tiles = p.n_groups * m_units * nt_total * p.k_splits;
  const int unit_stride = PAIR ? gridDim.x >> 1 : gridDim.x;
  const int unit_first = PAIR ? blockIdx.x >> 1 : blockIdx.x;
  const int HW = p.H * p.W;
  const int kb_per_img = HW >= 64 ? HW / 64 : 1;   // HW < 64: one K block spans kbn images
  const int kbx = p.W / p.kbw;

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      // the leader's full barriers (shared::cluster addresses; the CTA windows map linearly)
      const uint32_t full_remote0 = PAIR ? ptx::mapa_u32(&full_bar[0], 0) : 0u;
      for (int tile = unit_first; tile < total_tiles; tile += unit_stride) {
        const WTile t = decode_tile(p, tile, m_units, C::UNITS);
        const CUtensorMap* tb = t.is_h ? p.tm_h : p.tm_x;
        const int cpu = (t.is_h ? p.Co : p.Cin) / 64;
        // a pair: this CTA stages gate channels of M tile 2*mt + rank and columns [rank*nunits/2, ...) --
        // always one 2-box TMA (with nunits == 2 its second box is unused filler)
        const int mt = PAIR ? 2 * t.mt + static_cast<int>(cta_rank) : t.mt;
        // pairs: per MMA group of up to 4 units this CTA stages half of the group's units, always as one 2-box
        // TMA (with a 2-unit group its second box is unused filler)
        const int n_grp = WIDE ? (t.nunits + 3) >> 2 : 1;
        const int my_units = PAIR ? 2 * n_grp : t.nunits;
        const uint32_t bytes = (SPLIT ? 2 : 1) * (C::A_BYTES + my_units * BOX_BYTES);
        // per-tile constants of the B loads: (tap offset, chunk) of each 2-box TMA
        int u_dx[C::UNITS / 2], u_dy[C::UNITS / 2], u_chunk[C::UNITS / 2];
#pragma unroll
        for (int uu = 0; uu < C::UNITS / 2; ++uu) {
          const int u = 2 * uu;
          int unit = t.u0 + u;
          if (PAIR) {
            const int grp = WIDE ? u >> 1 : 0;
            const int cnt = WIDE ? min(4, t.nunits - 4 * grp) : t.nunits;
            unit = t.u0 + 4 * grp + static_cast<int>(cta_rank) * (cnt >> 1);
          }
          int tap = unit / cpu;
          u_chunk[uu] = unit - tap * cpu;
          if (WG_DBG(p) == 5) tap = 4;
          u_dx[uu] = tap % 3 - 1;
          u_dy[uu] = tap / 3 - 1;
        }
        // pixel box of the first K block; later blocks advance incrementally (no divisions in the loop)
        int n0, y0, x0;
        if (HW >= 64) {
          n0 = t.kb_lo / kb_per_img;
          const int rem = t.kb_lo - n0 * kb_per_img;
          y0 = (rem / kbx) * p.kbh;
          x0 = (rem % kbx) * p.kbw;
        } else {
          n0 = t.kb_lo * p.kbn; y0 = 0; x0 = 0;
        }
        for (int kb = t.kb_lo; kb < t.kb_hi; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          if (WG_DBG(p) == 1) {
            ptx::mbar_arrive(&full_bar[stage]);
            if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
            continue;
          }
          uint32_t xbytes = bytes;   // diagnostics: 3 = no A loads, 4 = no B loads, 5 = centre tap only
          if (WG_DBG(p) == 3) xbytes = (SPLIT ? 2 : 1) * (my_units * BOX_BYTES);
          if (WG_DBG(p) == 4) xbytes = (SPLIT ? 2 : 1) * C::A_BYTES;
          if (!PAIR) ptx::mbar_expect_tx(&full_bar[stage], xbytes);
          else if (leader) ptx::mbar_expect_tx(&full_bar[stage], 2 * xbytes);
          uint8_t* st = smem + stage * C::STAGE_BYTES;
          for (int pl = 0; pl < (SPLIT ? 2 : 1); ++pl) {
            uint8_t* sa = st + pl * C::A_BYTES;
            uint8_t* sb = st + (SPLIT ? 2 : 1) * C::A_BYTES + pl * C::B_BYTES;
            if (WG_DBG(p) == 3) {
            } else if (PAIR) ptx::tma_load_5d_pair(sa, &p.tm_g[pl], (full_remote0 + 8u * stage), 0, x0, y0, n0, mt * 2);
            else ptx::tma_load_5d(sa, &p.tm_g[pl], &full_bar[stage], 0, x0, y0, n0, mt * 2);
            if (WG_DBG(p) == 4) continue;
#pragma unroll
            for (int uu = 0; uu < C::UNITS / 2; ++uu) {   // a pair of 64-channel chunks of the same tap
              if (2 * uu >= my_units) break;
              if (PAIR)
                ptx::tma_load_5d_pair(sb + 2 * uu * BOX_BYTES, &tb[pl], (full_remote0 + 8u * stage), 0, x0 + u_dx[uu],
                                      y0 + u_dy[uu], n0, u_chunk[uu]);
              else
                ptx::tma_load_5d(sb + 2 * uu * BOX_BYTES, &tb[pl], &full_bar[stage], 0, x0 + u_dx[uu],
                                 y0 + u_dy[uu], n0, u_chunk[uu]);
            }
          }
          // next K block's pixel box
          if (HW >= 64) {
            x0 += p.kbw;
            if (x0 == p.W) {
              x0 = 0;
              y0 += p.kbh;
              if (y0 == p.H) { y0 = 0; ++n0; }
            }
          } else {
            n0 += p.kbn;
          }
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
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
      const uint64_t desc_a_hi = umma_desc_mn_sw128(smem0);
      const uint64_t desc_a_lo = umma_desc_mn_sw128(smem0 + C::A_BYTES);
      const uint64_t desc_b_hi = umma_desc_mn_sw128(smem0 + (SPLIT ? 2 : 1) * C::A_BYTES);
      const uint64_t desc_b_lo = umma_desc_mn_sw128(smem0 + (SPLIT ? 2 : 1) * C::A_BYTES + C::B_BYTES);
      for (int tile = unit_first; tile < total_tiles; tile += unit_stride) {
        const WTile t = decode_tile(p, tile, m_units, C::UNITS);
        if (t.kb_hi <= t.kb_lo) continue;  // empty node / empty k-split: nothing to add
        // MMA groups of up to 256 columns sharing the A operand (two only with WIDE tiles)
        const int cnt0 = WIDE ? min(4, t.nunits) : t.nunits;
        const int cnt1 = WIDE ? t.nunits - cnt0 : 0;
        const uint32_t idesc0 = idesc_mn(PAIR ? 256 : 128, cnt0 * 64);
        const uint32_t idesc1 = idesc_mn(PAIR ? 256 : 128, cnt1 * 64);
        const int acc = WIDE ? 0 : it & 1;
        const uint32_t acc_phase = WIDE ? (it & 1) : (it >> 1) & 1;
        ++it;
        ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_base = tmem_base + acc * WG_BLOCK_N;
        for (int kb = t.kb_lo; kb < t.kb_hi; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          // descriptors differ from the stage-0 / k-0 ones only in the 14-bit start-address field
          const uint32_t st_off = static_cast<uint32_t>(stage) * (C::STAGE_BYTES >> 4);
          if (WG_DBG(p) < 2)
#pragma unroll
          for (int grp = 0; grp < (WIDE ? 2 : 1); ++grp) {
            if (grp == 1 && cnt1 == 0) break;
            const uint32_t idesc = grp == 0 ? idesc0 : idesc1;
            const uint32_t d_tmem = d_base + grp * WG_BLOCK_N;
            const uint32_t g_off = st_off + grp * ((2 * BOX_BYTES) >> 4);  // this group's boxes in the CTA's B region
#pragma unroll
            for (int k = 0; k < 4; ++k) {            // 4 x UMMA_K(16 pixels) per 64-pixel K block
              const uint32_t k_off = k * ((16 * 128) >> 4);    // 16 K rows of 128 bytes
              const uint64_t da_hi = desc_a_hi + (st_off + k_off);
              const uint64_t db_hi = desc_b_hi + (g_off + k_off);
              const uint32_t first = (kb != t.kb_lo || k != 0) ? 1u : 0u;
              if (SPLIT) {
                const uint64_t da_lo = desc_a_lo + (st_off + k_off);
                const uint64_t db_lo = desc_b_lo + (g_off + k_off);
                if (PAIR) {
                  ptx::umma_f16_pair(d_tmem, da_lo, db_hi, idesc, first);
                  ptx::umma_f16_pair(d_tmem, da_hi, db_lo, idesc, 1);
                  ptx::umma_f16_pair(d_tmem, da_hi, db_hi, idesc, 1);
                } else {
                  ptx::umma_f16(d_tmem, da_lo, db_hi, idesc, first);
                  ptx::umma_f16(d_tmem, da_hi, db_lo, idesc, 1);
                  ptx::umma_f16(d_tmem, da_hi, db_hi, idesc, 1);
                }
              } else if (PAIR) {
                ptx::umma_f16_pair(d_tmem, da_hi, db_hi, idesc, first);
              } else {
                ptx::umma_f16(d_tmem, da_hi, db_hi, idesc, first);
              }
            }
          }
          if (PAIR) ptx::umma_commit_pair(&empty_bar[stage]); else ptx::umma_commit(&empty_bar[stage]);
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        if (PAIR) ptx::umma_commit_pair(&tmem_full[acc]); else ptx::umma_commit(&tmem_full[acc]);
      }
    }
  } else {
    // ===================================================== epilogue: fp32 vector atomics
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int Kw = 9 * (p.Cin + p.Co);
    const float scale = __ldg(p.g_inv_scale) * ACT_INV_SCALE;
    uint32_t tmem_empty_leader[2];
    if (PAIR)
      for (int a = 0; a < 2; ++a) tmem_empty_leader[a] = ptx::mapa_u32(&tmem_empty[a], 0);
    int it = 0;
    for (int tile = unit_first; tile < total_tiles; tile += unit_stride) {
      const WTile t = decode_tile(p, tile, m_units, C::UNITS);
      if (t.kb_hi <= t.kb_lo) continue;
      const int acc = WIDE ? 0 : it & 1;
      const uint32_t acc_phase = WIDE ? (it & 1) : (it >> 1) & 1;
      ++it;
      ptx::mbar_wait(&tmem_full[acc], acc_phase);
      ptx::tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * WG_BLOCK_N;
      const int mt = PAIR ? 2 * t.mt + static_cast<int>(cta_rank) : t.mt;
      float* dst = p.gw + (static_cast<size_t>(t.g) * 4 * p.Co + mt * 128 + row) * Kw +
                   (t.is_h ? 9 * p.Cin : 0) + t.u0 * 64;
#pragma unroll 1
      for (int j = 0; j < t.nunits * 4; ++j) {
        float v[16];
        ptx::tmem_ld16(t_row + j * 16, v);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int qd = 0; qd < 4; ++qd)
          atomicAdd(reinterpret_cast<float4*>(dst + j * 16) + qd,
                    make_float4(v[4 * qd] * scale, v[4 * qd + 1] * scale, v[4 * qd + 2] * scale, v[4 * qd + 3] * scale));
      }
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

template <bool SPLIT, bool PAIR, bool WIDE>
int launch_w(const WgradParams& p, int num_sms, cudaStream_t stream) {
  using C = WCfg<SPLIT, PAIR, WIDE>;
  auto kern = wgrad_kernel<SPLIT, PAIR, WIDE>;
  static bool configured[64] = {};  // per instantiation and device (the attribute is per device)
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return static_cast<int>(cudaErrorInvalidDevice);
  if (!configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) return static_cast<int>(e);
    configured[dev] = true;
  }
  const int nt_total = p.n_tiles_x + (p.has_h ? p.n_tiles_h : 0);
  const int total = p.n_groups * (PAIR ? p.m_tiles / 2 : p.m_tiles) * nt_total * p.k_splits;
  if (!PAIR) {
    const int grid = total < num_sms ? total : num_sms;
    kern<<<grid, WG_THREADS, C::SMEM_BYTES, stream>>>(p);
    return static_cast<int>(cudaGetLastError());
  }
  const int pairs = num_sms / 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * (total < pairs ? total : pairs));
  cfg.blockDim = dim3(WG_THREADS);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return static_cast<int>(cudaLaunchKernelEx(&cfg, kern, p));
}

}  // namespace

int launch_wgrad(const WgradParams& p, bool split, bool pair, int num_sms, cudaStream_t stream) {
  if (pair) {
    if (p.m_tiles & 1) return static_cast<int>(cudaErrorInvalidValue);
    if (p.wide) return split ? launch_w<true, true, true>(p, num_sms, stream) : launch_w<false, true, true>(p, num_sms, stream);
    return split ? launch_w<true, true, false>(p, num_sms, stream) : launch_w<false, true, false>(p, num_sms, stream);
  }
  if (p.wide) return static_cast<int>(cudaErrorInvalidValue);
  return split ? launch_w<true, false, false>(p, num_sms, stream) : launch_w<false, false, false>(p, num_sms, stream);
}

}  // namespace drasic
