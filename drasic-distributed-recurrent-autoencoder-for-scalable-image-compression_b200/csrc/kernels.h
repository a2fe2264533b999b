// Internal (non-ABI) declarations shared by the DRASIC sm_100a kernels and the C-ABI layer.
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <utility>

namespace drasic {

// Programmatic dependent launch (PDL): the kernels of the loop are launched so that each may start -- be scheduled,
// run its prologue -- while the tail of its predecessor still executes; every such kernel calls ptx::griddep_wait()
// before it touches global memory.  drasic_set_pdl(0) turns it off (plain stream order).
bool pdl_enabled();
void set_pdl(bool on);

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

constexpr int TILE_M = 128;     // UMMA M: pixels per output tile (= TMEM lanes)
constexpr int BLOCK_K = 64;     // fp16 elements per K block = one 128-byte swizzle row
constexpr int UMMA_K = 16;      // K per tcgen05.mma for 16-bit operands
constexpr int MAX_GROUPS = 64;  // sources (nodes) per launch
constexpr int ACT_SHIFT = 8;    // activations are stored as fp16(a * 2^8) (+ residual term)
constexpr float ACT_SCALE = 256.0f;
constexpr float ACT_INV_SCALE = 1.0f / 256.0f;

// physical channel order of an NHWC activation / state tensor
enum ChanOrder : int {
  ORDER_IDENT = 0,  // physical p == logical channel
  ORDER_QUAD = 1,   // physical p = q*(C/4) + c  <->  logical channel c*4 + q   (q = dy*2+dx)
};
// where the gate kernel writes h for its consumer (reference: PixelShuffleCell between the cells,
// modules/cell.py:147-165, functions/shuffle.py:4-25)
enum NextMode : int {
  NEXT_NONE = 0,
  NEXT_SAME = 1,  // same pixel, same physical channel
  NEXT_DOWN = 2,  // pixel_unshuffle(2): pixel (y/2,x/2), physical channel ((y&1)*2+(x&1))*Co + p
  NEXT_UP = 3,    // pixel_shuffle(2):  own order QUAD, p = q*(Co/4)+c -> pixel (2y+q/2, 2x+q%2), ch c
};
// where the dgrad kernel routes the input gradient (the inverse of the producer's NextMode)
enum DxMode : int {
  DX_SAME = 0,      // same pixel / channel (consumer is a pointwise kernel)
  DX_INV_DOWN = 1,  // input came through pixel_unshuffle: ch q*Cp+c at (y,x) -> producer (2y+q/2, 2x+q%2), ch c
  DX_INV_UP = 2,    // input came through pixel_shuffle:  ch c at (Y,X) -> producer (Y/2, X/2), ch q*Cin+c
};

// weight groups (nodes) packed by one launch: blockIdx.y = entry; `slot` = the group's plane in the packed arrays
struct PackBatch {
  const float* w_in[MAX_GROUPS];
  const float* w_h[MAX_GROUPS];
  int slot[MAX_GROUPS];
};

struct GateParams {
  CUtensorMap tm_x[2];  // layer input  NHWC fp16 {hi, lo}
  CUtensorMap tm_h[2];  // previous hidden state NHWC fp16 {hi, lo}
  CUtensorMap tm_w[2];  // packed weights, tile-contiguous: [groups][N tile][K block][tile rows][64] fp16 {hi, lo}, seen as rows of 64
  float* c_state;       // [B,H,W,Co] fp32 new cell state (in place when c_prev == nullptr)
  const float* c_prev;  // nullable: previous cell state in a separate buffer (training keeps every c_t)
  void* gates_out;      // nullable: post-activation gates [B,H,W,4,Co] (fp16, or fp32 when gates_f32), saved for backward
  int gates_f32;
  __half* h_out[2];     // new hidden state (own layout), {hi, lo}
  void* next[2];        // consumer copy of h: fp16 {hi, lo} or fp32 (next[0])
  const float* w_inv_scale;  // [n_groups] power of two: 1 / (ACT_SCALE * weight scale)
  int H, W, Cin, Co;
  int bw, bh, bn;  // TMA box (pixels): bw*bh*bn == 128
  int n_mtiles, n_ntiles;
  int tile_order;  // 0: N tile fastest in the persistent walk, 1: M unit fastest
  int has_hidden, has_cell;  // 0 on the first iteration (state is zero)
  int next_mode, next_f32;
  int precise;               // 1: IEEE exp/div/tanh (parity mode); 0: tanh.approx
  int n_groups;
  // n_groups > 1 (device arrays of n_groups ints): node g walks the M tiles [start[g], end[g]) -- neighbouring
  // ranges may share a tile -- as units (tiles, or pairs of tiles) numbered up to unit_end[g], and stores the
  // rows of the slots [img_end[g-1], img_end[g])
  const int* group_mtile_start;
  const int* group_mtile_end;
  const int* group_unit_end;
  const int* group_img_end;
  int n_munits;                // total work units of the M dimension
  // ---- position-major M tiles (n_groups == 1): the first pos_units units hold ONE pixel position of 128 images
  // each (TMA box 64 ch x 1 x 1 x 128 images), walked heaviest position first; the (tile, tap) K blocks whose tap
  // falls into the zero padding for that position are skipped altogether.  The images behind them (fewer than one
  // unit's worth) are covered by ordinary pixel-major tiles starting at M tile pos_tile_base.
  CUtensorMap tm_xp[2];
  CUtensorMap tm_hp[2];
  int pos_units, pos_upp;      // units in the position-major region; units (image tiles / pairs of them) per position
  int pos_tile_base;
  const int* pos_order;        // [H*W] pixel positions (y*W + x), most in-bounds taps first
  // positions with the same number of in-bounds taps form a class; inside a class the units walk the positions for one
  // block of images, then for the next (neighbouring positions of the same images share most of their halo in the L2).
  // Class c: units [pos_class_unit0[c], pos_class_unit0[c+1]), positions pos_order[pos_class_pos0[c] ...] (pos_class_n[c])
  int pos_classes;
  int pos_class_unit0[8], pos_class_pos0[8], pos_class_n[8], pos_class_taps[8];
  // pos_pair_mode (CTA pairs on fewer than 256 slots): a unit is TWO positions of one block of 128 images, one per
  // CTA of the pair -- pos_order then holds pairs (entries 2e, 2e+1) chosen with equal tap sets where possible; the
  // unit walks the union of the two sets (TMA zero fill covers a tap that is out of bounds for one of them)
  int pos_pair_mode;
  // ---- stream-K schedule (sk_on) of the LAST, partial wave: the first sk_static items are dealt whole, round
  // robin, like without it (the workers then run in lockstep over the same weight K blocks, which is what the L2
  // likes); the K blocks of the remaining items, laid end to end in item order, are cut into one equal range per
  // worker (CTA or CTA pair).  A worker whose range starts inside an item computes that part first and publishes
  // the fp32 partial accumulator (sk_scratch, sk_flags); the worker that computed the item's first K block adds the
  // partials of the workers behind it in its epilogue.  Items of class k (same K-block count: position-major units
  // with 9 / 6 / 4 ... in-bounds taps, then the pixel-major units) are [sk_item0[k], sk_item0[k+1]) and start at
  // K block sk_kb0[k] of the stream-K region.
  int sk_on, sk_classes;
  int sk_workers;             // CTAs (CTA pairs) of a stream-K launch: all co-resident
  int sk_static;              // items [0, sk_static) are dealt whole; a multiple of sk_workers
  int sk_item0[9], sk_kb0[9], sk_nkb[8];
  float* sk_scratch;           // [workers][CTAs per worker][BLOCK_N/16][128 rows][16] fp32
  unsigned int* sk_flags;      // [workers][CTAs per worker]: 1 while a published partial waits for its consumer
  // ---- dgrad epilogue (EPI_DGRAD): D[pixels, Nout] = conv3x3^T(dG) ; Cin here = 4*Co of the layer
  int w_group_tiles;           // N tiles per weight group (4Co / BLOCK_N for gates, ceil(Nout / 256) for dgrad)
  int w_nkb;                   // K blocks of the packed weight matrix (9 (Cin + Co) / 64 for gates, 36 Co / 64 for dgrad)
  int n_out;                   // dgrad: output columns = layer Cin + layer Co
  int n_x;                     // dgrad: columns [0,n_x) are dX, [n_x,n_out) are dH_{t-1}
  int dx_mode;                 // DxMode
  float* dx_out;               // dgrad: fp32 gradient w.r.t. the layer input, routed by dx_mode
  float* dhrec_out;            // dgrad: fp32 [B,H,W,n_out-n_x] gradient w.r.t. h_{t-1} (nullable at t=0)
  const float* a_inv_scale;    // dgrad: device scalar, 1 / scale of the fp16 dG planes
};

enum EpiKind : int { EPI_GATES = 0, EPI_DGRAD = 1 };

// launches the persistent tcgen05 gate kernel; returns a cudaError_t as int
int launch_convlstm_gates(const GateParams& p, bool split, int block_n, bool pair, int num_sms,
                          cudaStream_t stream);
int launch_convlstm_dgrad(const GateParams& p, bool split, bool pair, int num_sms, cudaStream_t stream);
// workers (CTAs, or CTA pairs) a stream-K launch of this configuration may use: all of them must be co-resident
int conv_gemm_max_workers(bool split, int block_n, bool dgrad, bool pair, int num_sms);
constexpr int SK_MAX_WORKERS = 148;                       // sizes of the stream-K scratch: one slot per CTA
constexpr size_t SK_SLOT_FLOATS = 128 * 256;              // one 128-row x 256-column fp32 accumulator
size_t convlstm_gates_smem_bytes(bool split, int block_n);

}  // namespace drasic
