// C ABI (include/drasic_b200.h) over the sm_100a kernels: argument validation, TMA descriptor
// construction, launches.  No torch types, no allocation, no synchronisation.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unordered_map>

#include "../../include/drasic_b200.h"
#include "kernels.h"
#include "pointwise.h"
#include "backward.h"
#include "optim.h"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int cuda_fail(int e, const char* what) {
  if (e == 0) return 0;
  return fail(e, "%s: %s", what, cudaGetErrorString(static_cast<cudaError_t>(e)));
}

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                   const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// Encoded tensor maps are cached per (kind, base pointer, shape, box): a training step launches ~300 GEMMs that name
// the same few dozen planes iteration after iteration, and cuTensorMapEncodeTiled is the most expensive thing the host
// does per launch.  A map encodes only the pointer and the geometry, never contents, so a cached map stays valid for as
// long as the same pointer is used with the same shape (a re-allocated buffer at the same address gets the same map).
struct MapKey {
  const void* base;
  long long a, b, c, d;
  int kind, e, f, g, h;
  bool operator==(const MapKey& o) const {
    return base == o.base && a == o.a && b == o.b && c == o.c && d == o.d && kind == o.kind && e == o.e && f == o.f &&
           g == o.g && h == o.h;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t x = reinterpret_cast<size_t>(k.base) * 0x9E3779B97F4A7C15ull;
    const long long v[9] = {k.a, k.b, k.c, k.d, k.kind, k.e, k.f, k.g, k.h};
    for (long long t : v) x = (x ^ static_cast<size_t>(t)) * 0x100000001B3ull;
    return x;
  }
};
thread_local std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;
bool cached_map(const MapKey& k, CUtensorMap* m) {
  auto it = g_maps.find(k);
  if (it == g_maps.end()) return false;
  *m = it->second;
  return true;
}
void remember_map(const MapKey& k, const CUtensorMap& m) {
  if (g_maps.size() > 4096) g_maps.clear();      // (arenas were re-allocated many times: start over)
  g_maps.emplace(k, m);
}

// NHWC fp16 activation plane [N,H,W,C]; box = (64 channels, bw, bh, bn) pixels, 128B swizzle,
// out-of-bounds elements (the conv halo) read as zero.
int make_act_map(CUtensorMap* m, const void* base, int N, int H, int W, int C, int bw, int bh, int bn) {
  const MapKey key{base, N, H, W, C, 0, bw, bh, bn, 0};
  if (cached_map(key, m)) return 0;
  EncodeTiledFn enc = encode_fn();
  if (!enc) return fail(DRASIC_ERR_NO_DEVICE, "cuTensorMapEncodeTiled entry point unavailable");
  const cuuint64_t dims[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(W),
                              static_cast<cuuint64_t>(H), static_cast<cuuint64_t>(N)};
  const cuuint64_t strides[3] = {static_cast<cuuint64_t>(C) * 2, static_cast<cuuint64_t>(W) * C * 2,
                                 static_cast<cuuint64_t>(H) * W * C * 2};
  const cuuint32_t box[4] = {drasic::BLOCK_K, static_cast<cuuint32_t>(bw), static_cast<cuuint32_t>(bh),
                             static_cast<cuuint32_t>(bn)};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(DRASIC_ERR_TENSORMAP, "activation tensor map rejected (CUresult %d) N=%d H=%d W=%d C=%d box=%dx%dx%d",
                static_cast<int>(r), N, H, W, C, bw, bh, bn);
  remember_map(key, *m);
  return 0;
}

// the same plane seen as [C/64 chunks][N,H,W][64 ch]: the chunk index is the outermost dimension so one
// box = `nchunk` stacked (64 ch x pixel-box) tiles, each in the MN-major SWIZZLE_128B layout (wgrad.cu)
int make_act_map_chunked(CUtensorMap* m, const void* base, int N, int H, int W, int C, int bw, int bh, int bn,
                         int nchunk) {
  const MapKey key{base, N, H, W, C, 1, bw, bh, bn, nchunk};
  if (cached_map(key, m)) return 0;
  EncodeTiledFn enc = encode_fn();
  if (!enc) return fail(DRASIC_ERR_NO_DEVICE, "cuTensorMapEncodeTiled entry point unavailable");
  const cuuint64_t dims[5] = {64, static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H),
                              static_cast<cuuint64_t>(N), static_cast<cuuint64_t>(C / 64)};
  const cuuint64_t strides[4] = {static_cast<cuuint64_t>(C) * 2, static_cast<cuuint64_t>(W) * C * 2,
                                 static_cast<cuuint64_t>(H) * W * C * 2, 128};
  const cuuint32_t box[5] = {64, static_cast<cuuint32_t>(bw), static_cast<cuuint32_t>(bh),
                             static_cast<cuuint32_t>(bn), static_cast<cuuint32_t>(nchunk)};
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<void*>(base), dims, strides, box,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(DRASIC_ERR_TENSORMAP, "chunked activation tensor map rejected (CUresult %d) N=%d H=%d W=%d C=%d box=%dx%dx%dx%d",
                static_cast<int>(r), N, H, W, C, bw, bh, bn, nchunk);
  remember_map(key, *m);
  return 0;
}

// packed weights, tile-contiguous ([N tile][K block][tile rows][64] fp16): `rows` rows of 64 elements, 128 bytes apart;
// box = (64, box_rows) = one contiguous run of box_rows * 128 bytes
int make_weight_map(CUtensorMap* m, const void* base, long long rows, int box_rows) {
  const MapKey key{base, rows, 0, 0, 0, 2, box_rows, 0, 0, 0};
  if (cached_map(key, m)) return 0;
  EncodeTiledFn enc = encode_fn();
  if (!enc) return fail(DRASIC_ERR_NO_DEVICE, "cuTensorMapEncodeTiled entry point unavailable");
  const cuuint64_t dims[2] = {drasic::BLOCK_K, static_cast<cuuint64_t>(rows)};
  const cuuint64_t strides[1] = {drasic::BLOCK_K * 2};
  const cuuint32_t box[2] = {drasic::BLOCK_K, static_cast<cuuint32_t>(box_rows)};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(DRASIC_ERR_TENSORMAP, "weight tensor map rejected (CUresult %d) rows=%lld box_rows=%d", static_cast<int>(r), rows, box_rows);
  remember_map(key, *m);
  return 0;
}

int g_num_sms = 0;
int num_sms() {
  if (g_num_sms > 0) return g_num_sms;
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
  g_num_sms = n;
  return n;
}

}  // namespace

extern "C" {

int drasic_version(void) { return DRASIC_ABI_VERSION; }
const char* drasic_last_error(void) { return g_err; }
int drasic_device_sms(void) {
  const int n = num_sms();
  if (n <= 0) return fail(DRASIC_ERR_NO_DEVICE, "no CUDA device");
  return n;
}

int drasic_set_pdl(int enable) {
  const int was = drasic::pdl_enabled() ? 1 : 0;
  drasic::set_pdl(enable != 0);
  return was;
}

size_t drasic_stream_k_scratch_bytes(void) { return drasic::SK_MAX_WORKERS * drasic::SK_SLOT_FLOATS * sizeof(float); }
size_t drasic_stream_k_flag_bytes(void) { return drasic::SK_MAX_WORKERS * sizeof(unsigned int); }

size_t drasic_packed_weight_bytes(int Cin, int Co) {
  return static_cast<size_t>(4) * Co * 9 * (Cin + Co) * 2;
}
size_t drasic_packed_dgrad_weight_bytes(int Cin, int Co) {
  return static_cast<size_t>((Cin + Co + 255) / 256 * 256) * 36 * Co * 2;
}

static int fill_pack_batch(drasic::PackBatch& b, const char* who, const float* const* w_in, const float* const* w_h,
                           const int* slots, int n, int n_slots) {
  if (!w_in || !w_h || !slots || n < 1 || n > DRASIC_MAX_GROUPS || n_slots < 1)
    return fail(DRASIC_ERR_INVALID, "%s: n=%d weight groups (1..%d)", who, n, DRASIC_MAX_GROUPS);
  for (int i = 0; i < n; ++i) {
    if (!w_in[i] || !w_h[i] || slots[i] < 0 || slots[i] >= n_slots) return fail(DRASIC_ERR_INVALID, "%s: group %d: null weights or slot out of range", who, i);
    b.w_in[i] = w_in[i]; b.w_h[i] = w_h[i]; b.slot[i] = slots[i];
  }
  return 0;
}

int drasic_pack_lstm_weights_batch(const float* const* w_in, const float* const* w_h, const int* slots, int n, int n_slots,
                                   int Cin, int Co, int in_order, int out_order, int block_n, void* out_hi, void* out_lo,
                                   float* inv_scale_out, void* scratch, void* stream) {
  if (!out_hi || !out_lo || !inv_scale_out || !scratch) return fail(DRASIC_ERR_INVALID, "pack_lstm_weights: null pointer");
  if (Cin % 64 || Co % 64 || (block_n != 256 && block_n != 128 && block_n != 64) || (4 * Co) % block_n)
    return fail(DRASIC_ERR_INVALID, "pack_lstm_weights: Cin=%d Co=%d block_n=%d unsupported", Cin, Co, block_n);
  if ((in_order == DRASIC_ORDER_QUAD && Cin % 4) || (out_order == DRASIC_ORDER_QUAD && (Co / 4) % 16))
    return fail(DRASIC_ERR_INVALID, "pack_lstm_weights: quad order needs Co/4 %% 16 == 0");
  drasic::PackBatch b;
  if (int rc = fill_pack_batch(b, "pack_lstm_weights", w_in, w_h, slots, n, n_slots)) return rc;
  return cuda_fail(drasic::launch_pack_lstm_weights(b, n, n_slots, Cin, Co, in_order, out_order, block_n,
                                                    static_cast<__half*>(out_hi), static_cast<__half*>(out_lo), inv_scale_out,
                                                    static_cast<unsigned int*>(scratch), static_cast<cudaStream_t>(stream)),
                   "pack_lstm_weights launch");
}

int drasic_pack_lstm_weights(const float* w_in, const float* w_h, int Cin, int Co, int in_order,
                             int out_order, int block_n, void* out_hi, void* out_lo,
                             float* inv_scale_out, void* scratch4, void* stream) {
  if (!w_in || !w_h) return fail(DRASIC_ERR_INVALID, "pack_lstm_weights: null pointer");
  const int slot = 0;
  return drasic_pack_lstm_weights_batch(&w_in, &w_h, &slot, 1, 1, Cin, Co, in_order, out_order, block_n, out_hi, out_lo,
                                        inv_scale_out, scratch4, stream);
}

// pixel box of one GEMM tile: bw*bh*bn == npix pixels.  bw is the largest power of two (<= npix) that
// divides W, so any even-sized plane tiles (768x512 images give 384/192/96-wide planes -> 128/64/32-wide
// boxes); boxes span several images only when one image is smaller than a tile.
static int pixel_box(int H, int W, int npix, int* bw, int* bh, int* bn) {
  if (H <= 0 || W <= 0) return -1;
  int w = npix;
  while (w > 1 && W % w) w >>= 1;
  *bw = w;
  *bh = H < npix / w ? H : npix / w;
  if ((npix / w) % *bh || H % *bh) return -1;
  *bn = npix / (w * *bh);
  if (*bn > 1 && (w != W || *bh != H)) return -1;
  return 0;
}

// positions of an H x W plane by number of in-bounds 3x3 taps: count[t], t = 1..9
static void tap_class_counts(int H, int W, int* count) {
  for (int i = 0; i < 10; ++i) count[i] = 0;
  int cy[4] = {0, 0, 0, 0}, cx[4] = {0, 0, 0, 0};     // rows / columns with 1, 2, 3 in-bounds taps along the axis
  if (H == 1) cy[1] = 1; else { cy[2] = 2; cy[3] = H - 2; }
  if (W == 1) cx[1] = 1; else { cx[2] = 2; cx[3] = W - 2; }
  for (int ty = 1; ty <= 3; ++ty)
    for (int tx = 1; tx <= 3; ++tx) count[ty * tx] += cy[ty] * cx[tx];
}

// position-major region of a conv GEMM launch (see drasic_convlstm_args): validates, builds the (64 ch, 1, 1, 128
// images) tensor maps of the planes in `hi` / `lo` and fills the unit bookkeeping
static int setup_pos_tiles(drasic::GateParams& p, const char* who, const int* pos_order, int pos_imgs, int pos_pairs,
                           int n_groups, bool pair, int B, int H, int W, int n_units) {
  p.pos_units = 0; p.pos_upp = 1; p.pos_tile_base = 0; p.pos_order = nullptr; p.pos_classes = 0; p.pos_pair_mode = 0;
  if (pos_imgs == 0) return 0;
  if (pos_pairs && (!pair || H < 4 || W < 4 || ((H | W) & 1)))
    return fail(DRASIC_ERR_INVALID, "%s: pos_pairs needs cta_pair and an even plane of at least 4x4 (H=%d W=%d)", who, H, W);
  const int unit_imgs = (pair && !pos_pairs) ? 256 : 128;
  if (pos_imgs < 0 || pos_imgs > B || pos_imgs % unit_imgs || n_groups != 1 || !pos_order)
    return fail(DRASIC_ERR_INVALID, "%s: pos_imgs=%d needs one group, a pos_order table and a multiple of %d slots <= B_pad", who, pos_imgs, unit_imgs);
  p.pos_upp = pos_imgs / unit_imgs;            // blocks of images, each walked by every table entry
  p.pos_pair_mode = pos_pairs ? 1 : 0;
  int c = 0, e0 = 0;
  if (pos_pairs) {
    // pairs of positions with equal tap sets: interior (9 taps), then edges and the two corner pairs (6 taps)
    const int n9 = (H - 2) * (W - 2) / 2, n6 = (H - 2) + (W - 2) + 2;
    const int n[2] = {n9, n6}, taps[2] = {9, 6};
    for (int k = 0; k < 2; ++k) {
      if (!n[k]) continue;
      p.pos_class_unit0[c] = e0 * p.pos_upp; p.pos_class_pos0[c] = e0; p.pos_class_n[c] = n[k]; p.pos_class_taps[c] = taps[k];
      e0 += n[k];
      ++c;
    }
  } else {
    // classes of positions by in-bounds tap count, in the order of pos_order (9, 6, 4, 3, 2, 1 taps)
    int count[10];
    tap_class_counts(H, W, count);
    for (int taps = 9; taps >= 1; --taps) {
      if (!count[taps]) continue;
      p.pos_class_unit0[c] = e0 * p.pos_upp; p.pos_class_pos0[c] = e0; p.pos_class_n[c] = count[taps]; p.pos_class_taps[c] = taps;
      e0 += count[taps];
      ++c;
    }
  }
  p.pos_classes = c;
  p.pos_units = e0 * p.pos_upp;
  p.pos_tile_base = static_cast<int>(static_cast<long long>(pos_imgs) * H * W / 128);
  p.pos_order = pos_order;
  const int rest = p.n_mtiles - p.pos_tile_base;
  if (n_units != p.pos_units + (pair ? (rest + 1) / 2 : rest))
    return fail(DRASIC_ERR_INVALID, "%s: n_units=%d does not match pos_imgs=%d", who, n_units, pos_imgs);
  return 0;
}

// stream-K schedule of a conv GEMM launch (see GateParams): the classes of work items with equal K-block counts, in
// the order the kernel numbers the items (position-major units sorted by in-bounds taps, then pixel-major units)
static int setup_stream_k(drasic::GateParams& p, const char* who, int stream_k, void* scratch, void* flags, int chunks_per_tap,
                          int max_workers) {
  p.sk_on = 0;
  if (!stream_k) return 0;
  if (!scratch || !flags) return fail(DRASIC_ERR_INVALID, "%s: stream_k needs sk_scratch and sk_flags", who);
  if (max_workers < 1) return fail(DRASIC_ERR_NO_DEVICE, "%s: stream_k: no co-resident workers on this device", who);
  // classes of the whole launch: (first item, items, K blocks per item)
  int c_item[9], c_n[9], c_nkb[9], nc = 0, item = 0;
  for (int c = 0; c < p.pos_classes; ++c) {
    c_item[nc] = item; c_n[nc] = p.pos_class_n[c] * p.pos_upp * p.n_ntiles; c_nkb[nc] = p.pos_class_taps[c] * chunks_per_tap;
    item += c_n[nc++];
  }
  const int rest = (p.n_munits - p.pos_units) * p.n_ntiles;
  if (rest > 0) { c_item[nc] = item; c_n[nc] = rest; c_nkb[nc] = 9 * chunks_per_tap; item += rest; ++nc; }
  if (nc == 0 || item != p.n_munits * p.n_ntiles) return fail(DRASIC_ERR_INVALID, "%s: stream_k: inconsistent work item count", who);
  // whole waves are dealt statically; the stream-K region is the partial last wave
  const int total_items = item;
  int workers = max_workers;
  long long kb_all = 0;
  for (int c = 0; c < nc; ++c) kb_all += static_cast<long long>(c_n[c]) * c_nkb[c];
  if (total_items < workers) {                      // fewer items than workers: at least four K blocks per worker
    const long long w = kb_all / 4 < 1 ? 1 : kb_all / 4;
    if (w < workers) workers = static_cast<int>(w);
  }
  int n_static = total_items / workers * workers;
  if (n_static == total_items) return 0;            // whole waves only: nothing to balance
  int k = 0;
  long long kb = 0;
  for (int attempt = 0; attempt < 2; ++attempt) {
    k = 0; kb = 0;
    for (int c = 0; c < nc; ++c) {
      const int lo = c_item[c] > n_static ? c_item[c] : n_static, hi = c_item[c] + c_n[c];
      if (lo >= hi) continue;
      p.sk_item0[k] = lo; p.sk_kb0[k] = static_cast<int>(kb); p.sk_nkb[k] = c_nkb[c];
      kb += static_cast<long long>(hi - lo) * c_nkb[c];
      ++k;
    }
    // every worker needs a non-empty range (a worker waits for the partials of ALL workers behind it inside its
    // item): a remainder of very few K blocks takes one whole wave along
    if (kb >= 4LL * workers || n_static < workers) break;
    n_static -= workers;
  }
  if (k == 0 || kb < workers || kb > 0x3fffffffLL) return fail(DRASIC_ERR_INVALID, "%s: stream_k: inconsistent K-block count", who);
  p.sk_item0[k] = total_items; p.sk_kb0[k] = static_cast<int>(kb);
  p.sk_classes = k;
  p.sk_on = 1;
  p.sk_static = n_static;
  p.sk_workers = workers;
  p.sk_scratch = static_cast<float*>(scratch);
  p.sk_flags = static_cast<unsigned int*>(flags);
  p.tile_order = 0;
  return 0;
}

int drasic_convlstm_fwd(const drasic_convlstm_args* a, void* stream) {
  if (!a) return fail(DRASIC_ERR_INVALID, "convlstm_fwd: null args");
  const int H = a->H, W = a->W, Cin = a->Cin, Co = a->Co, B = a->B_pad;
  if (Cin % 64 || Co % 64 || Cin <= 0 || Co <= 0)
    return fail(DRASIC_ERR_INVALID, "convlstm_fwd: channels must be multiples of 64 (Cin=%d Co=%d)", Cin, Co);
  if (a->block_n != 256 && a->block_n != 128 && a->block_n != 64)
    return fail(DRASIC_ERR_INVALID, "convlstm_fwd: block_n=%d", a->block_n);
  if ((4 * Co) % a->block_n) return fail(DRASIC_ERR_INVALID, "convlstm_fwd: 4*Co %% block_n != 0");
  // pixel box of one M tile
  int bw, bh, bn;
  if (pixel_box(H, W, 128, &bw, &bh, &bn)) return fail(DRASIC_ERR_INVALID, "convlstm_fwd: H=%d W=%d unsupported", H, W);
  if (B <= 0 || B % bn) return fail(DRASIC_ERR_INVALID, "convlstm_fwd: B_pad=%d not a multiple of %d", B, bn);
  if (a->n_groups < 1 || a->n_groups > DRASIC_MAX_GROUPS) return fail(DRASIC_ERR_INVALID, "convlstm_fwd: n_groups");
  if (a->n_groups > 1 && !a->group_mtile_end) return fail(DRASIC_ERR_INVALID, "convlstm_fwd: group_mtile_end");
  if (!a->x_hi || !a->w_hi || !a->w_inv_scale || !a->c_state || !a->h_out_hi)
    return fail(DRASIC_ERR_INVALID, "convlstm_fwd: null pointer");
  if (a->split && (!a->x_lo || !a->w_lo || !a->h_out_lo)) return fail(DRASIC_ERR_INVALID, "convlstm_fwd: split needs lo planes");
  if (a->has_state && (!a->h_hi || (a->split && !a->h_lo))) return fail(DRASIC_ERR_INVALID, "convlstm_fwd: has_state needs h");
  if (a->has_state && (a->h_hi == a->h_out_hi)) return fail(DRASIC_ERR_INVALID, "convlstm_fwd: h_out aliases h");
  if (a->next_mode != DRASIC_NEXT_NONE && (!a->next_hi || (a->split && !a->next_f32 && !a->next_lo)))
    return fail(DRASIC_ERR_INVALID, "convlstm_fwd: next buffers");
  if (a->next_mode == DRASIC_NEXT_DOWN && ((H | W) & 1)) return fail(DRASIC_ERR_INVALID, "convlstm_fwd: odd size for unshuffle");
  if (a->next_mode == DRASIC_NEXT_UP && (Co / 4) % 16) return fail(DRASIC_ERR_INVALID, "convlstm_fwd: Co/4 %% 16 for shuffle");

  drasic::GateParams p;
  memset(&p, 0, sizeof(p));
  int rc;
  if ((rc = make_act_map(&p.tm_x[0], a->x_hi, B, H, W, Cin, bw, bh, bn))) return rc;
  if (a->split && (rc = make_act_map(&p.tm_x[1], a->x_lo, B, H, W, Cin, bw, bh, bn))) return rc;
  if (a->has_state) {
    if ((rc = make_act_map(&p.tm_h[0], a->h_hi, B, H, W, Co, bw, bh, bn))) return rc;
    if (a->split && (rc = make_act_map(&p.tm_h[1], a->h_lo, B, H, W, Co, bw, bh, bn))) return rc;
  }
  const int K = 9 * (Cin + Co);
  const bool pair = a->cta_pair != 0;
  if (pair && a->block_n != 256) return fail(DRASIC_ERR_INVALID, "convlstm_fwd: cta_pair needs block_n == 256");
  if (a->n_units <= 0 || (a->n_groups > 1 && (!a->group_mtile_start || !a->group_unit_end || !a->group_img_end)))
    return fail(DRASIC_ERR_INVALID, "convlstm_fwd: n_units and (n_groups > 1) group_mtile_start / group_unit_end / group_img_end");
  const int w_box_rows = pair ? a->block_n / 2 : a->block_n;   // each CTA of a pair stages half of the N tile
  const long long w_rows = static_cast<long long>(a->n_groups) * 4 * Co * (K / 64);
  if ((rc = make_weight_map(&p.tm_w[0], a->w_hi, w_rows, w_box_rows))) return rc;
  if (a->split && (rc = make_weight_map(&p.tm_w[1], a->w_lo, w_rows, w_box_rows))) return rc;
  p.c_state = a->c_state;
  p.c_prev = a->c_prev;
  p.gates_out = a->gates_out;
  p.gates_f32 = a->gates_f32;
  p.w_group_tiles = 4 * Co / a->block_n;
  p.w_nkb = K / 64;
  p.h_out[0] = static_cast<__half*>(a->h_out_hi);
  p.h_out[1] = static_cast<__half*>(a->h_out_lo);
  p.next[0] = a->next_hi;
  p.next[1] = a->next_lo;
  p.w_inv_scale = a->w_inv_scale;
  p.H = H; p.W = W; p.Cin = Cin; p.Co = Co;
  p.bw = bw; p.bh = bh; p.bn = bn;
  p.n_mtiles = static_cast<int>(static_cast<long long>(B) * H * W / 128);
  p.n_ntiles = 4 * Co / a->block_n;
#ifdef DRASIC_DIAG
  { static const int ord = getenv("DRASIC_TILE_ORDER") ? atoi(getenv("DRASIC_TILE_ORDER")) : 0; p.tile_order = ord; }
#endif
  p.has_hidden = a->has_state ? 1 : 0;
  p.has_cell = a->has_state ? 1 : 0;
  p.next_mode = a->next_mode;
  p.next_f32 = a->next_f32;
  p.precise = a->precise;
  p.n_groups = a->n_groups;
  p.group_mtile_end = a->group_mtile_end;
  p.group_mtile_start = a->group_mtile_start; p.group_unit_end = a->group_unit_end; p.group_img_end = a->group_img_end;
  p.n_munits = a->n_units;
  if ((rc = setup_pos_tiles(p, "convlstm_fwd", a->pos_order, a->pos_imgs, a->pos_pairs, a->n_groups, pair, B, H, W, a->n_units))) return rc;
  if (p.pos_units) {
    if ((rc = make_act_map(&p.tm_xp[0], a->x_hi, B, H, W, Cin, 1, 1, 128))) return rc;
    if (a->split && (rc = make_act_map(&p.tm_xp[1], a->x_lo, B, H, W, Cin, 1, 1, 128))) return rc;
    if (a->has_state) {
      if ((rc = make_act_map(&p.tm_hp[0], a->h_hi, B, H, W, Co, 1, 1, 128))) return rc;
      if (a->split && (rc = make_act_map(&p.tm_hp[1], a->h_lo, B, H, W, Co, 1, 1, 128))) return rc;
    }
  }
  const int sms = num_sms();
  if (sms <= 0) return fail(DRASIC_ERR_NO_DEVICE, "no CUDA device");
  if (a->stream_k && (rc = setup_stream_k(p, "convlstm_fwd", a->stream_k, a->sk_scratch, a->sk_flags,
                                          Cin / 64 + (a->has_state ? Co / 64 : 0),
                                          drasic::conv_gemm_max_workers(a->split != 0, a->block_n, false, pair, sms))))
    return rc;
  return cuda_fail(drasic::launch_convlstm_gates(p, a->split != 0, a->block_n, pair, sms, static_cast<cudaStream_t>(stream)),
                   "convlstm_fwd launch");
}

int drasic_bottleneck_fwd(const drasic_bottleneck_args* a, void* stream) {
  if (!a) return fail(DRASIC_ERR_INVALID, "bottleneck_fwd: null args");
  if (!a->w_d0 || !a->perm || !a->code_out || !a->d1_in_hi || (a->split && !a->d1_in_lo) ||
      (!a->bits_in && (!a->h_e3 || !a->w_e4)))
    return fail(DRASIC_ERR_INVALID, "bottleneck_fwd: null pointer");
  if (a->B_pad <= 0 || a->HW <= 0 || a->Ch % 4 || a->code <= 0 || (a->code * a->HW) % 32)
    return fail(DRASIC_ERR_INVALID, "bottleneck_fwd: shape B_pad=%d HW=%d Ch=%d code=%d", a->B_pad, a->HW, a->Ch, a->code);
  if (a->n_enc_groups < 1 || a->n_dec_groups < 1 || a->n_enc_groups > DRASIC_MAX_GROUPS || a->n_dec_groups > DRASIC_MAX_GROUPS)
    return fail(DRASIC_ERR_INVALID, "bottleneck_fwd: groups");
  if ((a->n_enc_groups > 1 || a->n_dec_groups > 1) && !a->group_img_end)
    return fail(DRASIC_ERR_INVALID, "bottleneck_fwd: group_img_end");
  drasic::BottleneckParams p;
  memset(&p, 0, sizeof(p));
  p.h_e3 = a->h_e3; p.w_e4 = a->w_e4; p.w_d0 = a->w_d0; p.noise = a->noise; p.perm = a->perm;
  p.code_out = a->code_out; p.z_out = a->z_out; p.bits_out = a->bits_out;
  p.d1_in[0] = static_cast<__half*>(a->d1_in_hi); p.d1_in[1] = static_cast<__half*>(a->d1_in_lo);
  p.d0_out_f32 = a->d0_out_f32;
  p.B_pad = a->B_pad; p.HW = a->HW; p.Ch = a->Ch; p.code = a->code;
  p.n_enc_groups = a->n_enc_groups; p.n_dec_groups = a->n_dec_groups;
  p.precise = a->precise; p.split = a->split;
  p.group_img_end = a->group_img_end;
  p.bits_in = a->bits_in;
  return cuda_fail(drasic::launch_bottleneck(p, static_cast<cudaStream_t>(stream)), "bottleneck_fwd launch");
}

int drasic_tail_head_fwd(const drasic_tail_head_args* a, void* stream) {
  if (!a) return fail(DRASIC_ERR_INVALID, "tail_head_fwd: null args");
  if (!a->perm) return fail(DRASIC_ERR_INVALID, "tail_head_fwd: null pointer");
  const bool decode_only = !a->img;
  if (decode_only ? (a->mode == DRASIC_TAIL_INIT || a->e1_in_hi) : !a->w_e0)
    return fail(DRASIC_ERR_INVALID, "tail_head_fwd: img is required unless decoding only (mode != INIT, no e1_in)");
  if (a->mode != DRASIC_TAIL_INIT && (!a->d3_next || !a->w_d4 || !a->recon_out || (!decode_only && !a->loss_acc)))
    return fail(DRASIC_ERR_INVALID, "tail_head_fwd: step mode needs decoder buffers");
  if (a->mode == DRASIC_TAIL_STEP && !a->recon_prev) return fail(DRASIC_ERR_INVALID, "tail_head_fwd: recon_prev");
  if (a->mode < 0 || a->mode > 2 || a->loss_kind < 0 || a->loss_kind > 2) return fail(DRASIC_ERR_INVALID, "tail_head_fwd: mode/loss");
  if (a->C < 1 || a->C > 4 || (a->H * a->W) % 256 || ((a->H | a->W) & 1)) return fail(DRASIC_ERR_INVALID, "tail_head_fwd: C=%d H=%d W=%d", a->C, a->H, a->W);
  if (a->e1_in_hi && a->split && !a->e1_in_lo) return fail(DRASIC_ERR_INVALID, "tail_head_fwd: e1_in_lo");
  if ((a->n_enc_groups > 1 || a->n_dec_groups > 1) && !a->group_img_end)
    return fail(DRASIC_ERR_INVALID, "tail_head_fwd: group_img_end");
  drasic::TailParams p;
  memset(&p, 0, sizeof(p));
  p.d3_next = a->d3_next; p.w_d4 = a->w_d4; p.w_e0 = a->w_e0; p.b_e0 = a->b_e0; p.img = a->img;
  p.recon_prev = a->recon_prev; p.recon_out = a->recon_out; p.dec_out = a->dec_out; p.loss_acc = a->loss_acc;
  p.e1_in[0] = static_cast<__half*>(a->e1_in_hi); p.e1_in[1] = static_cast<__half*>(a->e1_in_lo);
  p.perm = a->perm;
  p.B_pad = a->B_pad; p.C = a->C; p.H = a->H; p.W = a->W; p.mode = a->mode; p.loss_kind = a->loss_kind;
  p.n_enc_groups = a->n_enc_groups < 1 ? 1 : a->n_enc_groups;
  p.n_dec_groups = a->n_dec_groups < 1 ? 1 : a->n_dec_groups;
  p.precise = a->precise; p.split = a->split;
  p.group_img_end = a->group_img_end;
  p.group_iters = a->group_iters; p.iter = a->iter;
  if (a->group_iters && p.n_enc_groups > 1 && !a->group_img_end) return fail(DRASIC_ERR_INVALID, "tail_head_fwd: group_iters needs group_img_end");
  return cuda_fail(drasic::launch_tail_head(p, static_cast<cudaStream_t>(stream)), "tail_head_fwd launch");
}


int drasic_lstm_bwd_fused(const drasic_lstm_bwd_fused_args* a, void* stream) {
  if (!a || !a->gates || !a->c_t || !a->dh_above || !a->dc_out || !a->dg_hi || !a->prev_amax_bits || !a->amax_bits ||
      !a->inv_scale_out || a->Co % 8 || a->n_pix == 0 || a->dc_in == a->dc_out)
    return fail(DRASIC_ERR_INVALID, "lstm_bwd_fused: bad argument");
  drasic::LstmBwdFusedParams p;
  p.gates = a->gates; p.gates_f32 = a->gates_f32; p.c_t = a->c_t; p.c_prev = a->c_prev; p.dh_above = a->dh_above;
  p.dh_rec = a->dh_rec; p.dc_in = a->dc_in; p.dc_out = a->dc_out;
  p.hi = static_cast<__half*>(a->dg_hi); p.lo = static_cast<__half*>(a->dg_lo);
  p.prev_amax_bits = static_cast<const unsigned int*>(a->prev_amax_bits);
  p.amax_bits = static_cast<unsigned int*>(a->amax_bits); p.inv_scale_out = a->inv_scale_out;
  p.n_pix = a->n_pix; p.Co = a->Co; p.verify = a->verify;
  return cuda_fail(drasic::launch_lstm_bwd_fused(p, static_cast<cudaStream_t>(stream)), "lstm_bwd_fused launch");
}

int drasic_pack_dgrad_weights_batch(const float* const* w_in, const float* const* w_h, const int* slots, int n, int n_slots,
                                    int Cin, int Co, int in_order, int out_order, void* out_hi, void* out_lo,
                                    float* inv_scale_out, void* scratch, void* stream) {
  if (!out_hi || !out_lo || !inv_scale_out || !scratch) return fail(DRASIC_ERR_INVALID, "pack_dgrad_weights: null pointer");
  if (Cin % 64 || Co % 64) return fail(DRASIC_ERR_INVALID, "pack_dgrad_weights: Cin=%d Co=%d", Cin, Co);
  drasic::PackBatch b;
  if (int rc = fill_pack_batch(b, "pack_dgrad_weights", w_in, w_h, slots, n, n_slots)) return rc;
  return cuda_fail(drasic::launch_pack_dgrad_weights(b, n, n_slots, Cin, Co, in_order, out_order, static_cast<__half*>(out_hi),
                                                     static_cast<__half*>(out_lo), inv_scale_out,
                                                     static_cast<unsigned int*>(scratch), static_cast<cudaStream_t>(stream)),
                   "pack_dgrad_weights launch");
}

int drasic_pack_dgrad_weights(const float* w_in, const float* w_h, int Cin, int Co, int in_order,
                              int out_order, void* out_hi, void* out_lo, float* inv_scale_out,
                              void* scratch4, void* stream) {
  if (!w_in || !w_h) return fail(DRASIC_ERR_INVALID, "pack_dgrad_weights: null pointer");
  const int slot = 0;
  return drasic_pack_dgrad_weights_batch(&w_in, &w_h, &slot, 1, 1, Cin, Co, in_order, out_order, out_hi, out_lo, inv_scale_out,
                                         scratch4, stream);
}

int drasic_convlstm_bwd_data(const drasic_dgrad_args* a, void* stream) {
  if (!a) return fail(DRASIC_ERR_INVALID, "convlstm_bwd_data: null args");
  const int H = a->H, W = a->W, Cin = a->Cin, Co = a->Co, B = a->B_pad;
  if (Cin % 64 || Co % 64 || Cin <= 0 || Co <= 0) return fail(DRASIC_ERR_INVALID, "convlstm_bwd_data: channels");
  int bw, bh, bn;
  if (pixel_box(H, W, 128, &bw, &bh, &bn)) return fail(DRASIC_ERR_INVALID, "convlstm_bwd_data: H=%d W=%d unsupported", H, W);
  if (B <= 0 || B % bn) return fail(DRASIC_ERR_INVALID, "convlstm_bwd_data: B_pad=%d", B);
  if (!a->dg_hi || !a->w_hi || !a->w_inv_scale || !a->dg_inv_scale || !a->dx_out || (a->split && (!a->dg_lo || !a->w_lo)))
    return fail(DRASIC_ERR_INVALID, "convlstm_bwd_data: null pointer");
  if (a->n_groups < 1 || a->n_groups > DRASIC_MAX_GROUPS || (a->n_groups > 1 && !a->group_mtile_end))
    return fail(DRASIC_ERR_INVALID, "convlstm_bwd_data: groups");
  if (a->dx_mode < 0 || a->dx_mode > 2 || (a->dx_mode == DRASIC_DX_INV_DOWN && Cin % 64) || (a->dx_mode == DRASIC_DX_INV_UP && ((H | W) & 1)))
    return fail(DRASIC_ERR_INVALID, "convlstm_bwd_data: dx_mode");
  drasic::GateParams p;
  memset(&p, 0, sizeof(p));
  int rc;
  const int Cg = 4 * Co, Nout = Cin + Co, K = 9 * Cg;
  if ((rc = make_act_map(&p.tm_x[0], a->dg_hi, B, H, W, Cg, bw, bh, bn))) return rc;
  if (a->split && (rc = make_act_map(&p.tm_x[1], a->dg_lo, B, H, W, Cg, bw, bh, bn))) return rc;
  const bool pair = a->cta_pair != 0;
  if (a->n_units <= 0 || (a->n_groups > 1 && (!a->group_mtile_start || !a->group_unit_end || !a->group_img_end)))
    return fail(DRASIC_ERR_INVALID, "convlstm_bwd_data: n_units and (n_groups > 1) group_mtile_start / group_unit_end / group_img_end");
  const int w_tiles = (Nout + 255) / 256;        // a group's plane holds whole 256-row tiles (drasic_pack_dgrad_weights)
  const long long w_rows = static_cast<long long>(a->n_groups) * w_tiles * 256 * (K / 64);
  if ((rc = make_weight_map(&p.tm_w[0], a->w_hi, w_rows, pair ? 128 : 256))) return rc;
  if (a->split && (rc = make_weight_map(&p.tm_w[1], a->w_lo, w_rows, pair ? 128 : 256))) return rc;
  p.w_inv_scale = a->w_inv_scale; p.a_inv_scale = a->dg_inv_scale;
  p.H = H; p.W = W; p.Cin = Cg; p.Co = Co;
  p.bw = bw; p.bh = bh; p.bn = bn;
  p.n_mtiles = static_cast<int>(static_cast<long long>(B) * H * W / 128);
  // without a dH_{t-1} destination (t == 0) only the dX columns are computed
  const int Nuse = a->dhrec_out ? Nout : Cin;
  p.n_ntiles = (Nuse + 255) / 256;
#ifdef DRASIC_DIAG
  { static const int ord = getenv("DRASIC_TILE_ORDER") ? atoi(getenv("DRASIC_TILE_ORDER")) : 0; p.tile_order = ord; }
#endif
  p.n_groups = a->n_groups; p.group_mtile_end = a->group_mtile_end;
  p.group_mtile_start = a->group_mtile_start; p.group_unit_end = a->group_unit_end; p.group_img_end = a->group_img_end;
  p.n_munits = a->n_units;
  p.w_group_tiles = w_tiles; p.w_nkb = K / 64; p.n_out = Nuse; p.n_x = Cin; p.dx_mode = a->dx_mode;
  p.dx_out = a->dx_out; p.dhrec_out = a->dhrec_out;
  if ((rc = setup_pos_tiles(p, "convlstm_bwd_data", a->pos_order, a->pos_imgs, a->pos_pairs, a->n_groups, pair, B, H, W, a->n_units))) return rc;
  if (p.pos_units) {
    if ((rc = make_act_map(&p.tm_xp[0], a->dg_hi, B, H, W, Cg, 1, 1, 128))) return rc;
    if (a->split && (rc = make_act_map(&p.tm_xp[1], a->dg_lo, B, H, W, Cg, 1, 1, 128))) return rc;
  }
  const int sms = num_sms();
  if (sms <= 0) return fail(DRASIC_ERR_NO_DEVICE, "no CUDA device");
  if (a->stream_k && (rc = setup_stream_k(p, "convlstm_bwd_data", a->stream_k, a->sk_scratch, a->sk_flags, Cg / 64,
                                          drasic::conv_gemm_max_workers(a->split != 0, 256, true, pair, sms))))
    return rc;
  return cuda_fail(drasic::launch_convlstm_dgrad(p, a->split != 0, pair, sms, static_cast<cudaStream_t>(stream)), "convlstm_bwd_data launch");
}

int drasic_convlstm_bwd_weight(const drasic_wgrad_args* a, void* stream) {
  if (!a) return fail(DRASIC_ERR_INVALID, "convlstm_bwd_weight: null args");
  const int H = a->H, W = a->W, Cin = a->Cin, Co = a->Co, B = a->B_pad;
  if (Cin % 128 || Co % 128 || Cin <= 0 || Co <= 0) return fail(DRASIC_ERR_INVALID, "convlstm_bwd_weight: channels must be multiples of 128 (Cin=%d Co=%d)", Cin, Co);
  int kbw, kbh, kbn;
  if (pixel_box(H, W, 64, &kbw, &kbh, &kbn)) return fail(DRASIC_ERR_INVALID, "convlstm_bwd_weight: H=%d W=%d unsupported", H, W);
  if (B <= 0 || B % kbn) return fail(DRASIC_ERR_INVALID, "convlstm_bwd_weight: B_pad=%d", B);
  if (!a->dg_hi || !a->x_hi || !a->gw || !a->dg_inv_scale || (a->split && (!a->dg_lo || !a->x_lo)))
    return fail(DRASIC_ERR_INVALID, "convlstm_bwd_weight: null pointer");
  if (a->has_h && (!a->h_hi || (a->split && !a->h_lo))) return fail(DRASIC_ERR_INVALID, "convlstm_bwd_weight: h planes");
  if (a->n_groups < 1 || a->n_groups > DRASIC_MAX_GROUPS || (a->n_groups > 1 && !a->group_kb_end) || a->k_splits < 1)
    return fail(DRASIC_ERR_INVALID, "convlstm_bwd_weight: groups / k_splits");
  drasic::WgradParams p;
  memset(&p, 0, sizeof(p));
  int rc;
  if ((rc = make_act_map_chunked(&p.tm_g[0], a->dg_hi, B, H, W, 4 * Co, kbw, kbh, kbn, 2))) return rc;
  if (a->split && (rc = make_act_map_chunked(&p.tm_g[1], a->dg_lo, B, H, W, 4 * Co, kbw, kbh, kbn, 2))) return rc;
  if ((rc = make_act_map_chunked(&p.tm_x[0], a->x_hi, B, H, W, Cin, kbw, kbh, kbn, 2))) return rc;
  if (a->split && (rc = make_act_map_chunked(&p.tm_x[1], a->x_lo, B, H, W, Cin, kbw, kbh, kbn, 2))) return rc;
  if (a->has_h) {
    if ((rc = make_act_map_chunked(&p.tm_h[0], a->h_hi, B, H, W, Co, kbw, kbh, kbn, 2))) return rc;
    if (a->split && (rc = make_act_map_chunked(&p.tm_h[1], a->h_lo, B, H, W, Co, kbw, kbh, kbn, 2))) return rc;
  }
  p.gw = a->gw; p.g_inv_scale = a->dg_inv_scale; p.group_kb_end = a->n_groups > 1 ? a->group_kb_end : nullptr;
  p.H = H; p.W = W; p.Cin = Cin; p.Co = Co; p.kbw = kbw; p.kbh = kbh; p.kbn = kbn;
  p.n_kb = static_cast<int>(static_cast<long long>(B) * H * W / 64);
  p.has_h = a->has_h ? 1 : 0; p.n_groups = a->n_groups; p.m_tiles = 4 * Co / 128;
  p.wide = (a->cta_pair == 2) ? 1 : 0;      // cta_pair: 0 single CTAs, 1 pairs with 256-column tiles, 2 pairs with 512-column tiles
  const int upt = p.wide ? 8 : 4;
  p.n_tiles_x = (9 * (Cin / 64) + upt - 1) / upt; p.n_tiles_h = (9 * (Co / 64) + upt - 1) / upt; p.k_splits = a->k_splits;
  {
#ifdef DRASIC_DIAG
    static const int dbg = getenv("DRASIC_WGRAD_DBG") ? atoi(getenv("DRASIC_WGRAD_DBG")) : 0;
    p.dbg = dbg;
#else
    p.dbg = 0;
#endif
  }
  const int sms = num_sms();
  if (sms <= 0) return fail(DRASIC_ERR_NO_DEVICE, "no CUDA device");
  return cuda_fail(drasic::launch_wgrad(p, a->split != 0, a->cta_pair != 0, sms, static_cast<cudaStream_t>(stream)), "convlstm_bwd_weight launch");
}

int drasic_unpack_wgrad(const float* gw, int Cin, int Co, int in_order, int out_order, float* dw_in,
                        float* dw_h, void* stream) {
  if (!gw || !dw_in || !dw_h || Cin <= 0 || Co <= 0) return fail(DRASIC_ERR_INVALID, "unpack_wgrad: bad argument");
  return cuda_fail(drasic::launch_unpack_wgrad(gw, Cin, Co, in_order, out_order, dw_in, dw_h, static_cast<cudaStream_t>(stream)), "unpack_wgrad launch");
}

int drasic_bottleneck_bwd(const drasic_bottleneck_bwd_args* a, void* stream) {
  if (!a || !a->dd0 || !a->h_e3 || !a->z || !a->codes || !a->w_e4 || !a->w_d0 || !a->perm || !a->dh_e3 || !a->dw_e4 || !a->dw_d0)
    return fail(DRASIC_ERR_INVALID, "bottleneck_bwd: null pointer");
  if (a->B_pad <= 0 || a->HW <= 0 || a->Ch <= 0 || a->code <= 0) return fail(DRASIC_ERR_INVALID, "bottleneck_bwd: shape");
  if (a->HW > 32) return fail(DRASIC_ERR_INVALID, "bottleneck_bwd: HW=%d -- un-tiled images are inference-only; train on patches of at most 32 code pixels", a->HW);
  if ((a->n_enc_groups > 1 || a->n_dec_groups > 1) && !a->group_img_end) return fail(DRASIC_ERR_INVALID, "bottleneck_bwd: group_img_end");
  drasic::BottleneckBwdParams p;
  p.dd0 = a->dd0; p.h_e3 = a->h_e3; p.z = a->z; p.codes = a->codes; p.w_e4 = a->w_e4; p.w_d0 = a->w_d0;
  p.perm = a->perm; p.group_img_end = a->group_img_end; p.dh_e3 = a->dh_e3; p.dw_e4 = a->dw_e4; p.dw_d0 = a->dw_d0;
  p.B_pad = a->B_pad; p.HW = a->HW; p.Ch = a->Ch; p.code = a->code;
  p.n_enc_groups = a->n_enc_groups < 1 ? 1 : a->n_enc_groups; p.n_dec_groups = a->n_dec_groups < 1 ? 1 : a->n_dec_groups;
  p.fix_scale = a->fix_scale;
  return cuda_fail(drasic::launch_bottleneck_bwd(p, static_cast<cudaStream_t>(stream)), "bottleneck_bwd launch");
}

int drasic_tail_bwd(const drasic_tail_bwd_args* a, void* stream) {
  if (!a || !a->d3_next || !a->w_d4 || !a->img || !a->recon_t || !a->gacc || !a->dh_d3 || !a->dw_d4 || !a->perm)
    return fail(DRASIC_ERR_INVALID, "tail_bwd: null pointer");
  if (a->C < 1 || a->C > 4 || (a->H * a->W) % 256 || ((a->H | a->W) & 1) || a->loss_kind < 0 || a->loss_kind > 2)
    return fail(DRASIC_ERR_INVALID, "tail_bwd: shape / loss");
  if (a->n_dec_groups > 1 && !a->group_img_end) return fail(DRASIC_ERR_INVALID, "tail_bwd: group_img_end");
  drasic::TailBwdParams p;
  p.d3_next = a->d3_next; p.w_d4 = a->w_d4; p.img = a->img; p.recon_t = a->recon_t; p.gacc = a->gacc;
  p.dh_d3 = a->dh_d3; p.dw_d4 = a->dw_d4; p.perm = a->perm; p.group_img_end = a->group_img_end;
  p.loss_scale = a->loss_scale; p.B_pad = a->B_pad; p.C = a->C; p.H = a->H; p.W = a->W; p.first = a->first;
  p.loss_kind = a->loss_kind; p.n_dec_groups = a->n_dec_groups < 1 ? 1 : a->n_dec_groups; p.has_acc = a->has_acc;
  p.group_iters = a->group_iters; p.iter = a->iter; p.n_enc_groups = a->n_enc_groups < 1 ? 1 : a->n_enc_groups;
  if (a->group_iters && p.n_enc_groups > 1 && !a->group_img_end) return fail(DRASIC_ERR_INVALID, "tail_bwd: group_iters needs group_img_end");
  p.fix_scale = a->fix_scale;
  return cuda_fail(drasic::launch_tail_bwd(p, static_cast<cudaStream_t>(stream)), "tail_bwd launch");
}

int drasic_head_bwd(const drasic_head_bwd_args* a, void* stream) {
  if (!a || !a->du1 || !a->w_e0 || !a->img || !a->gacc || !a->dw_e0 || !a->perm) return fail(DRASIC_ERR_INVALID, "head_bwd: null pointer");
  if (!a->first && !a->recon_prev) return fail(DRASIC_ERR_INVALID, "head_bwd: recon_prev");
  if (a->C < 1 || a->C > 4 || (a->H * a->W) % 256 || ((a->H | a->W) & 1)) return fail(DRASIC_ERR_INVALID, "head_bwd: shape");
  if (a->n_enc_groups > 1 && !a->group_img_end) return fail(DRASIC_ERR_INVALID, "head_bwd: group_img_end");
  drasic::HeadBwdParams p;
  p.du1 = a->du1; p.w_e0 = a->w_e0; p.b_e0 = a->b_e0; p.img = a->img; p.recon_prev = a->recon_prev; p.gacc = a->gacc;
  p.dw_e0 = a->dw_e0; p.db_e0 = a->db_e0; p.perm = a->perm; p.group_img_end = a->group_img_end;
  p.B_pad = a->B_pad; p.C = a->C; p.H = a->H; p.W = a->W; p.first = a->first; p.propagate = a->propagate;
  p.n_enc_groups = a->n_enc_groups < 1 ? 1 : a->n_enc_groups;
  p.fix_scale = a->fix_scale;
  return cuda_fail(drasic::launch_head_bwd(p, static_cast<cudaStream_t>(stream)), "head_bwd launch");
}

int drasic_conv1x1_fwd(const float* x, const float* w, const float* b, float* y, int N, int Cin,
                       int Cout, int HW, int act, void* stream) {
  if (!x || !w || !y || N <= 0 || Cin <= 0 || Cout <= 0 || HW <= 0) return fail(DRASIC_ERR_INVALID, "conv1x1_fwd: bad argument");
  if (act < 0 || act > 7) return fail(DRASIC_ERR_INVALID, "conv1x1_fwd: activation %d", act);
  return cuda_fail(drasic::launch_conv1x1(x, w, b, y, N, Cin, Cout, HW, act, static_cast<cudaStream_t>(stream)), "conv1x1_fwd launch");
}

int drasic_quantize_fwd(const float* x, const float* u, float* y, size_t n, void* stream) {
  if (!x || !y) return fail(DRASIC_ERR_INVALID, "quantize_fwd: null pointer");
  if (n == 0) return 0;
  return cuda_fail(drasic::launch_quantize(x, u, y, n, static_cast<cudaStream_t>(stream)), "quantize_fwd launch");
}

int drasic_adam_step(const drasic_adam_tensor* tensors, int n_tensors, double lr, double beta1, double beta2,
                     double eps, double weight_decay, double bias_correction1, double bias_correction2,
                     void* stream) {
  static_assert(sizeof(drasic_adam_tensor) == sizeof(drasic::AdamTensor), "ABI / kernel struct mismatch");
  if (n_tensors < 0 || (n_tensors > 0 && !tensors)) return fail(DRASIC_ERR_INVALID, "adam_step: null tensor list");
  if (!(lr >= 0.0) || !(eps >= 0.0) || !(beta1 >= 0.0 && beta1 < 1.0) || !(beta2 >= 0.0 && beta2 < 1.0) ||
      !(weight_decay >= 0.0))
    return fail(DRASIC_ERR_INVALID, "adam_step: lr=%g betas=(%g, %g) eps=%g weight_decay=%g", lr, beta1, beta2, eps, weight_decay);
  if (!(bias_correction1 > 0.0) || !(bias_correction2 > 0.0))
    return fail(DRASIC_ERR_INVALID, "adam_step: bias corrections must be positive (step >= 1)");
  for (int i = 0; i < n_tensors; ++i) {
    const drasic_adam_tensor& t = tensors[i];
    if (t.numel < 0 || (t.numel > 0 && (!t.param || !t.grad || !t.exp_avg || !t.exp_avg_sq)))
      return fail(DRASIC_ERR_INVALID, "adam_step: tensor %d has a null pointer", i);
  }
  drasic::AdamHyper h;
  h.lr = lr; h.beta1 = beta1; h.beta2 = beta2; h.eps = eps; h.weight_decay = weight_decay;
  h.bias_correction1 = bias_correction1; h.bias_correction2 = bias_correction2;
  return cuda_fail(drasic::launch_adam(reinterpret_cast<const drasic::AdamTensor*>(tensors), n_tensors, h,
                                       static_cast<cudaStream_t>(stream)), "adam_step launch");
}

}  // extern "C"
