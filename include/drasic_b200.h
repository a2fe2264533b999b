/*
 * drasic_b200.h -- C ABI of the B200 (sm_100a) implementation of DRASIC's recurrent residual
 * coding loop (ConvLSTM encoder -> sign binarizer -> ConvLSTM decoder, iterated T times).
 *
 * The reference (diaoenmao/DRASIC, pure PyTorch) has no FFI of its own; the entry points below are
 * what a binding for this path replaces, one per fused kernel, cited as file:line of the reference
 * under src/.  All functions return 0 on success or a negative drasic_status / positive
 * cudaError_t value; drasic_last_error() gives the message of the last failure on the calling
 * thread.  No function throws, allocates device memory, or synchronises the device: every buffer
 * is owned by the caller, every launch is asynchronous on `stream` (a cudaStream_t).  Pointers
 * are DEVICE pointers unless the name ends in `_host`.
 *
 * Layouts.  Activations and recurrent state are NHWC.  An activation tensor a (|a| <= 1) is stored
 * as one or two fp16 planes: hi = fp16(a * 2^8) and, in split mode, lo = fp16(a * 2^8 - hi).
 * Channels may be stored in "quad" order (DRASIC_ORDER_QUAD), the order in which the fused
 * pixel (un)shuffle stores are contiguous: physical p = q*(C/4) + c  <->  logical channel c*4 + q.
 * The batch is node-sorted and padded: slot s holds original image perm[s] (or -1), every node's
 * slot range is a multiple of 4 slots (the total a multiple of 8), group_img_end[g] is the exclusive end slot of node g.
 */
#ifndef DRASIC_B200_H_
#define DRASIC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DRASIC_ABI_VERSION 1
#define DRASIC_MAX_GROUPS 64

typedef enum {
  DRASIC_OK = 0,
  DRASIC_ERR_INVALID = -1,     /* bad argument / unsupported shape (mirrors the reference's ValueError) */
  DRASIC_ERR_NO_DEVICE = -2,   /* no sm_100 device / driver entry point unavailable */
  DRASIC_ERR_TENSORMAP = -3    /* cuTensorMapEncodeTiled rejected a descriptor */
} drasic_status;

enum { DRASIC_ORDER_IDENT = 0, DRASIC_ORDER_QUAD = 1 };
enum { DRASIC_NEXT_NONE = 0, DRASIC_NEXT_SAME = 1, DRASIC_NEXT_DOWN = 2, DRASIC_NEXT_UP = 3 };
enum { DRASIC_LOSS_MAE = 0, DRASIC_LOSS_MSE = 1, DRASIC_LOSS_BCE = 2 };
enum { DRASIC_TAIL_INIT = 0, DRASIC_TAIL_FIRST = 1, DRASIC_TAIL_STEP = 2 };

int drasic_version(void);
const char* drasic_last_error(void);
/* number of SMs of the current device (grid sizing); <0 on error */
int drasic_device_sms(void);

/* Programmatic dependent launch of the loop's kernels (EXPERIMENTAL, off by default): each kernel may be scheduled
 * and run its prologue while the tail of its predecessor in the stream still executes, and waits
 * (griddepcontrol.wait) before it touches global memory.  Results are identical either way and inference gains 1-2 %,
 * but the training step -- main-stream kernels next to the weight-gradient GEMMs of a side stream -- was observed to
 * hang intermittently with it on.  Process-wide switch; returns the previous setting. */
int drasic_set_pdl(int enable);

/* bytes of one packed fp16 plane for a ConvLSTM cell: 4*Co * 9*(Cin+Co) * 2 */
size_t drasic_packed_weight_bytes(int Cin, int Co);
/* sizes of the stream-K work buffers of drasic_convlstm_fwd / drasic_convlstm_bwd_data (see stream_k below) */
size_t drasic_stream_k_scratch_bytes(void);
size_t drasic_stream_k_flag_bytes(void);

/*
 * Pack the gate weights of one ConvLSTMCell (modules/cell.py:82-100: cell[0]['in'].weight
 * [4Co,Cin,3,3], cell[0]['hidden'].weight [4Co,Co,3,3], fp32 OIHW, no bias) into the K-major,
 * gate-interleaved fp16 {hi,lo} planes the gate kernel reads, scaled by a power of two.  The planes are stored
 * tile by tile -- [N tile of block_n rows][K block of 64][block_n][64] -- so that the operand tile of one TMA box is
 * one contiguous run (weights that do not fit the L2 stream from HBM at full rate).
 * out_hi/out_lo: drasic_packed_weight_bytes each; inv_scale_out: 1 float; scratch: 4 bytes.
 */
int drasic_pack_lstm_weights(const float* w_in, const float* w_h, int Cin, int Co, int in_order,
                             int out_order, int block_n, void* out_hi, void* out_lo,
                             float* inv_scale_out, void* scratch4, void* stream);
/* The same for n weight groups (nodes) of one layer in two launches: w_in / w_h / slots are HOST arrays of n entries
 * (device pointers of each group's weights; the group's plane index in the packed arrays).  out_hi / out_lo /
 * inv_scale_out / scratch are the BASES of the [n_slots] packed arrays (drasic_packed_weight_bytes per plane, one
 * float, 4 bytes per slot); only the planes named by slots[] are rewritten. */
int drasic_pack_lstm_weights_batch(const float* const* w_in, const float* const* w_h, const int* slots, int n, int n_slots,
                                   int Cin, int Co, int in_order, int out_order, int block_n, void* out_hi, void* out_lo,
                                   float* inv_scale_out, void* scratch, void* stream);

/*
 * One ConvLSTMCell step for the whole (node-grouped) batch, replacing ConvLSTMCell.forward
 * (modules/cell.py:110-144: in-conv :117, hidden-conv :132, gates :133-137, state update :138-139)
 * and the PixelShuffleCell that follows it (modules/cell.py:147-165).
 */
typedef struct {
  const void* x_hi; const void* x_lo;   /* [B_pad,H,W,Cin] fp16 planes (lo unused unless split) */
  const void* h_hi; const void* h_lo;   /* previous h [B_pad,H,W,Co]; unused when has_state == 0 */
  const void* w_hi; const void* w_lo;   /* packed weights, n_groups consecutive cells */
  const float* w_inv_scale;             /* [n_groups] from drasic_pack_lstm_weights */
  float* c_state;                       /* [B_pad,H,W,Co] fp32 new cell state (in place when c_prev is NULL) */
  const float* c_prev;                  /* nullable: previous cell state kept in its own buffer (training) */
  void* gates_out;                      /* nullable: activated gates [B_pad,H,W,4,Co] saved for backward (fp16; fp32 when gates_f32) */
  void* h_out_hi; void* h_out_lo;       /* new h (must not alias h_hi/h_lo) */
  void* next_hi; void* next_lo;         /* consumer copy; fp32 tensor in next_hi when next_f32 */
  const int* group_mtile_end;           /* device [n_groups]: exclusive end M-tile (128 pixels) per node, see below */
  int B_pad, H, W, Cin, Co;
  int n_groups;
  int has_state;                        /* 0 on iteration 0: h = c = 0 (cell.py:118-129) */
  int next_mode, next_f32;
  int split;                            /* 1: fp16 hi+lo operands (fp32-grade), 0: fp16 only */
  int precise;                          /* 1: IEEE exp/div/tanh, 0: tanh.approx */
  int block_n;                          /* N tile: 256, 128 or 64 (must match the packing) */
  int gates_f32;                        /* dtype of gates_out */
  int cta_pair;                         /* 1: CTA pairs (cta_group::2) on 256-pixel x 256-column tiles; needs block_n == 256 */
  /* Work units of the M dimension.  n_groups > 1: device arrays of n_groups ints -- node g walks the M tiles
   * [group_mtile_start[g], group_mtile_end[g]) (neighbouring ranges may share a tile: a node's slots are a multiple
   * of 4 images, a 4x4-layer tile holds 8) as units numbered up to group_unit_end[g] (single tiles, or with
   * cta_pair two consecutive tiles of the node, an odd last one alone), and stores only the rows of the slots
   * [group_img_end[g-1], group_img_end[g]).  n_units: total units = group_unit_end[n_groups-1], or with one
   * group the M tiles (cta_pair: ceil(M tiles / 2)). */
  const int* group_mtile_start; const int* group_unit_end; const int* group_img_end;
  int n_units;
  /* Position-major M tiles (n_groups == 1 only; pos_imgs == 0 switches them off).  The first pos_imgs slots -- a
   * multiple of 128, or of 256 with cta_pair -- are tiled as ONE pixel position x 128 images per M tile, so the
   * (tile, tap) blocks of the 3x3 convolutions (padding=1, cell.py:117,132) that fall entirely into the zero padding
   * are never issued: (4-|dx|)(4-|dy|)/16 of the taps are in bounds at 4x4, i.e. 30 % of the MACs of those layers are
   * skipped, 16 % at 8x8, 8 % at 16x16.  pos_order: device [H*W] ints, the pixel positions y*W+x in the order they are
   * walked (most in-bounds taps first, so the persistent CTAs end together).  The remaining B_pad - pos_imgs slots use
   * pixel-major tiles; n_units = H*W * pos_imgs/128 (cta_pair: /256) + the units of those tiles. */
  const int* pos_order;
  int pos_imgs;
  /* Stream-K schedule (stream_k != 0): instead of dealing whole (M unit, N tile) work items to the persistent CTAs,
   * the K blocks of all items laid end to end are cut into one equal range per CTA (pair); partial accumulators
   * travel through sk_scratch (drasic_stream_k_scratch_bytes() bytes) guarded by sk_flags
   * (drasic_stream_k_flag_bytes() bytes, zero before the first launch; every launch leaves them zero).  For
   * launches with few items per CTA -- small batches, the 4x4 layers -- where whole items would leave SMs idle. */
  int stream_k;
  void* sk_scratch; void* sk_flags;
  /* pos_pairs != 0 (with cta_pair, fewer than 256 slots to tile, even planes of at least 4x4): a unit is TWO pixel
   * positions of one block of 128 slots, one per CTA of the pair.  pos_order then holds H*W/2 PAIRS (2 ints each):
   * first the (H-2)(W-2)/2 pairs of interior positions (all nine taps), then the pairs of edge positions that share
   * their tap set and the two pairs of corners (six taps in their union); pos_imgs is a multiple of 128. */
  int pos_pairs;
} drasic_convlstm_args;
int drasic_convlstm_fwd(const drasic_convlstm_args* a, void* stream);

/*
 * Encoder tail + quantizer + bit-pack + decoder head, replacing models/dist_codec.py:83-85 (ConvCell
 * 128->code), functions/quantize.py:9-18, models/dist_codec.py:16-18 (pack) and :89-91 (ConvCell
 * code->128).
 */
typedef struct {
  const float* h_e3; const float* w_e4; const float* w_d0; const float* noise; const int* perm;
  float* code_out; float* z_out; uint8_t* bits_out; void* d1_in_hi; void* d1_in_lo;
  float* d0_out_f32;
  const int* group_img_end;       /* device [max(n_enc_groups,n_dec_groups)]: exclusive end slot per node */
  int B_pad, HW, Ch, code, n_enc_groups, n_dec_groups, precise, split;
  const uint8_t* bits_in;         /* nullable.  Non-null = decode-only: the symbols come from this packed stream
                                     ([B, code*HW/8], the layout of bits_out) instead of encoder tail + quantizer;
                                     h_e3 / w_e4 are then unused (may be NULL) */
} drasic_bottleneck_args;
int drasic_bottleneck_fwd(const drasic_bottleneck_args* a, void* stream);

/*
 * Decoder tail + reconstruction / loss / residual + encoder head, replacing
 * models/dist_codec.py:101-103 (ConvCell 32->C), :56-59 and :38, :71-74 (ConvCell C->32, unshuffle).
 */
typedef struct {
  const float* d3_next;           /* [B_pad,H/2,W/2,128] fp32: the last decoder ConvLSTM's h in its own (QUAD) order, i.e.
                                     pixel-major over 2x2 quads: element ((slot*H*W + quad*4 + dy*2+dx)*32 + c) */
  const float* w_d4; const float* w_e0; const float* b_e0;   /* [n_dec_groups][C,32], [n_enc_groups][32,C], nullable [..][32] */
  const float* img; const float* recon_prev; float* recon_out;   /* [B,C,H,W] fp32, the caller's batch order */
  float* dec_out;                 /* nullable: raw tanh decoder output */
  double* loss_acc;               /* [2]: += {sum of loss terms, sum of squared error} of this iteration */
  void* e1_in_hi; void* e1_in_lo; /* [B_pad,H/2,W/2,128] fp16 planes: first encoder ConvLSTM input (same quad-major order);
                                     NULL on the last iteration */
  const int* perm;
  const int* group_img_end;       /* device, as above */
  int B_pad, C, H, W, mode, loss_kind, n_enc_groups, n_dec_groups, precise, split;
  const int* group_iters;         /* nullable: device [n_enc_groups] iteration budget per node (heterogeneous rates,
                                     BASELINE config 4): at iter >= budget the node's decoded refinement is zero */
  int iter;                       /* iteration index of this call (used with group_iters) */
  /* img == NULL (with mode != INIT and e1_in_hi == NULL) = decode-only: no loss, no residual */
} drasic_tail_head_args;
int drasic_tail_head_fwd(const drasic_tail_head_args* a, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Backward-through-time: what autograd executes for `output['loss'].backward()` (train_model.py:115)
 * over the loop of models/dist_codec.py:36-65.  Gradients travel as fp32; the tensor-core GEMMs read
 * them as fp16 {hi,lo} planes scaled by an exact per-tensor power of two (drasic_lstm_bwd_fused).
 */
enum { DRASIC_DX_SAME = 0, DRASIC_DX_INV_DOWN = 1, DRASIC_DX_INV_UP = 2 };

/* LSTM cell pointwise backward (modules/cell.py:133-139 differentiated) writing the scaled fp16 {hi,lo}
 * gate-gradient planes directly (no fp32 round trip); dc carry in/out.  The power-of-two scale is predicted from *prev_amax_bits (the amax of the
 * previously processed iteration; 0 = unknown); the true amax is recorded in *amax_bits (must be zero
 * before the first call).  Call once with verify = 0, then once with verify = 1: the second call returns
 * immediately unless the prediction left the scaled amax outside [2^4, 2^15.9], in which case it redoes
 * the tensor with the exact scale.  dc_in / dc_out must be different buffers. */
typedef struct {
  const void* gates; const float* c_t; const float* c_prev; const float* dh_above; const float* dh_rec;
  const float* dc_in; float* dc_out; void* dg_hi; void* dg_lo;
  const void* prev_amax_bits; void* amax_bits; float* inv_scale_out;
  size_t n_pix;
  int Co, gates_f32, verify;
} drasic_lstm_bwd_fused_args;
int drasic_lstm_bwd_fused(const drasic_lstm_bwd_fused_args* a, void* stream);

/* conv-transpose weights of one cell for the dgrad GEMM: rows = Cin+Co input channels (physical
 * order), K = 9 taps (flipped) x 4Co gate channels, stored tile by tile ([256-row N tile][K block of 64][256][64]);
 * planes of drasic_packed_dgrad_weight_bytes(Cin, Co) = roundup(Cin+Co, 256)*36*Co fp16 each (zero the planes once:
 * the rows that pad the last tile are never written). */
size_t drasic_packed_dgrad_weight_bytes(int Cin, int Co);
int drasic_pack_dgrad_weights(const float* w_in, const float* w_h, int Cin, int Co, int in_order,
                              int out_order, void* out_hi, void* out_lo, float* inv_scale_out,
                              void* scratch4, void* stream);
int drasic_pack_dgrad_weights_batch(const float* const* w_in, const float* const* w_h, const int* slots, int n, int n_slots,
                                    int Cin, int Co, int in_order, int out_order, void* out_hi, void* out_lo,
                                    float* inv_scale_out, void* scratch, void* stream);

/* backward-data of one ConvLSTMCell step: [dX | dH_prev] = conv3x3^T(dGates) on tcgen05, the input
 * gradient routed through the inverse pixel (un)shuffle (dx_mode) into the producer's layout. */
typedef struct {
  const void* dg_hi; const void* dg_lo;   /* [B_pad,H,W,4Co] fp16 planes of the gate gradients */
  const void* w_hi; const void* w_lo;     /* from drasic_pack_dgrad_weights, n_groups consecutive cells */
  const float* w_inv_scale;               /* [n_groups] */
  const float* dg_inv_scale;              /* device scalar from drasic_lstm_bwd_fused */
  float* dx_out;                          /* fp32, layout per dx_mode */
  float* dhrec_out;                       /* nullable: fp32 [B_pad,H,W,Co] gradient w.r.t. h_{t-1} */
  const int* group_mtile_end;
  int B_pad, H, W, Cin, Co, n_groups, dx_mode, split;
  int cta_pair;                           /* as in drasic_convlstm_args */
  const int* group_mtile_start; const int* group_unit_end; const int* group_img_end;   /* as in drasic_convlstm_args */
  int n_units;
  const int* pos_order; int pos_imgs;   /* position-major M tiles, as in drasic_convlstm_args */
  int stream_k; void* sk_scratch; void* sk_flags;   /* stream-K schedule, as in drasic_convlstm_args */
  int pos_pairs;                                    /* as in drasic_convlstm_args */
} drasic_dgrad_args;
int drasic_convlstm_bwd_data(const drasic_dgrad_args* a, void* stream);

/* backward-weight of one ConvLSTMCell step, accumulated (atomics) into the packed-order fp32 gradient
 * gw [n_groups][4Co][9*(Cin+Co)]; drasic_unpack_wgrad turns it into the OIHW gradients. */
typedef struct {
  const void* dg_hi; const void* dg_lo;   /* [B_pad,H,W,4Co] */
  const void* x_hi; const void* x_lo;     /* layer input planes [B_pad,H,W,Cin] of this iteration */
  const void* h_hi; const void* h_lo;     /* h_{t-1} planes [B_pad,H,W,Co]; unused when has_h == 0 */
  float* gw;
  const float* dg_inv_scale;
  const int* group_kb_end;                /* device [n_groups]: exclusive end 64-pixel block per node */
  int B_pad, H, W, Cin, Co, n_groups, has_h, split, k_splits;
  int cta_pair;                           /* 1: CTA pairs, each pair owns 256 gate channels x 256 columns; 2: x 512 columns
                                             (one TMEM-wide accumulator, less L2 -> SMEM traffic per MMA) */
} drasic_wgrad_args;
int drasic_convlstm_bwd_weight(const drasic_wgrad_args* a, void* stream);
int drasic_unpack_wgrad(const float* gw, int Cin, int Co, int in_order, int out_order, float* dw_in,
                        float* dw_h, void* stream);

/* backward of drasic_bottleneck_fwd (straight-through quantizer, functions/quantize.py:21-22) */
typedef struct {
  const float* dd0; const float* h_e3; const float* z; const float* codes; const float* w_e4;
  const float* w_d0; const int* perm; const int* group_img_end; float* dh_e3; float* dw_e4; float* dw_d0;
  int B_pad, HW, Ch, code, n_enc_groups, n_dec_groups;
  /* Deterministic accumulation (fix_scale > 0, also in drasic_tail_bwd_args / drasic_head_bwd_args): the weight-gradient
   * destinations (dw_e4, dw_d0; dw_d4; dw_e0, db_e0) are then int64 buffers of the same shapes and every block adds
   * round(partial * fix_scale) -- integer addition is associative, so the sums do not depend on the order in which the
   * blocks arrive (float atomics, the default, differ from run to run at the 1e-6 level); the caller divides by
   * fix_scale afterwards.  Choose fix_scale so that |sum| * fix_scale < 2^62, e.g. 2^30 / loss_scale. */
  double fix_scale;
} drasic_bottleneck_bwd_args;
int drasic_bottleneck_bwd(const drasic_bottleneck_bwd_args* a, void* stream);

/* backward of the decoder-tail half of drasic_tail_head_fwd for iteration t; maintains the running
 * gradient w.r.t. the cumulative reconstruction (gacc) */
typedef struct {
  const float* d3_next; const float* w_d4; const float* img; const float* recon_t; float* gacc;
  float* dh_d3; float* dw_d4; const int* perm; const int* group_img_end;
  float loss_scale;
  int B_pad, C, H, W, first, loss_kind, n_dec_groups, has_acc;
  const int* group_iters;         /* nullable, as in drasic_tail_head_args: frozen nodes pass no gradient to the decoder */
  int iter, n_enc_groups;
  double fix_scale;               /* see drasic_bottleneck_bwd_args */
} drasic_tail_bwd_args;
int drasic_tail_bwd(const drasic_tail_bwd_args* a, void* stream);

/* backward of the encoder-head half for iteration t */
typedef struct {
  const float* du1; const float* w_e0; const float* b_e0; const float* img; const float* recon_prev;
  float* gacc; float* dw_e0; float* db_e0; const int* perm; const int* group_img_end;
  int B_pad, C, H, W, first, propagate, n_enc_groups;
  double fix_scale;               /* see drasic_bottleneck_bwd_args */
} drasic_head_bwd_args;
int drasic_head_bwd(const drasic_head_bwd_args* a, void* stream);

/*
 * Optimizer step (train_model.py:116 `optimizer.step()` on the optim.Adam of train_model.py:155-161): Adam with
 * L2 weight decay folded into the gradient, no amsgrad, over any number of fp32 tensors in ceil(n/64) launches.
 * `tensors` is a HOST array (the device pointers travel as kernel arguments); param / exp_avg / exp_avg_sq are
 * updated in place.  bias_correction{1,2} = 1 - beta{1,2}^step for the step being taken (computed by the caller
 * in double, like torch; the hyper-parameters are doubles for the same reason: 1 - beta is formed in double and
 * rounded to fp32 once).  All tensors of one call share the step count.
 */
typedef struct {
  float* param; const float* grad; float* exp_avg; float* exp_avg_sq;
  long long numel;
} drasic_adam_tensor;
int drasic_adam_step(const drasic_adam_tensor* tensors, int n_tensors, double lr, double beta1, double beta2,
                     double eps, double weight_decay, double bias_correction1, double bias_correction2,
                     void* stream);

/*
 * Stand-alone cells (called outside the fused loop; NCHW fp32 like the reference).
 * drasic_conv1x1_fwd: ConvCell.forward with kernel_size 1 (modules/cell.py:69-72): y = act(w*x + b);
 *   x [N,Cin,HW], w [Cout,Cin], b nullable [Cout], y [N,Cout,HW].
 * drasic_quantize_fwd: Quantize.forward (functions/quantize.py:9-18); u nullable = eval mode.
 */
enum { DRASIC_ACT_NONE = 0, DRASIC_ACT_TANH = 1, DRASIC_ACT_RELU = 2, DRASIC_ACT_SIGMOID = 3,
       DRASIC_ACT_HARDTANH = 4, DRASIC_ACT_ELU = 5, DRASIC_ACT_SELU = 6, DRASIC_ACT_CELU = 7 };
int drasic_conv1x1_fwd(const float* x, const float* w, const float* b, float* y, int N, int Cin,
                       int Cout, int HW, int act, void* stream);
int drasic_quantize_fwd(const float* x, const float* u, float* y, size_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DRASIC_B200_H_ */
