// Parameter blocks of the backward-through-time kernels (backward.cu, wgrad.cu).  Together they
// replace what autograd runs for `output['loss'].backward()` (reference train_model.py:115) over the
// loop of models/dist_codec.py:36-65.
#pragma once
#include "kernels.h"

namespace drasic {

// LSTM cell backward (modules/cell.py:133-139 differentiated), one launch per (layer, t): writes the scaled fp16 {hi, lo} operand planes of the dgrad / wgrad GEMMs
// directly (no fp32 gate-gradient round trip).  The per-tensor power-of-two scale is *predicted* from the
// amax of the previously processed iteration; the kernel also records the true amax.  A second launch
// with verify != 0 re-does the tensor with the exact scale only if the prediction left the scaled amax
// outside [2^4, 2^15.9] (it returns immediately otherwise).  Both launches read the same inputs.
struct LstmBwdFusedParams {
  const void* gates; int gates_f32;
  const float* c_t; const float* c_prev;   // c_prev nullable (t == 0)
  const float* dh_above; const float* dh_rec;   // dh_rec nullable (t == T-1)
  const float* dc_in;                      // nullable (t == T-1): dL/dc_t carried from step t+1
  float* dc_out;                           // dL/dc_{t-1} (a different buffer than dc_in)
  __half* hi; __half* lo;                  // [P, 4, Co] scaled gate gradients; lo nullable
  const unsigned int* prev_amax_bits;      // amax (float bits) of the previously processed iteration; 0 = unknown
  unsigned int* amax_bits;                 // out: true amax of this tensor (zero before the first launch)
  float* inv_scale_out;                    // out: 1 / scale actually in the planes
  size_t n_pix; int Co; int verify;
};

struct WgradParams {
  CUtensorMap tm_g[2];      // dG planes [B,H,W,4Co] fp16 {hi, lo}, box = 64 ch x 64 pixels
  CUtensorMap tm_x[2];      // layer input planes [B,H,W,Cin], same pixel box (shifted by the tap)
  CUtensorMap tm_h[2];      // h_{t-1} planes [B,H,W,Co]
  float* gw;                // [groups][4Co][9(Cin+Co)] fp32 packed-order weight gradient (atomic accumulate)
  const float* g_inv_scale; // device scalar: 1/scale of the dG planes
  const int* group_kb_end;  // device [n_groups]: exclusive end k-block (64 pixels) per node; null = 1 group
  int H, W, Cin, Co;
  int kbw, kbh, kbn;        // pixel box of one K block: kbw*kbh*kbn == 64
  int n_kb;                 // total K blocks = B_pad*H*W/64
  int has_h;                // 0 at t == 0
  int n_groups, m_tiles, n_tiles_x, n_tiles_h, k_splits;
  int wide;                 // CTA pairs only: 512-column tiles (n_tiles_* count units of 8 instead of 4)
  int dbg;                  // -DDRASIC_DIAG builds only (env DRASIC_WGRAD_DBG): 1 = skip the TMA loads, 2 = skip the MMAs
};

struct BottleneckBwdParams {
  const float* dd0;        // [B_pad, HW, Ch] gradient w.r.t. the decoder-head output (from D1 dgrad)
  const float* h_e3;       // [B_pad, HW, Ch] saved encoder ConvLSTM output of this iteration
  const float* z;          // [B, code, HW] saved pre-quantizer activations (original order)
  const float* codes;      // [B, code, HW] emitted symbols (original order)
  const float* w_e4;       // [n_enc_groups][code, Ch]
  const float* w_d0;       // [n_dec_groups][Ch, code]
  const int* perm;
  const int* group_img_end;
  float* dh_e3;            // [B_pad, HW, Ch] gradient w.r.t. the encoder ConvLSTM output
  float* dw_e4;            // [n_enc_groups][code, Ch] accumulated
  float* dw_d0;            // [n_dec_groups][Ch, code] accumulated
  int B_pad, HW, Ch, code, n_enc_groups, n_dec_groups;
  double fix_scale;        // > 0: deterministic accumulation, dw_e4 / dw_d0 are int64 fixed-point buffers (value * fix_scale)
};

struct TailBwdParams {
  const float* d3_next;    // [B_pad,H/2,W/2,128] saved decoder ConvLSTM output of iteration t (own quad-major order)
  const float* w_d4;       // [n_dec_groups][C,32]
  const float* img;        // [B,C,H,W]
  const float* recon_t;    // [B,C,H,W] reconstruction after iteration t
  float* gacc;             // [B,C,H,W] running dL/d(recon) accumulator (see backward.cu)
  float* dh_d3;            // [B_pad,H/2,W/2,128] gradient w.r.t. D3's h (own QUAD layout)
  float* dw_d4;            // [n_dec_groups][C,32] accumulated
  const int* perm;
  const int* group_img_end;
  float loss_scale;        // 1 / (T * B*C*H*W)
  int B_pad, C, H, W, first, loss_kind, n_dec_groups, has_acc;
  const int* group_iters;  // nullable: per-node iteration budget (heterogeneous rates)
  int iter, n_enc_groups;
  double fix_scale;        // > 0: deterministic accumulation, dw_d4 is an int64 fixed-point buffer
};

struct HeadBwdParams {
  const float* du1;        // [B_pad,H/2,W/2,128] gradient w.r.t. the first encoder ConvLSTM input
  const float* w_e0;       // [n_enc_groups][32,C]
  const float* b_e0;       // nullable
  const float* img;
  const float* recon_prev; // nullable (t == 0: x = img*2-1)
  float* gacc;             // [B,C,H,W]: -= dL/dx_t when propagate != 0
  float* dw_e0;            // [n_enc_groups][32,C]
  float* db_e0;            // nullable [n_enc_groups][32]
  const int* perm;
  const int* group_img_end;
  int B_pad, C, H, W, first, propagate, n_enc_groups;
  double fix_scale;        // > 0: deterministic accumulation, dw_e0 / db_e0 are int64 fixed-point buffers
};

int launch_lstm_bwd_fused(const LstmBwdFusedParams& p, cudaStream_t stream);
int launch_pack_dgrad_weights(const PackBatch& b, int n, int n_slots, int Cin, int Co, int in_order, int out_order,
                              __half* out_hi, __half* out_lo, float* inv_scale_out, unsigned int* scratch_amax,
                              cudaStream_t stream);
int launch_unpack_wgrad(const float* gw, int Cin, int Co, int in_order, int out_order, float* dw_in,
                        float* dw_h, cudaStream_t stream);
int launch_wgrad(const WgradParams& p, bool split, bool pair, int num_sms, cudaStream_t stream);
int launch_bottleneck_bwd(const BottleneckBwdParams& p, cudaStream_t stream);
int launch_tail_bwd(const TailBwdParams& p, cudaStream_t stream);
int launch_head_bwd(const HeadBwdParams& p, cudaStream_t stream);

}  // namespace drasic
