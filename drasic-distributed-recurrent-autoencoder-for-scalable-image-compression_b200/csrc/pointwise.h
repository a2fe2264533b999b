// Parameter blocks of the fused HBM-bound kernels (pointwise.cu) and of the weight packer (pack.cu).
#pragma once
#include "kernels.h"

namespace drasic {

enum ActKind : int { ACT_NONE = 0, ACT_TANH = 1, ACT_RELU = 2, ACT_SIGMOID = 3, ACT_HARDTANH = 4, ACT_ELU = 5, ACT_SELU = 6, ACT_CELU = 7 };  // modules/cell.py:23-46
enum LossKind : int { LOSS_MAE = 0, LOSS_MSE = 1, LOSS_BCE = 2 };  // dist_codec.py:24-34
enum TailMode : int {
  TAIL_INIT = 0,   // before iteration 0: x = img*2-1 -> encoder head only
  TAIL_FIRST = 1,  // iteration 0: decoded = (decoded+1)/2
  TAIL_STEP = 2,   // iterations >= 1
};

struct BottleneckParams {
  const float* h_e3;   // [B_pad, HW, Ch] fp32: last encoder ConvLSTM h (NHWC, identity order)
  const float* w_e4;   // [n_enc_groups][code, Ch]
  const float* w_d0;   // [n_dec_groups][Ch, code]
  const float* noise;  // nullable (eval): [B, code, HW] uniform [0,1), original batch order
  const int* perm;     // [B_pad] slot -> original image index, -1 = padding
  float* code_out;     // [B, code, HW] +-1 (NCHW, original order)
  float* z_out;        // nullable: pre-quantizer activations, same layout
  uint8_t* bits_out;   // nullable: [B, code*HW/8] packed sign bits
  __half* d1_in[2];    // [B_pad, HW, Ch] fp16 {hi, lo}: first decoder ConvLSTM input
  float* d0_out_f32;   // nullable: fp32 copy of the decoder head output (saved for backward)
  int B_pad, HW, Ch, code;
  int n_enc_groups, n_dec_groups;
  int precise, split;
  const int* group_img_end;  // device [n_groups]: exclusive end slot of each node
  const uint8_t* bits_in;    // nullable: decode-only, symbols come from this packed stream [B, code*HW/8]
};

struct TailParams {
  const float* d3_next;     // [B_pad, H/2, W/2, 4*32] fp32: last decoder ConvLSTM h in its own (quad-major) order
  const float* w_d4;        // [n_dec_groups][C, 32]
  const float* w_e0;        // [n_enc_groups][32, C]
  const float* b_e0;        // nullable: [n_enc_groups][32]
  const float* img;         // [B, C, H, W] original order
  const float* recon_prev;  // [B, C, H, W]; unused when mode != TAIL_STEP
  float* recon_out;         // [B, C, H, W]
  float* dec_out;           // nullable: raw tanh decoder output (saved for backward)
  double* loss_acc;         // [2]: {sum of loss terms, sum of squared error}
  __half* e1_in[2];         // [B_pad, H/2, W/2, 128] fp16 {hi, lo}; null = no next encoder pass
  const int* perm;
  int B_pad, C, H, W;
  int mode, loss_kind;
  int n_enc_groups, n_dec_groups;
  int precise, split;
  const int* group_img_end;  // device [n_groups]: exclusive end slot of each node
  const int* group_iters;    // nullable: per-node iteration budget (heterogeneous rates)
  int iter;
};

int launch_bottleneck(const BottleneckParams& p, cudaStream_t stream);
int launch_tail_head(const TailParams& p, cudaStream_t stream);
int pixel_blocks_per_image(int B_pad, int HW, int pix_per_pass, int min_blocks);
int launch_conv1x1(const float* x, const float* w, const float* b, float* y, int N, int Cin, int Cout,
                   int HW, int act, cudaStream_t stream);
int launch_quantize(const float* x, const float* u, float* y, size_t n, cudaStream_t stream);
int launch_pack_lstm_weights(const PackBatch& b, int n, int n_slots, int Cin, int Co, int in_order, int out_order, int block_n,
                             __half* out_hi, __half* out_lo, float* inv_scale_out, unsigned int* scratch_amax,
                             cudaStream_t stream);

}  // namespace drasic
