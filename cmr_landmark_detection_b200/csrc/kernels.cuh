// Launchers of the CUDA-core kernels (memory-bound passes, CUDA-core convs, head/loss, extraction, Adam).
#pragma once
#include "common.cuh"
#include "conv_tc.cuh"

namespace rvip {

// ------------------------------------------------------------------ CUDA-core convs (conv_simt.cu)
struct ConvSimtArgs {
  const void* in0;
  const void* in1;      // second concat source (channels C0..Ctot) or nullptr
  const float* w;       // [9][Ctot][Cout] fp32 (Keras HWIO; rotated copy for dgrad)
  const float* bias;
  void* out0;
  void* out1;           // EPI_LINEAR: channels >= out_split
  double* stats;        // [2][Cout]
  int B, H, W, C0, Ctot, Cout, mode, out_split;
  float floor = 0.f;    // activation floor of the EPI_RELU* modes: 0 = ReLU, -inf = none (BN_FIRST)
  // Conv2DTranspose(3, strides 2, 'same') in fp32 parity mode (KerasLayers.py:762-765) as a plain 3x3 convolution over the
  // VIRTUAL zero-stuffed input xs[2i + 1][2j + 1] = x[i][j] (zero elsewhere) with the flipped kernel:
  //   out[o] = sum_{2i + k = o} x[i] w[k]  ==  sum_{k'} xs[o + k' - 1] w[2 - k'].
  int in_stuffed = 0;   // in0 is the LOW-resolution tensor [B, H/2, W/2, C0]; H, W are the output (high) resolution
  int out_stuffed = 0;  // EPI_LINEAR: only the odd (y, x) outputs are kept, written to a low-resolution [B, H/2, W/2, .] tensor
                        // (gradient of the zero-stuffing = gather)
};
int conv_simt_launch(const ConvSimtArgs& a, int in_is_bf16, int out_is_bf16, cudaStream_t st);

struct WgradSimtArgs {
  const void* in0;
  const void* in1;
  const void* dz;
  float* dw;            // [9][Ctot][Cout] fp32, accumulated
  int B, H, W, C0, Ctot, Cout;
  int in_stuffed = 0;   // in0 is low-resolution and virtually zero-stuffed (see ConvSimtArgs)
  int transposed_out = 0;   // accumulate into a Conv2DTranspose kernel gradient (kh, kw, Cout, Ctot), taps flipped
};
int wgrad_simt_launch(const WgradSimtArgs& a, int in_is_bf16, int dz_is_bf16, cudaStream_t st);

// Cin == 1 first layer (x fp32 [B,H,W], w [9][Cout] fp32)
// scale / shift != nullptr: y = scale * relu(conv + bias) + shift (inference with BatchNorm folded in)
int conv_c1_fwd_launch(const float* x, const float* w, const float* bias, void* out, double* stats, int B, int H, int W,
                       int Cout, int want_stats, int out_is_bf16, const float* scale, const float* shift,
                       cudaStream_t st);
int wgrad_c1_launch(const float* x, const void* dz, float* dw, int B, int H, int W, int Cout, int dz_is_bf16,
                    cudaStream_t st);

// ------------------------------------------------------------------ BatchNorm passes (bn.cu)
constexpr int kRedStripes = kBnRedStripes;   // copies of the BN-backward partial sums (blocks add to copy blockIdx % 4)
enum PostOp { POST_NONE = 0, POST_DROPOUT = 1, POST_POOL = 2, POST_UPSAMPLE = 3 };

struct BnArgs {
  int B, H, W, C;          // geometry of the conv output `a`
  int post;
  const void* a;           // relu(conv) [B,H,W,C]
  const float* gamma;
  const float* beta;
  const float* mean;       // batch mean (training) or moving mean (inference)
  const float* rstd;
  // training forward: batch statistics straight from the conv epilogue (nullptr = inference, use mean / rstd)
  const double* stats;     // [2][C] sum, sum of squares
  double count, inv_count; // B*H*W
  float momentum, eps;
  float* mean_out;         // [C] published by block 0 for the backward pass
  float* rstd_out;
  float* mov_mean;
  float* mov_var;
  // forward outputs
  void* y;                 // [B,H,W,C]        (NONE / DROPOUT / POOL)
  void* y2;                // POOL: [B,H/2,W/2,C]; UPSAMPLE: [B,2H,2W,C]
  // dropout
  uint64_t seed;
  uint32_t site, thr16;
  float keep_scale;        // 1 / (1 - rate)
  // backward inputs
  const void* g0;          // NONE/DROPOUT: dL/d(y after dropout) [B,H,W,C]; POOL: d skip; UPSAMPLE: d(up) [B,2H,2W,C]
  const void* g1;          // POOL: d pooled [B,H/2,W/2,C]
  double* red;             // [kRedStripes][2][C]: sum dy, sum dy*a (striped partial sums)
  void* dz;                // [B,H,W,C]
  float* dgamma;
  float* dbeta;
  float* dbias;            // conv bias gradient = sum dz
  int identity;            // scale 1 / shift 0: the affine already happened in the conv epilogue (inference), or the block
                           // has no BatchNorm at all (BATCH_NORMALISATION false: backward is then the ReLU mask alone)
  int bn_first;            // BN_FIRST (KerasLayers.py:681-685): `a` holds z = conv + bias, y = relu(BN(z)); backward masks
                           // dy with [BN(z) > 0] and dz is NOT masked by [a > 0]
  // first layer (Cin = 1, training): `a` = relu(conv(x) + b) is never stored -- K = 9, so every pass that needs it
  // recomputes it from the 1-channel image (saves a 134 MB write and three 134 MB reads per step at C2)
  const float* x0;         // [B,H,W] fp32 image, nullptr = read `a`
  const float* w0;         // [9][C] fp32 (Keras HWIO with Cin = 1)
  const float* b0;         // [C]
};
// first layer, training: per-channel sum / sum of squares of relu(conv(x) + b) without storing it
int c1_stats_launch(const float* x, const float* w, const float* bias, double* stats, int B, int H, int W, int C,
                    cudaStream_t st);
int bn_eval_prepare_launch(const float* mov_mean, const float* mov_var, float* mean, float* rstd, int n, float eps,
                           cudaStream_t st);
// inference: scale = gamma / sqrt(moving_var + eps), shift = beta - moving_mean * scale for EVERY BatchNorm layer in
// one launch (table entry per layer: offsets into params / bn_state / the per-channel coefficient arrays)
struct BnEvalEntry {
  long long off_g, off_be, off_mm, off_mv, off_stat;
  int C;
};
int bn_eval_coef_launch(const float* params, const float* bn_state, const BnEvalEntry* table_dev, int n_layers, int max_c,
                        float eps, float* scale, float* shift, cudaStream_t st);
int bn_apply_launch(const BnArgs& a, int is_bf16, cudaStream_t st);
int bn_bwd_reduce_launch(const BnArgs& a, int is_bf16, cudaStream_t st);
int bn_bwd_apply_launch(const BnArgs& a, int is_bf16, cudaStream_t st);
// conv + ReLU without BN (decoder up-conv): dz = du * [u > 0], dbias = sum dz
int relu_bwd_launch(const void* u, const void* du, void* dz, float* dbias, size_t pixels, int C, int is_bf16,
                    cudaStream_t st);
int dropout_mask_launch(uint64_t seed, uint32_t site, uint32_t thr16, size_t n_vec8, uint8_t* keep, cudaStream_t st);

// ------------------------------------------------------------------ head + loss (head_loss.cu)
enum LossKind { LOSS_MSE = 0, LOSS_MASKED = 1, LOSS_WEIGHTED = 2, LOSS_BCE_DICE = 3 };
struct HeadArgs {
  int B, H, W, Cin, NC;
  const void* y;           // [B,H,W,Cin]
  const float* w;          // [Cin][NC]
  const float* b;          // [NC]
  float* heat;             // [B,H,W,NC] fp32 sigmoid output
  // training
  const float* target;     // [B,H,W,NC]
  const float* inplane;    // [H,W] or nullptr
  int loss_kind;
  float mask_thr, eps;
  void* dy;                // [B,H,W,Cin]
  float* dw;
  float* db;
  double* loss_acc;        // scalar accumulator (sum of per-pixel losses)
  // LOSS_BCE_DICE (Loss_and_metrics.py:208-245): w_bce * BCE - w_dice * Dice; the Dice term needs the batch-global
  // sums {sum t*p, sum p, sum t} before any gradient can be formed (head_dice_sums_launch)
  const double* dice_sums;
  float w_bce, w_dice;
  // BatchNorm of the last decoder block folded into the head (training): `y` is then that block's relu(conv) output
  // `a`, the head derives scale / shift from the conv epilogue's sum / sum^2 (bn_stats != nullptr) and normalises on the
  // fly -- the block's bn_apply pass (read a, write y) disappears.  dw then accumulates sum_p a * dlogit, which
  // head_bn_finalize_launch turns into the true head-kernel gradient AND into the BatchNorm-backward sums of that block
  // (sum dy, sum dy * a): its bn_bwd_reduce pass disappears as well.
  const double* bn_stats;  // [2][Cin] sum, sum of squares of `a`
  double bn_count, bn_inv_count;
  float bn_momentum, bn_eps;
  const float* bn_gamma;
  const float* bn_beta;
  float* bn_mean_out;      // [Cin] published by block 0 when bn_publish (for the backward pass)
  float* bn_rstd_out;
  float* bn_mov_mean;
  float* bn_mov_var;
  int bn_publish;
};
int head_launch(const HeadArgs& a, int training, int is_bf16, cudaStream_t st);
// dwa [Cin][NC] = sum_p a * dlogit, db [NC] = sum_p dlogit (both from the folded training head), w the head kernel:
//   dw[c][k]  = sc[c] * dwa[c][k] + sh[c] * db[k]          (y = sc * a + sh)
//   red[c]    = sum_p dy[c]     = sum_k w[c][k] * db[k]
//   red[C+c]  = sum_p dy[c] a   = sum_k w[c][k] * dwa[c][k]      (dy = dlogit . w^T: exact, a 1x1 conv has no border)
int head_bn_finalize_launch(const float* dwa, const float* db, const float* w, const float* gamma, const float* beta,
                            const float* mean, const float* rstd, int Cin, int NC, float* dw, double* red,
                            cudaStream_t st);
// heat, target [n] fp32 -> sums[3] += {sum t*p, sum p, sum t} (double)
int head_dice_sums_launch(const float* heat, const float* target, size_t n, double* sums, cudaStream_t st);

// heat, target [n_pixels][NC] fp32 -> out[0] = sum of the per-pixel loss terms, out[1 + 3c ..] = {sum t*p, sum p, sum t}
// of channel c (double; zeroed by the launch)
int heat_stats_launch(const float* heat, const float* target, const float* inplane, size_t n_pixels, int HW, int NC,
                      int loss_kind, float mask_thr, double* out, cudaStream_t st);

// ------------------------------------------------------------------ landmark extraction (extract.cu)
int extract_launch(const float* heat, int Z, int H, int W, int C, float thr, double* yx, int* count, int* argmax,
                   float* maxv, unsigned long long* scratch, cudaStream_t st);
size_t extract_scratch_bytes(int Z, int C);
int label_map_launch(const float* heat, size_t n_pix, int C, float thr, uint8_t* out, cudaStream_t st);
// gt / pred [Z][2 landmarks][y, x] float64 (NaN = missing) -> angle [2][Z], dist / dist_thr / dist_ub [2][Z], summary [18]
int landmark_metrics_launch(const double* gt, const double* pred, int Z, double spacing, double thr, double dim,
                            double* angle, double* dist, double* dist_thr, double* dist_ub, double* summary,
                            cudaStream_t st);

// ------------------------------------------------------------------ largest-connected-component filter (cc.cu)
size_t cc_scratch_bytes(int Z, int H, int W);
int cc_filter_launch(const uint8_t* labels, int Z, int H, int W, int connectivity, uint8_t* out, void* scratch,
                     cudaStream_t st);

// ------------------------------------------------------------------ optimizer / weight packing (optim.cu)
int adam_launch(float* p, const float* g, float* m, float* v, size_t n, float lr_t, float b1, float b2, float eps,
                float grad_scale, cudaStream_t st);
int sgd_launch(float* p, const float* g, float* velocity, size_t n, float lr, float momentum, int nesterov,
               float grad_scale, cudaStream_t st);
struct PackEntry {
  long long src;      // float offset of the HWIO kernel in the parameter buffer
  long long dst_f;    // element offset in the packed buffer: forward  [Cout][9][Ctot]
  long long dst_d;    // element offset in the packed buffer: dgrad    [Ctot][9][Cout] (taps rotated), -1 = none
  int Ctot, Cout;
};
// bf16 packs for the tensor-core kernels (packed = __nv_bfloat16*) or fp32 rotated copies for the
// CUDA-core dgrad (packed = float*, only dst_d is written, layout [9][Cout][Ctot] rotated)
int pack_weights_launch(const float* params, void* packed, const PackEntry* table_dev, int n_entries, int to_bf16,
                        cudaStream_t st);

// phase-decomposed up-convolution operands (conv_halo.cuh): C == 32 -> ns = 3, C % 64 == 0 -> ns = 2
struct UpPackEntry {
  long long src;      // float offset of the HWIO kernel [3][3][Cin][C]
  long long dst_f;    // element offset of the forward copy
  long long dst_d;    // element offset of the dgrad copy, -1 = none
  int Cin, C, ns;
  int transposed;     // source is a Conv2DTranspose kernel (kh, kw, C, Cin): every phase weight is ONE tap (or zero)
};
int pack_up_launch(const float* params, void* packed, const UpPackEntry* table_dev, int n_entries, cudaStream_t st);

}  // namespace rvip
