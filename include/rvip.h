/* rvip.h -- C ABI of the B200-native RVIP heat-map U-Net hot path (librvip_b200.so).
 *
 * The reference (Cardio-AI/cmr-landmark-detection) has no FFI layer of its own: its hot path is
 * reached through the tf.keras object returned by create_unet().  Each entry point below names
 * the reference interface it stands in for; INTEGRATION.md shows the ctypes binding a maintainer
 * adds behind src/models/Unets.py:create_unet.
 *
 * Conventions: all pointers are DEVICE pointers unless the name ends in _host; tensors are NHWC;
 * every call is asynchronous on the caller's cudaStream_t (passed as void*); functions return 0
 * on success, non-zero on failure with a message in rvip_last_error().  A handle is owned by one
 * host thread and one GPU.  There is no CPU fallback anywhere in this library.
 */
#ifndef RVIP_H_
#define RVIP_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RVIP_MAX_DEPTH 8

/* Network description == the config keys create_unet() reads (src/models/Unets.py:77-106). */
typedef struct rvip_cfg {
  int H, W;              /* DIM */
  int in_ch;             /* IMG_CHANNELS */
  int classes;           /* MASK_CLASSES */
  int depth;             /* DEPTH */
  int filters;           /* FILTERS */
  int batch_norm;        /* BATCH_NORMALISATION: 0 = the blocks are Conv -> ReLU only (KerasLayers.py:684,691 skipped) */
  int bn_first;          /* BN_FIRST: 0 = Conv -> ReLU -> BN (KerasLayers.py:687-691, every shipped config),
                            1 = Conv -> BN -> ReLU (:681-685) */
  int use_upsample;      /* USE_UPSAMPLE truthiness: 1 = UpSampling2D + Conv (KerasLayers.py:753-759), 0 = Conv2DTranspose
                            (:762-765; bf16: shapes that fit the phase-decomposed kernels; fp32: CUDA-core convolution
                            over the zero-stuffed input) */
  int precision;         /* 0 = fp32 storage, CUDA-core convs; 1 = bf16 storage, tcgen05 convs */
  float dropout[RVIP_MAX_DEPTH]; /* encoder level l; decoder pops from the back (Unets.py:105-106, :832) */
  float dropout_mid;     /* DROPOUT_MAX at the bottleneck (Unets.py:813) */
  float bn_momentum;     /* 0.99  (Keras BatchNormalization default) */
  float bn_eps;          /* 1e-3 */
} rvip_cfg;

typedef struct rvip_handle rvip_handle;

/* RVIP_LOSS_BCE_DICE: w_bce * binary_crossentropy - w_dice * dice_coef (src/models/Loss_and_metrics.py:165-171,
 * 208-245; the reference's default LOSS_FUNCTION 'BcdDiceLoss'), weights via rvip_set_loss_weights (default 1, 1). */
enum { RVIP_LOSS_MSE = 0, RVIP_LOSS_MASKED = 1, RVIP_LOSS_WEIGHTED = 2, RVIP_LOSS_BCE_DICE = 3 };

const char* rvip_last_error(void);
int rvip_abi_version(void);

/* ---- model construction: replaces unet()/create_unet() graph building (Unets.py:61-133, 755-869) */
int rvip_create(const rvip_cfg* cfg, rvip_handle** out);
void rvip_destroy(rvip_handle* h);

/* Flat buffers: `params`/`grads` hold the trainable tensors, `bn_state` the BN moving statistics.
 * The tensor table lists model.get_weights() order (Conv: kernel HWIO, bias; BN: gamma, beta,
 * moving_mean, moving_variance) with the offset of each tensor in its flat buffer. */
long long rvip_param_count(const rvip_handle* h);
long long rvip_state_count(const rvip_handle* h);
int rvip_num_tensors(const rvip_handle* h);
int rvip_tensor_info(const rvip_handle* h, int index, char* name, int name_cap, int* is_state, long long* offset,
                     int* ndim, int dims[4]);

/* Workspace (activations, gradients of activations, packed weights, tensor maps are host side). */
size_t rvip_workspace_bytes(const rvip_handle* h, int batch, int training);
int rvip_bind(rvip_handle* h, float* params, float* grads, float* bn_state, void* workspace, size_t workspace_bytes,
              int batch, int training);
/* Re-derive the tensor-core operand copies after `params` changed (set_weights / load_weights). */
int rvip_pack_weights(rvip_handle* h, void* stream);

/* ---- model.predict (predict_model.py:143): inference forward, BN moving stats, no dropout.
 * x [B,H,W,in_ch] fp32 -> heat [B,H,W,classes] fp32 */
int rvip_predict(rvip_handle* h, const float* x, float* heat, void* stream);

/* ---- one model.fit step minus the optimizer (train_model.py:105): forward with batch statistics and
 * dropout(seed), loss, full backward.  grads (bound buffer) receive d(mean loss)/d(param);
 * loss_out is a device double.  inplane [H,W] is only read for RVIP_LOSS_WEIGHTED. */
int rvip_train_step(rvip_handle* h, const float* x, const float* target, const float* inplane, int loss_kind,
                    float mask_thr, uint64_t seed, float* heat, double* loss_out, void* stream);

int rvip_set_loss_weights(rvip_handle* h, float w_bce, float w_dice);

/* ---- validation loss and Dice metrics (model.evaluate / fit(validation_data=...), train_model.py:54-59, 105-112;
 * src/models/Loss_and_metrics.py:124-171): one pass over heat and target [n_pixels][classes] fp32.
 * out (device, 1 + 3 * classes doubles, zeroed by the call): out[0] = sum over pixels of the compiled loss's per-pixel
 * term (MSE: mean_c (p - t)^2; masked / weighted: loss_with_zero_mask; BCE+Dice: mean_c binary cross-entropy),
 * out[1 + 3c ..] = {sum t*p, sum p, sum t} of channel c, from which every dice_coef* metric follows.  hw = H * W
 * (period of the in-plane weights, only read for RVIP_LOSS_WEIGHTED). */
int rvip_heat_stats(const float* heat, const float* target, const float* inplane, long long n_pixels, int hw, int classes,
                    int loss_kind, float mask_thr, double* out, void* stream);

/* ---- Adam apply (ModelUtils.py:107; Keras epsilon-hat form) + operand re-pack.
 * grad_scale folds the data-parallel 1/world into the update. step counts from 1. */
int rvip_adam_step(rvip_handle* h, float* m, float* v, float lr, float beta1, float beta2, float eps, long long step,
                   float grad_scale, void* stream);
/* Arms the NEXT rvip_train_step to apply Adam itself (same update as rvip_adam_step): every gradient bucket is stepped and
 * its tensor-core operand copies re-packed on the weight-gradient stream as soon as the bucket's gradients are final, so
 * the optimizer hides behind the rest of the backward pass.  Single-replica training only (under data parallelism the
 * all-reduce sits between backward and the optimizer: use rvip_adam_step).  One-shot: cleared by that train step. */
int rvip_set_inline_adam(rvip_handle* h, float* m, float* v, float lr, float beta1, float beta2, float eps, long long step,
                         float grad_scale);
/* Data-parallel counterpart: Adam + operand re-pack of ONE gradient bucket (rvip_bucket) on the caller's stream -- issued on
 * the communication stream right behind that bucket's all-reduce, so the optimizer of early buckets overlaps the backward
 * pass and the all-reduce of later ones. */
int rvip_adam_bucket(rvip_handle* h, int bucket, float* m, float* v, float lr, float beta1, float beta2, float eps,
                     long long step, float grad_scale, void* stream);
/* tf.keras.optimizers.SGD apply (OPTIMIZER='sgd', ModelUtils.py:109-111, and the Adam -> SGD switch of
 * utils/KerasCallbacks.py:280-306) + operand re-pack: v = momentum v - lr g; w += nesterov ? momentum v - lr g : v.
 * velocity [n_params] may be NULL when momentum == 0. */
int rvip_sgd_step(rvip_handle* h, float* velocity, float lr, float momentum, int nesterov, float grad_scale, void* stream);


/* ---- data-parallel plumbing (MirroredStrategy, Unets.py:70-75): gradient buckets are contiguous
 * ranges of `grads` in backward-completion order; each records a cudaEvent_t when complete so the
 * caller can start that bucket's all-reduce while the rest of backward still runs.  The event is
 * recorded on the handle's weight-gradient stream once every gradient of the bucket is final (the
 * main chain is never made to wait for a weight gradient at a bucket boundary); the step's own
 * stream has joined that stream by the time rvip_train_step's launches end. */
int rvip_num_buckets(const rvip_handle* h);
int rvip_bucket(const rvip_handle* h, int index, long long* offset, long long* count);
int rvip_set_bucket_event(rvip_handle* h, int index, void* cuda_event);

/* ---- landmark extraction: threshold/label map (predict_model.py:153-156) + per-slice centroid
 * (evaluate_cv.py:418-442) + argmax/max.
 * heat [Z,H,W,C] fp32 -> yx [Z,C,2] float64 (NaN when the label is absent), count [Z,C] int32,
 * argmax [Z,C] int32 flat index, maxv [Z,C] fp32.  scratch: rvip_extract_scratch_bytes(Z,C) bytes. */
size_t rvip_extract_scratch_bytes(int Z, int C);
int rvip_extract(const float* heat, int Z, int H, int W, int C, float thr, double* yx, int* count, int* argmax,
                 float* maxv, void* scratch, void* stream);

/* threshold -> uint8 label map exactly as predict_model.py:153-156 (0 background, c+1 for the LAST channel
 * whose value is > thr); heat [n_pixels, C] fp32 -> labels [n_pixels] */
int rvip_label_map(const float* heat, long long n_pixels, int C, float thr, uint8_t* labels, void* stream);

/* ---- largest-connected-component filter (src/data/Postprocess.py:108-120 clean_3d_prediction_2d_cc; CC_FILTER in
 * predict_model.py:159-161): labels [Z,H,W] uint8 (0 = background, values 1..4) -> out, only the largest component of
 * each value per slice. connectivity 8 is what the reference's cv2 call actually runs (its positional 4 lands in the
 * `labels` slot); 4 is offered too. Equal maximal areas: the component whose first pixel comes first in raster order.
 * scratch: rvip_cc_scratch_bytes(Z,H,W) bytes (8-byte aligned). */
size_t rvip_cc_scratch_bytes(int Z, int H, int W);
int rvip_cc_filter(const uint8_t* labels, int Z, int H, int W, int connectivity, uint8_t* out, void* scratch, void* stream);

/* ---- per-volume landmark metrics, the step after extraction (src/models/evaluate_cv.py): get_angle2x :508-536,
 * get_distances :549-561, get_distances_upper_bound :572-595, calc_mean_ip :113-120, calc_tpr_thresh :267-308,
 * calc_ppv_thresh :311-353.  gt_yx / pred_yx [Z][2 landmarks: anterior, inferior][y, x] float64, NaN = missing.
 * Outputs (device, float64): angle [2: gt, pred][Z] degrees in [0,360) or NaN; dist / dist_thr / dist_ub
 * [2 landmarks][Z] (spacing applied; dist_thr NaN above `threshold`; dist_ub = distance to the farthest corner of the
 * dim x dim image when the prediction is missing); summary [18] = mean points [gt,pred][ant,inf][y,x] (8),
 * tpr[2], ppv[2], counters [landmark][tp, fn, fp] (6). */
int rvip_landmark_metrics(const double* gt_yx, const double* pred_yx, int Z, double spacing, double threshold, double dim,
                          double* angle, double* dist, double* dist_thr, double* dist_ub, double* summary, void* stream);

/* ---- introspection for parity tests and profiling */
/* which: 0 = relu(conv) output `a`, 1 = block output `y`, 2 = pooled / up-sampled output, 3 = dL/d(in0),
 * 4 = dL/d(in1). Returns the device pointer, element count and element size of layer `name`. */
int rvip_debug_buffer(const rvip_handle* h, const char* name, int which, void** ptr, long long* count,
                      int* elem_bytes);
int rvip_dropout_mask(uint64_t seed, uint32_t site, float rate, long long n_elems, uint8_t* keep, void* stream);
int rvip_dropout_site(const rvip_handle* h, const char* name, uint32_t* site, float* rate);
/* Per-kernel-class device timing: after rvip_profile(h, 1) every launch group is bracketed with CUDA
 * events; rvip_profile_read sums them (ms) and the launch counts per class and clears the log. */
#define RVIP_NUM_KERNEL_CLASSES 10
int rvip_profile(rvip_handle* h, int enable);
int rvip_profile_read(rvip_handle* h, float ms[RVIP_NUM_KERNEL_CLASSES], long long launches[RVIP_NUM_KERNEL_CLASSES]);
/* "class,layer:op,ms" lines, one per launch group, of the log consumed by the last rvip_profile_read */
const char* rvip_profile_detail(const rvip_handle* h);
const char* rvip_kernel_class_name(int cls);
long long rvip_launch_count(const rvip_handle* h); /* kernels launched by this handle so far */

/* ---- single-op entry points (unit tests drive the tensor-core kernels directly).
 * tcgen05 implicit-GEMM conv: in0/in1 NHWC bf16 (C0 + C1 channels), w_packed [Cout][9][C0+C1] bf16,
 * out NHWC bf16; mode 0 = bias+ReLU+stats, 1 = bias+ReLU, 2 = linear (dgrad; out1 gets channels >= out_split) */
int rvip_conv3x3_tc(const void* in0, const void* in1, int C0, int C1, const void* w_packed, const float* bias,
                    void* out0, void* out1, int out_split, double* stats, int B, int H, int W, int Cout, int mode,
                    void* stream);
/* tcgen05 wgrad: dw [9][C0+C1][Cout] fp32 += sum_p x[p+tap] * dz[p] */
int rvip_wgrad3x3_tc(const void* x0, const void* x1, int C0, int C1, const void* dz, float* dw, int B, int H, int W,
                     int Cout, void* stream);

/* Row-tiled variants (rows of 128-pixel tiles; the conv accepts a partial last tile when >= 3/4 of it is image, the
 * weight gradient needs W % 128 == 0): the input tile is staged once with its halo and all nine taps are
 * issued from shifted shared-memory descriptors. Same contracts as the two entry points above. */
int rvip_conv3x3_row(const void* in0, const void* in1, int C0, int C1, const void* w_packed, const float* bias,
                     void* out0, void* out1, int out_split, double* stats, int B, int H, int W, int Cout, int mode,
                     int base_offset_mode, void* stream);
int rvip_wgrad3x3_row(const void* x0, const void* x1, int C0, int C1, const void* dz, float* dw, int B, int H, int W,
                      int Cout, void* stream);

/* Halo-staged conv for the deep levels (any H, W; channels % 64 == 0): 16 x 16 pixel blocks staged once with
 * their halo, two 128-row accumulators per CTA share every weight tile. Same contract as rvip_conv3x3_tc. */
int rvip_conv3x3_halo(const void* in0, const void* in1, int C0, int C1, const void* w_packed, const float* bias,
                      void* out0, void* out1, int out_split, double* stats, int B, int H, int W, int Cout, int mode,
                      void* stream);
/* profiling aid: dbg (device, [148][8] int64) receives per CTA {issue-loop cycles, cycles waiting for a free
 * accumulator, for the activation block, for a weight tile, whole-kernel cycles, epilogue cycles} of subsequent rvip_conv3x3_halo calls; NULL = off */
int rvip_conv3x3_halo_debug(long long* dbg);
/* Halo-staged wgrad (any H, W; channels % 32 == 0): one input-channel chunk per CTA, the x tile staged once
 * with its halo, all nine taps accumulated in TMEM. Same contract as rvip_wgrad3x3_tc. */
int rvip_wgrad3x3_halo(const void* x0, const void* x1, int C0, int C1, const void* dz, float* dw, int B, int H, int W,
                       int Cout, void* stream);
/* Phase-decomposed decoder up-convolution (UpSampling2D(2) -> Conv2D 3x3 -> ReLU, src/models/KerasLayers.py:756-759)
 * computed from the LOW-resolution tensor: output pixel (2i + a, 2j + b) sees only a 2 x 2 low-resolution neighbourhood,
 * so the 3x3 taps that coincide are pre-summed in fp32 and each of the four phases is a 4-tap convolution (2.25x fewer
 * MMAs; the up-sampled tensor is never materialised for this op).  h, w = low-resolution size (any),
 * Cin % 64 == 0, C % 64 == 0 or C == 32.  w_hwio: fp32 [3][3][Cin][C] (device); packed_scratch: device bf16 buffer of
 * at least 2 * max(16*Cin*C, 768*Cin) elements (receives the packed forward and dgrad operands).
 *   dir 0: high[B,2h,2w,C] = relu(conv3x3(upsample2x(low[B,h,w,Cin])) + bias)
 *   dir 1: low[B,h,w,Cin]  = gradient of that convolution w.r.t. its low-resolution input, from high = dz[B,2h,2w,C]
 * transposed = 1: the same kernels compute Conv2DTranspose(3x3, strides 2, padding 'same') -> ReLU (USE_UPSAMPLE falsy,
 * KerasLayers.py:762-765): w is then the Keras kernel (kh, kw, C, Cin) and every phase weight is a single tap
 * (output parity 0 sees tap 2 on neighbour i - 1 and tap 0 on i, parity 1 sees tap 1 on i).
 * Synchronises the stream before returning (test / profiling entry point). */
int rvip_upconv3x3_halo(int dir, const void* low, const void* high, const float* w_hwio, const float* bias,
                        void* packed_scratch, int B, int h, int w, int Cin, int C, int transposed, void* stream);
/* Weight gradient of the same up-convolution from the LOW-resolution input: dw[3][3][Cin][C] (fp32, accumulated) from
 * x_low[B,h,w,Cin] and dz[B,2h,2w,C] (bf16).  Cin % 64 == 0, C % 32 == 0. */
int rvip_upconv_wgrad_halo(const void* x_low, const void* dz, float* dw, int B, int h, int w, int Cin, int C,
                           int transposed, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RVIP_H_ */
