/* pacingpseudo_b200.h — C ABI of libpacingpseudo_b200.so (hand-written CUDA for sm_100a).
 *
 * Drop-in boundary for the PacingPseudo training-step hot path. The reference is pure PyTorch, so
 * the "FFI" each entry point replaces is the torch op call site named beside it (paths relative to
 * /root/reference). Conventions:
 *   - every function returns 0 on success or a negative code; pp_last_error() gives the message
 *     (thread-local); nothing throws across the boundary;
 *   - all pointers are DEVICE pointers owned by the caller (PyTorch's allocator); the library
 *     allocates nothing and never synchronises the host, except pp_init();
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it;
 *   - activations are NHWC of element type `dtype` (PP_DTYPE_BF16 = tcgen05 path, PP_DTYPE_F32 =
 *     CUDA-core fp32 precision mode); logits, losses, parameters and gradients are fp32, with the
 *     reference's layouts at the API (NCHW logits, OIHW weights).
 *   - "G" = BatchNorm statistics groups: the weak and strong branch of the siamese step are batched
 *     into one tensor of N = G * Ng images but keep per-branch batch statistics.
 */
#ifndef PACINGPSEUDO_B200_H_
#define PACINGPSEUDO_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PP_DTYPE_F32 0
#define PP_DTYPE_BF16 1

/* consistency-loss variants (train_chaos.py:138 --loss_cr_variants; consistency_reglur_memory.py:56-63) */
#define PP_CR_NONE 0
#define PP_CR_CE 1
#define PP_CR_L1 2
#define PP_CR_L2 3
#define PP_CR_KL 4

/* ---- library ------------------------------------------------------------------------------- */
int pp_init(int device);                 /* selects the device, checks sm_100, resolves cuTensorMapEncodeTiled */
const char* pp_last_error(void);
int pp_version(void);
/* measurement hooks (bench.py): kernels launched so far by this library; optional CUDA-event bracketing of
 * every tcgen05 conv launch (and the fused loss passes) on its own stream. family 0 = conv3x3 forward/dgrad,
 * 1 = conv3x3 wgrad, 2 = fused scribble loss forward/backward (its "flops" are algorithmic BYTES), < 0 = both conv families.
 * pp_profile_collect returns the device time during which at least one launch of the family was running (union of
 * the launch intervals: weight-gradient kernels overlap dgrad on a side stream), the algorithmic FLOPs and the
 * launch count (HOST pointers). */
long long pp_launch_count(void);
void pp_profile_enable(int on);
void pp_profile_reset(void);
int pp_profile_collect(int family, double* ms, double* flops, long long* launches);

/* ---- whole-UNet executor (models/unet.py:10-98 UNet.__init__/forward) ------------------------- */
typedef struct UNetPlan* pp_unet_t;
int pp_unet_create(int input_ch, int init_ch, int max_ch, int num_classes, int output_stride, int dtype,
                   pp_unet_t* out);
/* strided = 1: the is_stride_conv + is_trans_conv variant (unet.py:25,113-116,141): the first conv of every
 * sub-sampling encoder block has stride 2 (no max-pool) and every decoder block up-samples with
 * ConvTranspose2d(lower_ch, skip_ch, k = s, stride = s, bias=False). pp_unet_create == strided 0. */
int pp_unet_create_ex(int input_ch, int init_ch, int max_ch, int num_classes, int output_stride, int dtype,
                      int strided, pp_unet_t* out);
void pp_unet_destroy(pp_unet_t u);
int pp_unet_num_convs(pp_unet_t u);
/* Whole-pass CUDA-graph replay: pp_unet_forward / pp_unet_backward called again with the SAME pointers, shapes and
 * flags are captured once (on an internal stream; the ~200-260 launches over the plan's internal streams become one
 * graph) and replayed with one cudaGraphLaunch on the caller's stream. PP_GRAPHS=0 disables it; passes that record
 * data-parallel gradient events (pp_unet_set_grad_events) always run eagerly. Number of replays so far: */
long long pp_graph_replays(void);
/* layer -> Cin, Cout, dilation and module path ("enc_block1.conv_block.conv_layer1", ...) */
int pp_unet_conv_info(pp_unet_t u, int layer, int* cin, int* cout, int* dil, const char** name);
/* kind 0: Conv2d 3x3 stride 1 + BN + LeakyReLU; 1: the same with stride 2 (scale = 2); 2: ConvTranspose2d with
 * kernel = stride = scale, weight [Cin][Cout][scale][scale], no bias / BN: its params slots 1..6 and grads slots 1..3
 * are NULL. pp_unet_conv_info reports the PARAMETER's Cin / Cout for every kind. */
int pp_unet_conv_kind(pp_unet_t u, int layer, int* kind, int* scale);
long long pp_unet_workspace_bytes(pp_unet_t u, int N, int H, int W, int G);
/* byte offset / shape of a named end point ("encoder/stage6", ... unet.py:82-97) inside the workspace */
int pp_unet_activation(pp_unet_t u, const char* name, int N, int H, int W, int G, int* act_id, long long* offset,
                       int* C, int* h, int* w);
/* params: per conv layer [weight OIHW, bias, bn.weight, bn.bias, bn.running_mean, bn.running_var,
 * bn.num_batches_tracked(int64)] in pp_unet_conv_info order, then head [weight, bias].
 * x: fp32 NCHW [N,input_ch,H,W]; logits out: fp32 NCHW [N,num_classes,H,W]. training=1: batch statistics +
 * running-stat update (unet.py:189 BatchNorm2d in train()), 0: running statistics. */
int pp_unet_forward(pp_unet_t u, const float* x, void* const* params, void* workspace, int N, int H, int W, int G,
                    int training, float* logits, void* stream);
/* grads: per conv layer [dweight, dbias, dgamma, dbeta], head [dweight, dbias]; fp32, ACCUMULATED (+=).
 * dfeat: optional gradients w.r.t. named end points (NHWC, activation dtype), e.g. from the aux path. */
int pp_unet_backward(pp_unet_t u, const float* x, void* const* params, void* workspace, int N, int H, int W, int G,
                     int training, const float* dlogits, int n_dfeat, const int* dfeat_act,
                     const void* const* dfeat, float* const* grads, void* stream);
/* Data-parallel overlap (NEW functionality; the reference is single-GPU). pp_unet_backward walks the
 * layers from the last to the first; after pp_unet_set_grad_events(u, n, layers) it records event i on its stream as
 * soon as every parameter gradient of conv layers >= layers[i] (and of the head) has been enqueued.
 * pp_unet_wait_grad_event makes `stream` (e.g. the NCCL stream) wait for event i, so the all-reduce of that slice of
 * the gradient buffer overlaps the rest of the backward pass. The library owns the events. */
int pp_unet_set_grad_events(pp_unet_t u, int n, const int* layers);
int pp_unet_wait_grad_event(pp_unet_t u, int i, void* stream);

/* ---- single operators (also used by the aux path and the op-level parity tests) --------------- */
/* nn.Conv2d 3x3 stride 1 pad=dil (unet.py:188; aux_path_memory.py:24): y[N,H,W,oc0(+oc1)] from the virtual
 * concat of x0[N,H,W,C0] and x1[N,H,W,C1] (unet.py:151; aux_path_memory.py:49), packed weights
 * [9][Cout][C0+C1] (pp_pack_weights). Output columns [0,oc0) go to out0, the rest to out1 (dgrad of a
 * concat); accN=1 adds to the destination. dtype BF16 -> tcgen05 kernel, F32 -> CUDA-core kernel. */
int pp_conv3x3(int dtype, const void* x0, int C0, const void* x1, int C1, const void* wpack, const float* bias,
               void* out0, int oc0, int acc0, void* out1, int oc1, int acc1, int N, int H, int W, int dil,
               void* stream);
/* bf16 forward conv with the BatchNorm batch statistics fused into the epilogue (unet.py:188-189): besides y it
 * accumulates sum / sum-of-squares of the ROUNDED output per statistics group into
 * stats[pp_stat_replicas()][groups][Cout][2] doubles (caller zeroes; the replicas are summed when folding).
 * A 128-pixel tile must not straddle two groups (error otherwise). */
int pp_conv3x3_bn_stats(const void* x0, int C0, const void* x1, int C1, const void* wpack, const float* bias, void* y,
                        int Cout, double* stats, int groups, int N, int H, int W, int dil, void* stream);
int pp_stat_replicas(void);
/* weight gradient of the same conv: dwp[9][Cout][C0+C1] fp32 += (caller zeroes) */
int pp_conv3x3_wgrad(int dtype, const void* dy, int Cout, const void* x0, int C0, const void* x1, int C1, float* dwp,
                     int N, int H, int W, int dil, void* stream);
/* bf16 weight gradient ACCUMULATED straight into the reference-layout gradient g_oihw[Cout][C0+C1][3][3] (the
 * nn.Conv2d weight.grad, unet.py:188). dwp: packed scratch of 9*Cout*(C0+C1) floats (used, and zeroed here, only for
 * 32/64-channel sources). ws_split (optional, ws_floats floats, ideally >= 2*9*Cout*(C0+C1)): scratch of the
 * deterministic split-K path - every K split stores its partial gradient and a fixed-order reduction folds them in,
 * so the result is bit-reproducible; without it the splits are accumulated with fp32 atomics. */
int pp_conv3x3_wgrad_oihw(const void* dy, int Cout, const void* x0, int C0, const void* x1, int C1, float* dwp,
                          float* g_oihw, float* ws_split, long long ws_floats, int N, int H, int W, int dil,
                          void* stream);
/* CUDA-core twin of the bf16 tcgen05 kernels (test cross-check only) */
int pp_conv3x3_reference(int dtype, const void* x0, int C0, const void* x1, int C1, const void* wpack,
                         const float* bias, void* out0, int oc0, int acc0, void* out1, int oc1, int acc1, int N,
                         int H, int W, int dil, void* stream);
int pp_conv3x3_wgrad_reference(int dtype, const void* dy, int Cout, const void* x0, int C0, const void* x1, int C1,
                               float* dwp, int N, int H, int W, int dil, void* stream);
/* OIHW fp32 -> forward pack wf[9][Cout][Cin] and dgrad pack wd[9][Cin][Cout] (flipped taps) */
int pp_pack_weights(int dtype, const float* w_oihw, void* wf, void* wd, int Cout, int Cin, void* stream);
int pp_unpack_wgrad(const float* dwp, float* g_oihw, int Cout, int Cin, int accumulate, void* stream);
/* first conv, Cin = 1 (unet.py:28); the whole-UNet executor also handles --input_ch 2..16 (NCHW fp32 input) */
int pp_first_conv_fwd(int dtype, const float* x, const float* w, const float* bias, void* y, int N, int H, int W,
                      int Cout, void* stream);
int pp_first_conv_wgrad(int dtype, const void* dy, const float* x, float* dw, int N, int H, int W, int Cout,
                        void* stream);
/* 1x1 heads (unet.py:60 final_conv; aux_path_memory.py:32 fc_cls): NHWC in, NCHW fp32 logits out */
int pp_head_fwd(int dtype, const void* a, const float* w, const float* bias, float* logits, long long P, int HW,
                int Cin, int C, void* stream);
int pp_head_bwd(int dtype, const float* dlogits, const void* a, const float* w, void* da, float* dw, float* db,
                long long P, int HW, int Cin, int C, void* stream);
/* BatchNorm2d + LeakyReLU(0.01) (unet.py:189-190,193). coef = [G][4][C] scale, shift, mean, rstd. */
int pp_bn_stats(int dtype, const void* y, double* sums, int G, long long Pg, int C, void* stream);
int pp_bn_finalize(const double* sums, const float* gamma, const float* beta, float* running_mean,
                   float* running_var, long long* num_batches_tracked, float* coef, int G, long long Pg, int C,
                   int training, float eps, float momentum, void* stream);
int pp_bn_apply(int dtype, const void* y, const float* coef, void* a, int G, long long Pg, int C, float slope,
                void* stream);
int pp_bn_bwd(int dtype, const void* da, const void* y, const float* coef, double* bsums, float* bcoef,
              float* dgamma, float* dbeta, float* dbias, void* dy, int G, long long Pg, int C, int training,
              float slope, void* stream);
/* Eval-mode BatchNorm (running statistics; the reference's steady state: train_chaos.py:370 calls model.eval() for
 * validation and never model.train() again) folded into the bf16 convolution (unet.py:188-190):
 *   pp_bn_eval_coef:     coef[4*C] = scale | shift | beta | 1/gamma, scale = gamma / sqrt(running_var + eps),
 *                        shift = beta + (conv_bias - running_mean) * scale  (conv_bias may be NULL)
 *   pp_conv3x3_bn_eval:  a = LeakyReLU(conv3x3(x0 ++ x1, wpack) * scale + shift) written as the bf16 NHWC activation —
 *                        no pre-BatchNorm tensor, no separate normalisation pass (coef must be 16-byte aligned)
 *   pp_bn_bwd_eval:      backward of BatchNorm + LeakyReLU from the saved ACTIVATION in one pass:
 *                        dy = da * lrelu'(a) * scale; dgamma += sum dz * xhat, dbeta += sum dz, dbias += scale * sum dz
 *                        with xhat = (lrelu^-1(a) - beta) / gamma. sums: 2*C doubles + 16 bytes, zeroed by the caller. */
int pp_bn_eval_coef(const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                    const float* conv_bias, float* coef, int C, float eps, void* stream);
int pp_conv3x3_bn_eval(const void* x0, int C0, const void* x1, int C1, const void* wpack, const float* coef, void* a,
                       int Cout, float slope, int N, int H, int W, int dil, void* stream);
int pp_bn_bwd_eval(int dtype, const void* da, const void* a, const float* coef, double* sums, float* dgamma,
                   float* dbeta, float* dbias, void* dy, long long P, int C, float slope, void* stream);
/* nn.Dropout2d of the aux path (aux_path_memory.py:23,31), forward and backward: y[n,p,c] = x[n,p,c] *
 * scale[n*ld + c] on NHWC tensors [N][HW][C]; scale holds 0 or 1/(1-p) per (sample, channel); y may alias x. */
int pp_channel_scale(int dtype, const void* x, const float* scale, void* y, int N, int HW, int C, int ld,
                     void* stream);
/* Layout shuffles of the strided-conv / transposed-conv UNet variant (unet.py:113-116,141; see pp_unet_create_ex):
 * x[N, 2Hs, 2Ws, C] <-> y[N, Hs, Ws, 4C] with y channel (sy*2+sx)*C + c = x[2Y+sy, 2X+sx, c]; accumulate: x += . */
int pp_space_to_depth(int dtype, const void* x, void* y, int N, int Hs, int Ws, int C, void* stream);
int pp_depth_to_space(int dtype, const void* y, void* x, int N, int Hs, int Ws, int C, int accumulate, void* stream);
/* MaxPool2d(2,2) (unet.py:109) */
int pp_maxpool_fwd(int dtype, const void* x, void* y, int N, int H, int W, int C, void* stream);
int pp_maxpool_bwd(int dtype, const void* x, const void* gy, void* gx, int N, int H, int W, int C, int accumulate,
                   void* stream);
/* bilinear, align_corners=True (unet.py:144; aux_path_memory.py:52) */
int pp_upsample_nhwc_fwd(int dtype, const void* x, void* y, int N, int h, int w, int H, int W, int C, void* stream);
int pp_upsample_nhwc_bwd(int dtype, const void* gy, void* gx, int N, int h, int w, int H, int W, int C,
                         int accumulate, void* stream);
int pp_upsample_planes_fwd(const float* x, float* y, long long NC, int h, int w, int H, int W, void* stream);
int pp_upsample_planes_bwd(const float* gy, float* gx, long long NC, int h, int w, int H, int W, void* stream);
int pp_nchw_to_nhwc(int dtype, const float* src, void* dst, int N, int C, int HW, void* stream);
int pp_nhwc_to_nchw(int dtype, const void* src, float* dst, int N, int C, int HW, void* stream);

/* ---- losses (losses/losses.py) ------------------------------------------------------------------ */
/* torch.argmax(one_hot, 1) (consistency_reglur_memory.py:31; aux_path_memory.py:55; upper_bound_chaos.py:160) */
int pp_onehot_argmax(const float* x, uint8_t* out, int N, int K, int HW, void* stream);
/* ONE pass: partial CE (losses.py:35-43), entropy (9-24), consistency (45-116), aux partial CE.
 * zs / za / mask may be NULL. acc: 8 doubles of scratch kept for the backward. Each loss is written to
 * its own fp32 scalar. */
int pp_scribble_loss_fwd(const float* zw, const float* zs, const float* za, const uint8_t* target,
                         const float* mask, double* acc, float* loss_pce, float* loss_ent, float* loss_cr,
                         float* loss_aux, int N, int C, int HW, int ignore_index, int do_ent, int cr_variant,
                         void* stream);
int pp_scribble_loss_bwd(const float* zw, const float* zs, const float* za, const uint8_t* target,
                         const float* mask, const double* acc, const float* g_pce, const float* g_ent,
                         const float* g_cr, const float* g_aux, float* dzw, float* dzs, float* dza, int N, int C,
                         int HW, int ignore_index, int do_ent, int cr_variant, int detach_weak, void* stream);
/* The same pass with the aux logits taken at the resolution fc_cls produces them, za_low [N][C][aux_h][aux_w]
 * (aux_path_memory.py:51), instead of the F.interpolate(bilinear, align_corners=True) output of aux_path_memory.py:52:
 * only labelled pixels read the aux logits, so the kernels interpolate them there (same arithmetic as
 * pp_upsample_planes_fwd) and no full-resolution aux tensor is written or read. The backward pass accumulates the
 * interpolation's transpose in dza_scratch (N*C*aux_h*aux_w int64, 2^-44 fixed point, integer atomics: the sum does
 * not depend on the order the blocks arrive in, so the gradient is bit-reproducible) and writes dza_low
 * [N][C][aux_h][aux_w] from it. */
int pp_scribble_loss_lowaux_fwd(const float* zw, const float* zs, const float* za_low, int aux_h, int aux_w,
                                const uint8_t* target, const float* mask, double* acc, float* loss_pce,
                                float* loss_ent, float* loss_cr, float* loss_aux, int N, int C, int H, int W,
                                int ignore_index, int do_ent, int cr_variant, void* stream);
int pp_scribble_loss_lowaux_bwd(const float* zw, const float* zs, const float* za_low, int aux_h, int aux_w,
                                const uint8_t* target, const float* mask, const double* acc, const float* g_pce,
                                const float* g_ent, const float* g_cr, const float* g_aux, float* dzw, float* dzs,
                                float* dza_low, long long* dza_scratch, int N, int C, int H, int W, int ignore_index,
                                int do_ent, int cr_variant, int detach_weak, void* stream);
/* stand-alone soft_label_cross_entropy_loss (a = logits, b = probabilities; losses.py:45-62), l1_loss and
 * l2_loss (a, b = probabilities; losses.py:64-96); variant = PP_CR_CE / PP_CR_L1 / PP_CR_L2. pacc: 8 doubles. */
int pp_pair_loss_fwd(const float* a, const float* b, const float* mask, double* pacc, float* loss, int N, int C,
                     int HW, int variant, void* stream);
int pp_pair_loss_bwd(const float* a, const float* b, const float* mask, const double* pacc, const float* g,
                     float* da, float* db, int N, int C, int HW, int variant, void* stream);
/* dice_loss_fn (losses.py:147-162). sums: [N][C][3] doubles, coef: [N][C][2] floats (kept for backward) */
int pp_dice_fwd(const float* z, const float* label, double* sums, float* coef, float* loss, int N, int C, int HW,
                void* stream);
int pp_dice_bwd(const float* z, const float* label, const float* coef, const float* g, float* dz, int N, int C,
                int HW, int accumulate, void* stream);
/* AuxPath.memory_update (aux_path_memory.py:68-116): sample 0 only; feat = aux_features [N,h,w,hid].
 * scratch: pp_memory_update_scratch_floats(C, hid) floats (per-slice partial sums, reduced in a fixed order). */
int pp_memory_update_scratch_floats(int C, int hid);
int pp_memory_update(int dtype, const void* feat, const float* scribble, float* bank, float* scratch, int C, int h,
                     int w, int H, int W, int hid, int cosine_mode, float m, float one_minus_m, void* stream);
/* the same update from a COMPACT scribble: uint8 class-index map [N][H][W] (C = unlabelled) instead of the fp32 one-hot
 * tensor ToTorchTensor builds (datasets/augmentations.py:421-446); SURVEY 8f N3: 24x less host->device traffic. */
int pp_memory_update_idx(int dtype, const void* feat, const uint8_t* scribble_idx, float* bank, float* scratch, int C,
                         int h, int w, int H, int W, int hid, int cosine_mode, float m, float one_minus_m, void* stream);
/* Strong colour augmentation on the device: Brightness -> Contrast -> GammaAugmentation(retain_stats)
 * of datasets/augmentations.py:98-166 as CHAOSTwoStream applies them (chaos_dataset.py:68-75) to each base-transformed
 * slice. image/out: fp32 [N][HW]; params: [N][8] = {apply_brightness, b, apply_contrast, a, apply_gamma, gamma, 0, 0}. */
int pp_strong_color_augment(const float* image, const float* params, float* out, int N, int HW, void* stream);
/* cross_entropy_loss(fc_cls(memory_bank), arange(C)) (aux_path_memory.py:61; consistency_reglur_memory.py:94) */
int pp_memory_loss_fwd(const float* bank, const float* wfc, float* loss, float* probs, int C, int hid,
                       void* stream);
int pp_memory_loss_bwd(const float* bank, const float* probs, const float* g, float* dwfc, int C, int hid,
                       void* stream);

/* ---- validation metric --------------------------------------------------------------
 * utils/metrics.py:7-34 compute_dice, which train_chaos.py:386-390 calls per sample on host copies of the softmax
 * values: here for the whole batch in one pass. scores, label: fp32 NCHW [N][C][HW] (label one-hot); dice: fp32
 * [N][C], NaN where the class is absent from both the arg-max prediction and the label (the caller skips those);
 * scratch: 3*N*C doubles followed by N*C uint32 (zeroed by the call). */
int pp_dice_metric(const float* scores, const float* label, float* dice, void* scratch, int N, int C, int HW,
                   void* stream);

/* ---- optimizer (train_chaos.py:219 torch.optim.Adam(lr, weight_decay)) --------------------------- */
int pp_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                 float eps, float weight_decay, int step, float grad_scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PACINGPSEUDO_B200_H_ */
