// pp_ops.h — internal C++ declarations of every operator launcher (defined in conv_tc.cu,
// conv_simt.cu, ops.cu, loss.cu, common.cu). The public C ABI in capi.cu forwards to these.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pp {

int init_device(int device);
const char* last_error();

// ---- convolutions -------------------------------------------------------------------------------
// Optional epilogue of the forward convolutions: eval-mode BatchNorm + LeakyReLU folded into the conv (unet.py:188-190
// with running statistics): out = lrelu(acc * scale[c] + shift[c]); `shift` already contains bias * scale, so the
// conv is called WITHOUT a bias. Both arrays hold Cout floats and must be 16-byte aligned.
struct ConvAffine {
  const float* scale;
  const float* shift;
  float slope;
};
// stats/groups (optional): fused BatchNorm partial sums [groups][cout][2] of the rounded output (caller zeroes)
int conv3x3_tc(const void* x0, int C0, const void* x1, int C1, const void* wpack, const float* bias, void* out0,
               int outc0, int acc0, void* out1, int outc1, int acc1, int N, int H, int W, int dil, cudaStream_t s,
               double* stats = nullptr, int groups = 1, const ConvAffine* affine = nullptr);
// conv_rows.cu (wide layers, shared-memory-resident im2col, optional CTA pairs): candidate tilings and the launcher
struct RowsPlan {
  int block_n, bk, mt, pair, R, rbox, a_bytes, nb, smem;
  double cost;
};
int conv3x3_rows_plans(int N, int H, int W, int dil, int C0, int C1, int cout, RowsPlan* out, int max_out);
bool conv3x3_rows_applicable(int C0, int C1, int cout, int N, int H, int W, int dil);
int conv3x3_rows_tc(const void* x0, int C0, const void* x1, int C1, const void* wpack, const float* bias, void* out0,
                    int outc0, int acc0, void* out1, int outc1, int acc1, int N, int H, int W, int dil, cudaStream_t s,
                    double* stats, int groups, const ConvAffine* affine, const RowsPlan* plan = nullptr);
// g_oihw == nullptr: packed dwp[9][Cout][C0+C1] += ; else the OIHW gradient is accumulated in place (dwp is scratch
// for narrow sources, zeroed by the caller when conv3x3_wgrad_tc_uses_scratch())
int conv3x3_wgrad_tc(const void* dy, int Cout, const void* x0, int C0, const void* x1, int C1, float* dwp,
                     float* g_oihw, int N, int H, int W, int dil, cudaStream_t s, float* ws_split = nullptr,
                     long long ws_floats = 0);
bool conv3x3_wgrad_tc_uses_scratch(int Cout, int C0, int C1, int W, int dil);
int unpack_wgrad_range(const float* dwp, float* g, int Cout, int Cin, int ci_begin, int ci_count, int accumulate,
                       cudaStream_t s);
int conv3x3_simt(int dtype, const void* x0, int C0, const void* x1, int C1, const void* wpack, const float* bias,
                 void* out0, int outc0, int acc0, void* out1, int outc1, int acc1, int N, int H, int W, int dil,
                 cudaStream_t s);
int conv3x3_wgrad_simt(int dtype, const void* dy, int Cout, const void* x0, int C0, const void* x1, int C1, float* dw,
                       int N, int H, int W, int dil, cudaStream_t s);

// ---- ops.cu ---------------------------------------------------------------------------------------
int pack_weights(int dtype, const float* w, void* wf, void* wd, int Cout, int Cin, cudaStream_t s);
int pack_weights_multi(int dtype, int n, const float* const* w, void* const* wf, void* const* wd, const int* cout,
                       const int* cin, cudaStream_t s);
int unpack_wgrad(const float* dwp, float* g, int Cout, int Cin, int accumulate, cudaStream_t s);
// Cin > 1: x is NCHW fp32 [N][Cin][H][W] (multi-channel input, --input_ch)
int first_conv_fwd(int dtype, const float* x, const float* w, const float* bias, void* y, int N, int H, int W, int Cout,
                   cudaStream_t s, int Cin = 1, const ConvAffine* affine = nullptr);
int first_conv_wgrad(int dtype, const void* dy, const float* x, float* dw, int N, int H, int W, int Cout,
                     cudaStream_t s, int Cin = 1);
int head_fwd(int dtype, const void* a, const float* w, const float* bias, float* logits, long long P, int HW, int Cin,
             int C, cudaStream_t s);
int head_bwd(int dtype, const float* dlogits, const void* a, const float* w, void* da, float* dw, float* db, long long P,
             int HW, int Cin, int C, cudaStream_t s);
int bn_stats(int dtype, const void* y, double* sums, int G, long long Pg, int C, cudaStream_t s);
int bn_finalize(const double* sums, const float* gamma, const float* beta, float* running_mean, float* running_var,
                long long* nbt, float* coef, int G, long long Pg, int C, int training, float eps, float momentum,
                cudaStream_t s, int replicas = 1);
int bn_apply(int dtype, const void* y, const float* coef, void* a, int G, long long Pg, int C, float slope,
             cudaStream_t s);
int bn_bwd(int dtype, const void* da, const void* y, const float* coef, double* bsums, float* bcoef, float* dgamma,
           float* dbeta, float* dbias, void* dy, int G, long long Pg, int C, int training, float slope, cudaStream_t s,
           bool sums_zeroed = false);
// Eval-mode BatchNorm folded into the convolutions (bf16 path). coef[l] = [scale | shift | beta | 1/gamma] (4*C floats):
//   scale = gamma / sqrt(running_var + eps), shift = beta + (bias - running_mean) * scale  (one launch for all layers)
int bn_eval_coef_multi(int n, const float* const* gamma, const float* const* beta, const float* const* rmean,
                       const float* const* rvar, const float* const* bias, float* const* coef, const int* C, float eps,
                       cudaStream_t s);
// Backward of conv -> eval-BN -> LeakyReLU given only the saved ACTIVATION a (no pre-BN tensor exists in this mode):
//   dz = da * (a > 0 ? 1 : slope), dy = dz * scale; z = a > 0 ? a : a / slope, xhat = (z - beta) / gamma;
//   dgamma += sum dz * xhat, dbeta += sum dz, dbias += scale * sum dz   (ONE pass; the last block folds the sums).
// sums: reps * 2*C doubles ([replica][C][2]: the blocks spread their atomics over the replicas) + one unsigned ticket
// placed after them, zeroed by the caller. bn_bwd_replicas(C) is the count the UNet plan uses (<= kBnBwdReplicas).
constexpr int kBnBwdReplicas = 8;
int bn_bwd_replicas(int C);
int bn_bwd_eval(int dtype, const void* da, const void* a, const float* coef, double* sums, float* dgamma, float* dbeta,
                float* dbias, void* dy, long long P, int C, float slope, cudaStream_t s, int reps = 1);
int maxpool_fwd(int dtype, const void* x, void* y, int N, int H, int W, int C, cudaStream_t s);
int maxpool_bwd(int dtype, const void* x, const void* gy, void* gx, int N, int H, int W, int C, int accumulate,
                cudaStream_t s);
int upsample_nhwc_fwd(int dtype, const void* x, void* y, int N, int h, int w, int H, int W, int C, cudaStream_t s);
int upsample_nhwc_bwd(int dtype, const void* gy, void* gx, int N, int h, int w, int H, int W, int C, int accumulate,
                      cudaStream_t s);
int upsample_planes_fwd(const float* x, float* y, long long NC, int h, int w, int H, int W, cudaStream_t s);
int upsample_planes_bwd(const float* gy, float* gx, long long NC, int h, int w, int H, int W, cudaStream_t s);
int nchw_to_nhwc(int dtype, const float* src, void* dst, int N, int C, int HW, cudaStream_t s);
int nhwc_to_nchw(int dtype, const void* src, float* dst, int N, int C, int HW, cudaStream_t s);
// strided-conv / transposed-conv variant (unet.py:113-116,141): layout shuffles and zero-embedded kernels
int space_to_depth(int dtype, const void* x, void* y, int N, int Hs, int Ws, int C, cudaStream_t s);
int depth_to_space(int dtype, const void* y, void* x, int N, int Hs, int Ws, int C, int accumulate, cudaStream_t s);
int embed_s2_weight(const float* w, float* we, int Cout, int C, cudaStream_t s);
int collapse_s2_wgrad(const float* dwe, float* dw, int Cout, int C, cudaStream_t s);
int embed_ct_weight(const float* wt, float* we, int Cin, int Cout, int S, cudaStream_t s);
int collapse_ct_wgrad(const float* dwe, float* dwt, int Cin, int Cout, int S, cudaStream_t s);
int channel_scale(int dtype, const void* x, const float* scale, void* y, int N, int HW, int C, int ld, cudaStream_t s);
int strong_color_augment(const float* image, const float* params, float* out, int N, int HW, cudaStream_t s);
int adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
              float wd, int step, float grad_scale, cudaStream_t s);

// ---- loss.cu --------------------------------------------------------------------------------------
int onehot_argmax(const float* x, uint8_t* out, int N, int K, int HW, cudaStream_t s);
// aux_h > 0: za (and dza) are the LOW-resolution aux logits [N][C][aux_h][aux_w]; the kernels interpolate them at the
// labelled pixels of the W-wide label map (bilinear, align_corners=True) instead of reading an up-sampled tensor, and
// the backward pass scatters into a zeroed fixed-point scratch with integer atomics (order-independent) and converts it
// to dza. aux_h == 0: full-resolution planes [N][C][HW].
int scribble_loss_fwd(const float* zw, const float* zs, const float* za, const uint8_t* target, const float* mask,
                      double* acc, float* loss_pce, float* loss_ent, float* loss_cr, float* loss_aux, int N, int C,
                      int HW, int ignore_index, int do_ent, int cr_variant, cudaStream_t s, int aux_h = 0,
                      int aux_w = 0, int W = 0);
int scribble_loss_bwd(const float* zw, const float* zs, const float* za, const uint8_t* target, const float* mask,
                      const double* acc, const float* g_pce, const float* g_ent, const float* g_cr, const float* g_aux,
                      float* dzw, float* dzs, float* dza, int N, int C, int HW, int ignore_index, int do_ent,
                      int cr_variant, int detach_weak, cudaStream_t s, int aux_h = 0, int aux_w = 0, int W = 0,
                      long long* aux_scratch = nullptr);   // aux_h > 0: N*C*aux_h*aux_w int64 (fixed-point accumulators)
int pair_loss_fwd(const float* a, const float* b, const float* mask, double* pacc, float* loss, int N, int C, int HW,
                  int variant, cudaStream_t s);
int pair_loss_bwd(const float* a, const float* b, const float* mask, const double* pacc, const float* g, float* da,
                  float* db, int N, int C, int HW, int variant, cudaStream_t s);
int dice_fwd(const float* z, const float* label, double* sums, float* coef, float* loss, int N, int C, int HW,
             cudaStream_t s);
int dice_bwd(const float* z, const float* label, const float* coef, const float* g, float* dz, int N, int C, int HW,
             int accumulate, cudaStream_t s);
int memory_update_scratch_floats(int C, int hid);
// scribble: fp32 one-hot [N][K][H][W] (reference format), or nullptr with scribble_idx: uint8 index map [N][H][W]
int memory_update(int dtype, const void* feat, const float* scribble, float* bank, float* scratch, int C, int h, int w,
                  int H, int W, int hid, int cosine_mode, float m, float one_minus_m, cudaStream_t s,
                  const uint8_t* scribble_idx = nullptr);
int memory_loss_fwd(const float* bank, const float* wfc, float* loss, float* probs, int C, int hid, cudaStream_t s);
int memory_loss_bwd(const float* bank, const float* probs, const float* g, float* dwfc, int C, int hid, cudaStream_t s);

// validation Dice metric (utils/metrics.py:7-34), whole batch in one pass; scratch = 3*N*C doubles + N*C uint32
int dice_metric(const float* scores, const float* label, float* dice, void* scratch, int N, int C, int HW,
                cudaStream_t s);

}  // namespace pp
