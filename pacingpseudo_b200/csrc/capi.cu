// capi.cu — extern "C" surface declared in include/pacingpseudo_b200.h. Pure forwarding plus the
// bf16/fp32 dispatch of the conv entry points; no torch types, no allocation.
#include "../../include/pacingpseudo_b200.h"

#include "pp_common.cuh"
#include "pp_ops.h"

namespace pp {
struct UNetPlan;
UNetPlan* unet_create(int input_ch, int init_ch, int max_ch, int num_classes, int output_stride, int dtype,
                      int strided);
int unet_conv_kind(const UNetPlan* pl, int layer, int* kind, int* scale);
void unet_destroy(UNetPlan* pl);
int unet_set_grad_events(UNetPlan* pl, int n, const int* layers);
int unet_wait_grad_event(const UNetPlan* pl, int i, cudaStream_t s);
int unet_num_convs(const UNetPlan* pl);
long long unet_graph_replays();
int unet_conv_info(const UNetPlan* pl, int layer, int* cin, int* cout, int* dil, const char** name);
long long unet_workspace_bytes(const UNetPlan* pl, int N, int H, int W, int G);
int unet_activation(const UNetPlan* pl, const char* name, int N, int H, int W, int G, int* act_id, long long* offset,
                    int* C, int* h, int* w);
int unet_forward(const UNetPlan& pl, const float* x, void* const* params, void* ws, int N, int H, int W, int G,
                 int training, float* logits, cudaStream_t s);
int unet_backward(const UNetPlan& pl, const float* x, void* const* params, void* ws, int N, int H, int W, int G,
                  int training, const float* dlogits, int n_dfeat, const int* dfeat_act, const void* const* dfeat,
                  float* const* grads, cudaStream_t s);
}  // namespace pp

using namespace pp;
#define ST(s) static_cast<cudaStream_t>(s)

extern "C" {

int pp_init(int device) { return init_device(device); }
const char* pp_last_error(void) { return last_error(); }
int pp_version(void) { return 100; }
long long pp_launch_count(void) { return launch_count(); }
void pp_profile_enable(int on) { prof_enable(on); }
void pp_profile_reset(void) { prof_reset(); }
int pp_profile_collect(int family, double* ms, double* flops, long long* launches) {
  return prof_collect(family, ms, flops, launches);
}

int pp_unet_create(int input_ch, int init_ch, int max_ch, int num_classes, int output_stride, int dtype,
                   pp_unet_t* out) {
  *out = reinterpret_cast<pp_unet_t>(unet_create(input_ch, init_ch, max_ch, num_classes, output_stride, dtype, 0));
  return *out ? PP_OK : PP_ERR_INVALID;
}
int pp_unet_create_ex(int input_ch, int init_ch, int max_ch, int num_classes, int output_stride, int dtype,
                      int strided, pp_unet_t* out) {
  *out = reinterpret_cast<pp_unet_t>(
      unet_create(input_ch, init_ch, max_ch, num_classes, output_stride, dtype, strided));
  return *out ? PP_OK : PP_ERR_INVALID;
}
int pp_unet_conv_kind(pp_unet_t u, int layer, int* kind, int* scale) {
  return unet_conv_kind(reinterpret_cast<pp::UNetPlan*>(u), layer, kind, scale);
}
void pp_unet_destroy(pp_unet_t u) { unet_destroy(reinterpret_cast<pp::UNetPlan*>(u)); }
int pp_unet_set_grad_events(pp_unet_t u, int n, const int* layers) {
  return unet_set_grad_events(reinterpret_cast<pp::UNetPlan*>(u), n, layers);
}
int pp_unet_wait_grad_event(pp_unet_t u, int i, void* stream) {
  return unet_wait_grad_event(reinterpret_cast<pp::UNetPlan*>(u), i, static_cast<cudaStream_t>(stream));
}
int pp_unet_num_convs(pp_unet_t u) { return unet_num_convs(reinterpret_cast<pp::UNetPlan*>(u)); }
long long pp_graph_replays(void) { return unet_graph_replays(); }
int pp_unet_conv_info(pp_unet_t u, int layer, int* cin, int* cout, int* dil, const char** name) {
  return unet_conv_info(reinterpret_cast<pp::UNetPlan*>(u), layer, cin, cout, dil, name);
}
long long pp_unet_workspace_bytes(pp_unet_t u, int N, int H, int W, int G) {
  return unet_workspace_bytes(reinterpret_cast<pp::UNetPlan*>(u), N, H, W, G);
}
int pp_unet_activation(pp_unet_t u, const char* name, int N, int H, int W, int G, int* act_id, long long* offset,
                       int* C, int* h, int* w) {
  return unet_activation(reinterpret_cast<pp::UNetPlan*>(u), name, N, H, W, G, act_id, offset, C, h, w);
}
int pp_unet_forward(pp_unet_t u, const float* x, void* const* params, void* workspace, int N, int H, int W, int G,
                    int training, float* logits, void* stream) {
  return unet_forward(*reinterpret_cast<pp::UNetPlan*>(u), x, params, workspace, N, H, W, G, training, logits,
                      ST(stream));
}
int pp_unet_backward(pp_unet_t u, const float* x, void* const* params, void* workspace, int N, int H, int W, int G,
                     int training, const float* dlogits, int n_dfeat, const int* dfeat_act,
                     const void* const* dfeat, float* const* grads, void* stream) {
  return unet_backward(*reinterpret_cast<pp::UNetPlan*>(u), x, params, workspace, N, H, W, G, training, dlogits,
                       n_dfeat, dfeat_act, dfeat, grads, ST(stream));
}

int pp_conv3x3(int dtype, const void* x0, int C0, const void* x1, int C1, const void* wpack, const float* bias,
               void* out0, int oc0, int acc0, void* out1, int oc1, int acc1, int N, int H, int W, int dil,
               void* stream) {
  if (dtype == PP_BF16)
    return conv3x3_tc(x0, C0, x1, C1, wpack, bias, out0, oc0, acc0, out1, oc1, acc1, N, H, W, dil, ST(stream));
  return conv3x3_simt(dtype, x0, C0, x1, C1, wpack, bias, out0, oc0, acc0, out1, oc1, acc1, N, H, W, dil, ST(stream));
}
int pp_conv3x3_bn_stats(const void* x0, int C0, const void* x1, int C1, const void* wpack, const float* bias, void* y,
                        int Cout, double* stats, int groups, int N, int H, int W, int dil, void* stream) {
  return conv3x3_tc(x0, C0, x1, C1, wpack, bias, y, Cout, 0, nullptr, 0, 0, N, H, W, dil, ST(stream), stats, groups);
}
int pp_stat_replicas(void) { return kStatReplicas; }
int pp_conv3x3_wgrad(int dtype, const void* dy, int Cout, const void* x0, int C0, const void* x1, int C1, float* dwp,
                     int N, int H, int W, int dil, void* stream) {
  if (dtype == PP_BF16) return conv3x3_wgrad_tc(dy, Cout, x0, C0, x1, C1, dwp, nullptr, N, H, W, dil, ST(stream));
  return conv3x3_wgrad_simt(dtype, dy, Cout, x0, C0, x1, C1, dwp, N, H, W, dil, ST(stream));
}
int pp_conv3x3_wgrad_oihw(const void* dy, int Cout, const void* x0, int C0, const void* x1, int C1, float* dwp,
                          float* g_oihw, float* ws_split, long long ws_floats, int N, int H, int W, int dil,
                          void* stream) {
  PP_REQUIRE(g_oihw != nullptr && dwp != nullptr, "pp_conv3x3_wgrad_oihw: null gradient / scratch pointer");
  if (conv3x3_wgrad_tc_uses_scratch(Cout, C0, C1, W, dil))
    PP_CHECK_CUDA(cudaMemsetAsync(dwp, 0, sizeof(float) * 9 * Cout * (C0 + C1), ST(stream)));
  return conv3x3_wgrad_tc(dy, Cout, x0, C0, x1, C1, dwp, g_oihw, N, H, W, dil, ST(stream), ws_split, ws_floats);
}
int pp_conv3x3_reference(int dtype, const void* x0, int C0, const void* x1, int C1, const void* wpack,
                         const float* bias, void* out0, int oc0, int acc0, void* out1, int oc1, int acc1, int N,
                         int H, int W, int dil, void* stream) {
  return conv3x3_simt(dtype, x0, C0, x1, C1, wpack, bias, out0, oc0, acc0, out1, oc1, acc1, N, H, W, dil, ST(stream));
}
int pp_conv3x3_wgrad_reference(int dtype, const void* dy, int Cout, const void* x0, int C0, const void* x1, int C1,
                               float* dwp, int N, int H, int W, int dil, void* stream) {
  return conv3x3_wgrad_simt(dtype, dy, Cout, x0, C0, x1, C1, dwp, N, H, W, dil, ST(stream));
}
int pp_pack_weights(int dtype, const float* w, void* wf, void* wd, int Cout, int Cin, void* stream) {
  return pack_weights(dtype, w, wf, wd, Cout, Cin, ST(stream));
}
int pp_unpack_wgrad(const float* dwp, float* g, int Cout, int Cin, int accumulate, void* stream) {
  return unpack_wgrad(dwp, g, Cout, Cin, accumulate, ST(stream));
}
int pp_first_conv_fwd(int dtype, const float* x, const float* w, const float* bias, void* y, int N, int H, int W,
                      int Cout, void* stream) {
  return first_conv_fwd(dtype, x, w, bias, y, N, H, W, Cout, ST(stream));
}
int pp_first_conv_wgrad(int dtype, const void* dy, const float* x, float* dw, int N, int H, int W, int Cout,
                        void* stream) {
  return first_conv_wgrad(dtype, dy, x, dw, N, H, W, Cout, ST(stream));
}
int pp_head_fwd(int dtype, const void* a, const float* w, const float* bias, float* logits, long long P, int HW,
                int Cin, int C, void* stream) {
  return head_fwd(dtype, a, w, bias, logits, P, HW, Cin, C, ST(stream));
}
int pp_head_bwd(int dtype, const float* dlogits, const void* a, const float* w, void* da, float* dw, float* db,
                long long P, int HW, int Cin, int C, void* stream) {
  return head_bwd(dtype, dlogits, a, w, da, dw, db, P, HW, Cin, C, ST(stream));
}
int pp_bn_stats(int dtype, const void* y, double* sums, int G, long long Pg, int C, void* stream) {
  return bn_stats(dtype, y, sums, G, Pg, C, ST(stream));
}
int pp_bn_finalize(const double* sums, const float* gamma, const float* beta, float* running_mean,
                   float* running_var, long long* nbt, float* coef, int G, long long Pg, int C, int training,
                   float eps, float momentum, void* stream) {
  return bn_finalize(sums, gamma, beta, running_mean, running_var, nbt, coef, G, Pg, C, training, eps, momentum,
                     ST(stream));
}
int pp_bn_apply(int dtype, const void* y, const float* coef, void* a, int G, long long Pg, int C, float slope,
                void* stream) {
  return bn_apply(dtype, y, coef, a, G, Pg, C, slope, ST(stream));
}
int pp_bn_bwd(int dtype, const void* da, const void* y, const float* coef, double* bsums, float* bcoef,
              float* dgamma, float* dbeta, float* dbias, void* dy, int G, long long Pg, int C, int training,
              float slope, void* stream) {
  return bn_bwd(dtype, da, y, coef, bsums, bcoef, dgamma, dbeta, dbias, dy, G, Pg, C, training, slope, ST(stream));
}
int pp_bn_eval_coef(const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                    const float* conv_bias, float* coef, int C, float eps, void* stream) {
  return bn_eval_coef_multi(1, &gamma, &beta, &running_mean, &running_var, &conv_bias, &coef, &C, eps, ST(stream));
}
int pp_conv3x3_bn_eval(const void* x0, int C0, const void* x1, int C1, const void* wpack, const float* coef, void* a,
                       int Cout, float slope, int N, int H, int W, int dil, void* stream) {
  PP_REQUIRE(coef != nullptr, "pp_conv3x3_bn_eval: null coefficients");
  const ConvAffine af{coef, coef + Cout, slope};
  return conv3x3_tc(x0, C0, x1, C1, wpack, nullptr, a, Cout, 0, nullptr, 0, 0, N, H, W, dil, ST(stream), nullptr, 1, &af);
}
int pp_bn_bwd_eval(int dtype, const void* da, const void* a, const float* coef, double* sums, float* dgamma,
                   float* dbeta, float* dbias, void* dy, long long P, int C, float slope, void* stream) {
  return bn_bwd_eval(dtype, da, a, coef, sums, dgamma, dbeta, dbias, dy, P, C, slope, ST(stream));
}
int pp_channel_scale(int dtype, const void* x, const float* scale, void* y, int N, int HW, int C, int ld,
                     void* stream) {
  return channel_scale(dtype, x, scale, y, N, HW, C, ld, ST(stream));
}
int pp_space_to_depth(int dtype, const void* x, void* y, int N, int Hs, int Ws, int C, void* stream) {
  return space_to_depth(dtype, x, y, N, Hs, Ws, C, ST(stream));
}
int pp_depth_to_space(int dtype, const void* y, void* x, int N, int Hs, int Ws, int C, int accumulate, void* stream) {
  return depth_to_space(dtype, y, x, N, Hs, Ws, C, accumulate, ST(stream));
}
int pp_maxpool_fwd(int dtype, const void* x, void* y, int N, int H, int W, int C, void* stream) {
  return maxpool_fwd(dtype, x, y, N, H, W, C, ST(stream));
}
int pp_maxpool_bwd(int dtype, const void* x, const void* gy, void* gx, int N, int H, int W, int C, int accumulate,
                   void* stream) {
  return maxpool_bwd(dtype, x, gy, gx, N, H, W, C, accumulate, ST(stream));
}
int pp_upsample_nhwc_fwd(int dtype, const void* x, void* y, int N, int h, int w, int H, int W, int C, void* stream) {
  return upsample_nhwc_fwd(dtype, x, y, N, h, w, H, W, C, ST(stream));
}
int pp_upsample_nhwc_bwd(int dtype, const void* gy, void* gx, int N, int h, int w, int H, int W, int C,
                         int accumulate, void* stream) {
  return upsample_nhwc_bwd(dtype, gy, gx, N, h, w, H, W, C, accumulate, ST(stream));
}
int pp_upsample_planes_fwd(const float* x, float* y, long long NC, int h, int w, int H, int W, void* stream) {
  return upsample_planes_fwd(x, y, NC, h, w, H, W, ST(stream));
}
int pp_upsample_planes_bwd(const float* gy, float* gx, long long NC, int h, int w, int H, int W, void* stream) {
  return upsample_planes_bwd(gy, gx, NC, h, w, H, W, ST(stream));
}
int pp_nchw_to_nhwc(int dtype, const float* src, void* dst, int N, int C, int HW, void* stream) {
  return nchw_to_nhwc(dtype, src, dst, N, C, HW, ST(stream));
}
int pp_nhwc_to_nchw(int dtype, const void* src, float* dst, int N, int C, int HW, void* stream) {
  return nhwc_to_nchw(dtype, src, dst, N, C, HW, ST(stream));
}
int pp_onehot_argmax(const float* x, uint8_t* out, int N, int K, int HW, void* stream) {
  return onehot_argmax(x, out, N, K, HW, ST(stream));
}
int pp_scribble_loss_fwd(const float* zw, const float* zs, const float* za, const uint8_t* target,
                         const float* mask, double* acc, float* loss_pce, float* loss_ent, float* loss_cr,
                         float* loss_aux, int N, int C, int HW, int ignore_index, int do_ent, int cr_variant,
                         void* stream) {
  return scribble_loss_fwd(zw, zs, za, target, mask, acc, loss_pce, loss_ent, loss_cr, loss_aux, N, C, HW,
                           ignore_index, do_ent, cr_variant, ST(stream));
}
int pp_scribble_loss_bwd(const float* zw, const float* zs, const float* za, const uint8_t* target,
                         const float* mask, const double* acc, const float* g_pce, const float* g_ent,
                         const float* g_cr, const float* g_aux, float* dzw, float* dzs, float* dza, int N, int C,
                         int HW, int ignore_index, int do_ent, int cr_variant, int detach_weak, void* stream) {
  return scribble_loss_bwd(zw, zs, za, target, mask, acc, g_pce, g_ent, g_cr, g_aux, dzw, dzs, dza, N, C, HW,
                           ignore_index, do_ent, cr_variant, detach_weak, ST(stream));
}
int pp_scribble_loss_lowaux_fwd(const float* zw, const float* zs, const float* za_low, int aux_h, int aux_w,
                                const uint8_t* target, const float* mask, double* acc, float* loss_pce,
                                float* loss_ent, float* loss_cr, float* loss_aux, int N, int C, int H, int W,
                                int ignore_index, int do_ent, int cr_variant, void* stream) {
  PP_REQUIRE(za_low != nullptr && aux_h > 0 && aux_w > 0 && H > 0 && W > 0, "pp_scribble_loss_lowaux_fwd: bad aux tensor");
  return scribble_loss_fwd(zw, zs, za_low, target, mask, acc, loss_pce, loss_ent, loss_cr, loss_aux, N, C, H * W,
                           ignore_index, do_ent, cr_variant, ST(stream), aux_h, aux_w, W);
}
int pp_scribble_loss_lowaux_bwd(const float* zw, const float* zs, const float* za_low, int aux_h, int aux_w,
                                const uint8_t* target, const float* mask, const double* acc, const float* g_pce,
                                const float* g_ent, const float* g_cr, const float* g_aux, float* dzw, float* dzs,
                                float* dza_low, long long* dza_scratch, int N, int C, int H, int W, int ignore_index,
                                int do_ent, int cr_variant, int detach_weak, void* stream) {
  PP_REQUIRE(za_low != nullptr && aux_h > 0 && aux_w > 0 && H > 0 && W > 0, "pp_scribble_loss_lowaux_bwd: bad aux tensor");
  PP_REQUIRE((dza_low == nullptr) == (dza_scratch == nullptr), "pp_scribble_loss_lowaux_bwd: dza_low and dza_scratch go together");
  return scribble_loss_bwd(zw, zs, za_low, target, mask, acc, g_pce, g_ent, g_cr, g_aux, dzw, dzs, dza_low, N, C, H * W,
                           ignore_index, do_ent, cr_variant, detach_weak, ST(stream), aux_h, aux_w, W, dza_scratch);
}
int pp_pair_loss_fwd(const float* a, const float* b, const float* mask, double* pacc, float* loss, int N, int C,
                     int HW, int variant, void* stream) {
  return pair_loss_fwd(a, b, mask, pacc, loss, N, C, HW, variant, ST(stream));
}
int pp_pair_loss_bwd(const float* a, const float* b, const float* mask, const double* pacc, const float* g,
                     float* da, float* db, int N, int C, int HW, int variant, void* stream) {
  return pair_loss_bwd(a, b, mask, pacc, g, da, db, N, C, HW, variant, ST(stream));
}
int pp_dice_fwd(const float* z, const float* label, double* sums, float* coef, float* loss, int N, int C, int HW,
                void* stream) {
  return dice_fwd(z, label, sums, coef, loss, N, C, HW, ST(stream));
}
int pp_dice_bwd(const float* z, const float* label, const float* coef, const float* g, float* dz, int N, int C,
                int HW, int accumulate, void* stream) {
  return dice_bwd(z, label, coef, g, dz, N, C, HW, accumulate, ST(stream));
}
int pp_memory_update_scratch_floats(int C, int hid) { return memory_update_scratch_floats(C, hid); }
int pp_memory_update(int dtype, const void* feat, const float* scribble, float* bank, float* scratch, int C, int h,
                     int w, int H, int W, int hid, int cosine_mode, float m, float one_minus_m, void* stream) {
  return memory_update(dtype, feat, scribble, bank, scratch, C, h, w, H, W, hid, cosine_mode, m, one_minus_m,
                       ST(stream));
}
int pp_memory_update_idx(int dtype, const void* feat, const uint8_t* scribble_idx, float* bank, float* scratch, int C,
                         int h, int w, int H, int W, int hid, int cosine_mode, float m, float one_minus_m, void* stream) {
  return memory_update(dtype, feat, nullptr, bank, scratch, C, h, w, H, W, hid, cosine_mode, m, one_minus_m, ST(stream),
                       scribble_idx);
}
int pp_strong_color_augment(const float* image, const float* params, float* out, int N, int HW, void* stream) {
  return strong_color_augment(image, params, out, N, HW, ST(stream));
}
int pp_memory_loss_fwd(const float* bank, const float* wfc, float* loss, float* probs, int C, int hid,
                       void* stream) {
  return memory_loss_fwd(bank, wfc, loss, probs, C, hid, ST(stream));
}
int pp_memory_loss_bwd(const float* bank, const float* probs, const float* g, float* dwfc, int C, int hid,
                       void* stream) {
  return memory_loss_bwd(bank, probs, g, dwfc, C, hid, ST(stream));
}
int pp_dice_metric(const float* scores, const float* label, float* dice, void* scratch, int N, int C, int HW,
                   void* stream) {
  return dice_metric(scores, label, dice, scratch, N, C, HW, ST(stream));
}
int pp_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                 float eps, float weight_decay, int step, float grad_scale, void* stream) {
  return adam_step(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, step, grad_scale, ST(stream));
}

}  // extern "C"
