// unet_plan.cu — host-side executor for the whole UNet forward / backward on one stream.
//
// Mirrors the dataflow of /root/reference/models/unet.py:10-98 (6 encoder stages, 5 decoder stages,
// 1x1 head; output_stride 8/16/32; maxpool + bilinear variant) as a static op list over NHWC buffers
// carved from ONE caller-owned workspace. One C call launches the ~200 kernels of a pass, so the
// Python side pays no per-layer overhead. The channel concat of the decoder is never materialised
// (two-source K loop in the conv kernels); scale-1 "upsampling" (output_stride 8) is a no-op alias.
#include <atomic>
#include <cstdlib>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "pp_common.cuh"
#include "pp_ops.h"

namespace pp {

namespace {

struct Act { int C; int res; };  // channels, spatial divisor relative to the input
// Layer kinds. KIND_S2 / KIND_CT belong to the strided-conv / transposed-conv variant (unet.py:113-116,141) and run on
// the same stride-1 conv kernels with zero-embedded weights (ops.cu: embed_s2_weight / embed_ct_weight):
//   KIND_S2: Conv2d(3x3, stride 2) + BN + LeakyReLU; the input activation is the space-to-depth tensor (cin0 = 4 C)
//   KIND_CT: ConvTranspose2d(k = s = S, bias=False): a plain (no BN, no bias) conv to S*S*Cout channels (cout), then
//            a depth-to-space op when S == 2
enum ConvKind { KIND_CONV = 0, KIND_S2 = 1, KIND_CT = 2 };
struct ConvL {
  int in0, in1;            // activation ids (in0 == -1: the network input, Cin = 1)
  int cin0, cin1, cout, dil, out;
  std::string name;        // module path, e.g. "enc_block1.conv_block.conv_layer1"
  int kind = KIND_CONV;
  int S = 1;               // KIND_CT: kernel size == stride
};
enum OpKind { OP_CONV = 0, OP_POOL = 1, OP_UP = 2, OP_S2D = 3, OP_D2S = 4 };
struct Op { OpKind kind; int layer; int src; int dst; };

inline long long align_up(long long v) { return (v + 255) & ~255LL; }

}  // namespace

struct UNetPlan {
  int input_ch, init_ch, max_ch, num_classes, output_stride, dtype;
  int strided = 0;         // 1: is_stride_conv and is_trans_conv (the reference only allows both or neither, unet.py:25)
  std::vector<Act> acts;
  std::vector<ConvL> convs;
  std::vector<Op> ops;
  int head_in = -1;
  std::map<std::string, int> named;
  // data-parallel overlap: ev[i] is recorded on the backward stream as soon as every parameter gradient of
  // conv layers >= ev_layer[i] (and of the head) is complete, so a communication stream can start reducing
  // that slice of the gradient buffer while the rest of the backward pass still runs (dp.py).
  std::vector<int> ev_layer;
  std::vector<cudaEvent_t> ev;
  // backward overlap: the weight-gradient kernels (tensor / L2 bound) run on an internal side stream while the
  // critical chain (BatchNorm backward -> dgrad -> pool/upsample backward, mostly HBM bound) continues on the
  // caller's stream. dY lives in kDyBufs rotating buffers guarded by events. PP_NO_OVERLAP=1 disables it.
  static constexpr int kDyBufs = 3;
  static constexpr int kMaxParts = 2;   // forward: statistics groups on separate streams
  mutable cudaStream_t side = nullptr, hi = nullptr;
  mutable cudaEvent_t dy_ready[kDyBufs] = {}, buf_free[kDyBufs] = {}, join = nullptr, hi_done = nullptr;
  mutable cudaStream_t part_stream[kMaxParts] = {};
  mutable cudaEvent_t fork = nullptr, stat_order[kMaxParts] = {}, part_done[kMaxParts] = {};
  mutable int overlap = -1;   // -1: not initialised, 0: off, 1: on
  // CUDA-graph replay of a whole pass (PP_GRAPHS=0 turns it off). A pass is ~200-260 launches over up to four internal
  // streams; when the SAME call (same pointers, shapes, flags) has been seen once eagerly, the next one is captured
  // (the internal fork / join events become graph edges) and later identical calls are ONE cudaGraphLaunch: no per-launch
  // host cost, shorter gaps between dependent kernels. PyTorch's caching allocator hands the same blocks back every
  // step, so in steady state every step hits; a call whose pointers differ simply runs eagerly (or gets its own graph).
  struct GraphEntry {
    std::vector<unsigned long long> key;
    cudaGraphExec_t exec = nullptr;
    long long launches = 0;
    int seen = 0;
    unsigned long long last_use = 0;
  };
  mutable std::vector<GraphEntry> graphs;
  mutable std::mutex graph_mu;
  mutable cudaStream_t cap_stream = nullptr;   // passes are captured on this stream (the caller's may be the legacy
                                               // default stream, which cannot capture) and replayed on the caller's
  mutable unsigned long long graph_clock = 0;

  int new_act(int C, int res) { acts.push_back({C, res}); return int(acts.size()) - 1; }

  int add_conv(const std::string& name, int in0, int in1, int cin0, int cin1, int cout, int dil, int res,
               int kind = KIND_CONV, int S = 1) {
    const int out = new_act(cout, res);
    ConvL c{in0, in1, cin0, cin1, cout, dil, out, name};
    c.kind = kind;
    c.S = S;
    convs.push_back(c);
    ops.push_back({OP_CONV, int(convs.size()) - 1, -1, out});
    return out;
  }
  int double_conv(const std::string& block, int in0, int in1, int cin0, int cin1, int cout, int dil, int res,
                  int kind1 = KIND_CONV) {
    const int a = add_conv(block + ".conv_block.conv_layer1", in0, in1, cin0, cin1, cout, dil, res, kind1);
    return add_conv(block + ".conv_block.conv_layer2", a, -1, cout, 0, cout, dil, res);
  }

  void build() {
    int ch[6];
    for (int k = 0; k < 6; ++k) ch[k] = std::min(max_ch, (1 << k) * init_ch);
    bool pool[6] = {false, true, true, true, false, false};
    int dil[6] = {1, 1, 1, 1, 1, 1};
    int scale5 = 1, scale4 = 1;
    if (output_stride == 32) { pool[4] = pool[5] = true; scale5 = scale4 = 2; }
    else if (output_stride == 16) { pool[4] = true; dil[5] = 2; scale4 = 2; }
    else { dil[4] = 2; dil[5] = 4; }
    int enc[6];
    int cur = -1, cur_c = input_ch, res = 1;
    for (int k = 0; k < 6; ++k) {
      int kind1 = KIND_CONV, cin = cur_c;
      if (pool[k] && !strided) {
        const int p = new_act(cur_c, res * 2);
        ops.push_back({OP_POOL, -1, cur, p});
        cur = p;
        res *= 2;
      } else if (pool[k]) {   // stride-2 first conv (unet.py:113-116) == stride-1 conv over the space-to-depth input
        const int p = new_act(4 * cur_c, res * 2);
        ops.push_back({OP_S2D, -1, cur, p});
        cur = p;
        res *= 2;
        kind1 = KIND_S2;
        cin = 4 * cur_c;
      }
      cur = double_conv("enc_block" + std::to_string(k + 1), cur, -1, cin, 0, ch[k], dil[k], res, kind1);
      cur_c = ch[k];
      enc[k] = cur;
      named["encoder/stage" + std::to_string(k + 1)] = cur;
    }
    // decoder: stage 5..1; skip = enc[stage-1]; scale per stage
    const int scales[5] = {scale5, scale4, 2, 2, 2};
    for (int i = 0; i < 5; ++i) {
      const int stage = 5 - i;
      const int skip = enc[stage - 1];
      int low = cur;
      if (strided) {   // ConvTranspose2d(lower_ch, skip_ch, S, S, bias=False) (unet.py:141), also for S == 1
        const int S = scales[i], skip_c = acts[skip].C;
        low = add_conv("dec_block" + std::to_string(stage) + ".up_samp", cur, -1, cur_c, 0, S * S * skip_c, 1, res,
                       KIND_CT, S);
        if (S == 2) {
          const int u = new_act(skip_c, res / 2);
          ops.push_back({OP_D2S, -1, low, u});
          low = u;
          res /= 2;
        }
        cur_c = skip_c;
      } else if (scales[i] > 1) {
        const int u = new_act(cur_c, res / scales[i]);
        ops.push_back({OP_UP, -1, cur, u});
        low = u;
        res /= scales[i];
      }
      cur = double_conv("dec_block" + std::to_string(stage), low, skip, cur_c, acts[skip].C, acts[skip].C, 1, res);
      cur_c = acts[skip].C;
      named["decoder/stage" + std::to_string(stage)] = cur;
    }
    head_in = cur;
  }
};

struct UNetLayout {
  long long total = 0;
  std::vector<long long> act_data, act_grad;          // per activation
  std::vector<long long> yraw, coef, sums, wf, wd;     // per conv layer
  std::vector<long long> wexp;                         // per conv layer: zero-embedded fp32 OIHW kernel (KIND_S2 / KIND_CT), else -1
  long long dwe_scratch = 0;                           // embedded weight gradient of one such layer (fp32 OIHW)
  long long dy_scratch = 0, dy_stride = 0, dwp_scratch = 0, dws_scratch = 0, dws_floats = 0, bsums = 0, bcoef = 0;
  long long sums_begin = 0, sums_bytes = 0, bsums_begin = 0, bsums_bytes = 0;
  std::vector<long long> bsums_layer;
};

static UNetLayout make_layout(const UNetPlan& pl, int N, int H, int W, int G) {
  UNetLayout L;
  const long long es = pl.dtype == PP_BF16 ? 2 : 4;
  long long off = 0;
  auto take = [&](long long bytes) { long long o = off; off += align_up(bytes); return o; };
  auto act_bytes = [&](const Act& a) { return static_cast<long long>(N) * (H / a.res) * (W / a.res) * a.C * es; };
  long long max_act = 0, max_w = 0, max_we = 0;
  int max_c = 0;
  for (const Act& a : pl.acts) {
    L.act_data.push_back(take(act_bytes(a)));
    max_act = std::max(max_act, act_bytes(a));
  }
  for (const Act& a : pl.acts) L.act_grad.push_back(take(act_bytes(a)));
  for (const ConvL& c : pl.convs) {
    L.yraw.push_back(take(act_bytes(pl.acts[c.out])));
    L.coef.push_back(take(sizeof(float) * 4 * G * c.cout));
    const long long wn = 9LL * c.cout * (c.cin0 + c.cin1);
    L.wf.push_back(take(wn * es));
    L.wd.push_back(take(wn * es));
    L.wexp.push_back(c.kind != KIND_CONV ? take(wn * 4) : -1);
    if (c.kind != KIND_CONV) max_we = std::max(max_we, wn);
    max_w = std::max(max_w, wn);
    max_c = std::max(max_c, c.cout);
  }
  // BatchNorm partial sums of every layer, forward ([replica][group][C][2]) and backward ([group][C][2]), in ONE
  // contiguous region each: a single memset per pass instead of one tiny memset per layer on the critical chain
  L.sums_begin = off;
  for (const ConvL& c : pl.convs) L.sums.push_back(take(sizeof(double) * 2 * G * c.cout * kStatReplicas));
  L.sums_bytes = off - L.sums_begin;
  L.bsums_begin = off;
  // (eval mode: up to kBnBwdReplicas replicas of [C][2] and, behind them, the 16-byte ticket of the one-pass backward)
  for (const ConvL& c : pl.convs)
    L.bsums_layer.push_back(take(sizeof(double) * 2 * c.cout * std::max(G, kBnBwdReplicas) + 16));
  L.bsums_bytes = off - L.bsums_begin;
  L.dy_scratch = take(max_act);
  L.dy_stride = align_up(max_act);
  for (int k = 1; k < UNetPlan::kDyBufs; ++k) take(max_act);
  L.dwp_scratch = take(max_w * 4);
  L.dws_floats = 2 * max_w;                 // split-K partial gradients of the wide layers (>= 2 splits each)
  L.dws_scratch = take(L.dws_floats * 4);
  if (max_we > 0) L.dwe_scratch = take(max_we * 4);
  L.bsums = take(sizeof(double) * 2 * G * max_c);
  L.bcoef = take(sizeof(float) * 2 * G * max_c);
  L.total = off;
  return L;
}

// Parameter pointer table, in conv-layer order then head:
//   per conv layer: [weight, bias, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.num_batches_tracked]
//   head: [weight, bias]
// Gradient pointer table: per conv layer [dweight, dbias, dgamma, dbeta]; head [dweight, dbias] (fp32, +=).
static constexpr int kParamsPerConv = 7;
static constexpr int kGradsPerConv = 4;

static int check_shape(const UNetPlan& pl, int N, int H, int W, int G) {
  const int div = pl.output_stride == 8 ? 8 : pl.output_stride;
  PP_REQUIRE(N > 0 && G > 0 && N % G == 0, "unet: batch %d not divisible into %d statistics groups", N, G);
  PP_REQUIRE(H % div == 0 && W % div == 0 && H >= div && W >= div,
             "unet: H=%d W=%d must be multiples of %d (maxpool stages)", H, W, div);
  return PP_OK;
}

// Eval-mode BatchNorm (running statistics) on the bf16 path is folded into the convolutions: the conv epilogue writes
// lrelu(acc * scale + shift) directly, there is no pre-BN tensor, no finalize / apply pass, and the backward pass is
// one pass per layer from the saved activation (ops.cu: bn_eval_coef_multi, bn_bwd_eval). PP_NO_EVAL_FUSION=1 keeps the
// unfused kernels (A/B experiments); the fp32 precision mode always uses them.
static bool eval_fusion(const UNetPlan& pl, int training) {
  static const int off = [] { const char* e = getenv("PP_NO_EVAL_FUSION"); return (e && e[0] == '1') ? 1 : 0; }();
  return !training && pl.dtype == PP_BF16 && !off;
}

// internal streams / events of the overlap machinery (see UNetPlan), created on first use on the current device
static int overlap_init(const UNetPlan& pl) {
  if (pl.overlap >= 0) return PP_OK;
  const char* off = getenv("PP_NO_OVERLAP");
  const int on = (off != nullptr && off[0] == '1') ? 0 : 1;
  if (on) {
    PP_CHECK_CUDA(cudaStreamCreateWithFlags(&pl.side, cudaStreamNonBlocking));
    const char* nohi = getenv("PP_NO_PRIORITY");
    if (nohi == nullptr || nohi[0] != '1') {
      int least = 0, greatest = 0;
      PP_CHECK_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
      PP_CHECK_CUDA(cudaStreamCreateWithPriority(&pl.hi, cudaStreamNonBlocking, greatest));
      PP_CHECK_CUDA(cudaEventCreateWithFlags(&pl.hi_done, cudaEventDisableTiming));
    }
    for (int k = 0; k < UNetPlan::kDyBufs; ++k) {
      PP_CHECK_CUDA(cudaEventCreateWithFlags(&pl.dy_ready[k], cudaEventDisableTiming));
      PP_CHECK_CUDA(cudaEventCreateWithFlags(&pl.buf_free[k], cudaEventDisableTiming));
    }
    PP_CHECK_CUDA(cudaEventCreateWithFlags(&pl.join, cudaEventDisableTiming));
    PP_CHECK_CUDA(cudaEventCreateWithFlags(&pl.fork, cudaEventDisableTiming));
    for (int k = 0; k < UNetPlan::kMaxParts; ++k) {
      if (k > 0) PP_CHECK_CUDA(cudaStreamCreateWithFlags(&pl.part_stream[k], cudaStreamNonBlocking));
      PP_CHECK_CUDA(cudaEventCreateWithFlags(&pl.stat_order[k], cudaEventDisableTiming));
      PP_CHECK_CUDA(cudaEventCreateWithFlags(&pl.part_done[k], cudaEventDisableTiming));
    }
  }
  pl.overlap = on;
  return PP_OK;
}

static int unet_forward_eager(const UNetPlan& pl, const float* x, void* const* params, void* ws, int N, int H, int W,
                              int G, int training, float* logits, cudaStream_t s) {
  int rc = check_shape(pl, N, H, W, G);
  if (rc) return rc;
  const UNetLayout L = make_layout(pl, N, H, W, G);
  char* base = static_cast<char*>(ws);
  const int dt = pl.dtype;
  {  // fp32 OIHW master weights -> packed forward / dgrad operands of every tensor-core layer, one launch
    std::vector<const float*> pw;
    std::vector<void*> pf, pd;
    std::vector<int> co, ci;
    for (size_t l = 0; l < pl.convs.size(); ++l) {
      const ConvL& c = pl.convs[l];
      if (c.in0 < 0) continue;
      if (c.kind != KIND_CONV) {   // rebuild the zero-embedded fp32 kernel from the master weights, then pack that
        float* we = reinterpret_cast<float*>(base + L.wexp[l]);
        const float* w = static_cast<const float*>(params[l * kParamsPerConv]);
        rc = c.kind == KIND_S2 ? embed_s2_weight(w, we, c.cout, c.cin0 / 4, s)
                               : embed_ct_weight(w, we, c.cin0, c.cout / (c.S * c.S), c.S, s);
        if (rc) return rc;
        pw.push_back(we);
      } else
      pw.push_back(static_cast<const float*>(params[l * kParamsPerConv]));
      pf.push_back(base + L.wf[l]);
      pd.push_back(base + L.wd[l]);
      co.push_back(c.cout);
      ci.push_back(c.cin0 + c.cin1);
    }
    rc = pack_weights_multi(dt, int(pw.size()), pw.data(), pf.data(), pd.data(), co.data(), ci.data(), s);
    if (rc) return rc;
    if (training) PP_CHECK_CUDA(cudaMemsetAsync(base + L.sums_begin, 0, L.sums_bytes, s));   // all layers' statistics
  }
  const bool fuse_eval = eval_fusion(pl, training);
  if (fuse_eval) {   // scale / shift of every BatchNorm layer from the running statistics, one launch
    std::vector<const float*> g, b, rm, rv, bias;
    std::vector<float*> cf;
    std::vector<int> cc;
    for (size_t l = 0; l < pl.convs.size(); ++l) {
      const ConvL& c = pl.convs[l];
      if (c.kind == KIND_CT) continue;
      void* const* pp = params + l * kParamsPerConv;
      bias.push_back(static_cast<const float*>(pp[1]));
      g.push_back(static_cast<const float*>(pp[2]));
      b.push_back(static_cast<const float*>(pp[3]));
      rm.push_back(static_cast<const float*>(pp[4]));
      rv.push_back(static_cast<const float*>(pp[5]));
      cf.push_back(reinterpret_cast<float*>(base + L.coef[l]));
      cc.push_back(c.cout);
    }
    rc = bn_eval_coef_multi(int(cc.size()), g.data(), b.data(), rm.data(), rv.data(), bias.data(), cf.data(), cc.data(),
                            1e-5f, s);
    if (rc) return rc;
  }
  // Statistics groups (the weak and the strong branch of the siamese step) are independent until the losses, so
  // with overlap enabled every group runs on its own stream: the HBM-bound BatchNorm / pool / upsample kernels of one
  // branch overlap the tensor-core kernels of the other. Only the running-statistics update is ordered (weak, then
  // strong, as two reference forward passes would do): group g's finalize waits for group g-1's of the same layer.
  rc = overlap_init(pl);
  if (rc) return rc;
  static int fwd_parts_on = -1;
  if (fwd_parts_on < 0) {
    const char* e = getenv("PP_FWD_PARTS");
    fwd_parts_on = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  const int parts = (pl.overlap == 1 && fwd_parts_on && G > 1 && G <= UNetPlan::kMaxParts) ? G : 1;
  const int Np = N / parts, Gp = G / parts;   // images and statistics groups per part
  const long long es = dt == PP_BF16 ? 2 : 4;
  cudaStream_t st[UNetPlan::kMaxParts];
  st[0] = s;
  for (int k = 1; k < parts; ++k) {
    st[k] = pl.part_stream[k];
    PP_CHECK_CUDA(cudaEventRecord(pl.fork, s));
    PP_CHECK_CUDA(cudaStreamWaitEvent(st[k], pl.fork, 0));
  }
  auto act_part = [&](long long off, const Act& a, int k) -> char* {   // images [k*Np, (k+1)*Np) of an activation
    return base + off + static_cast<long long>(k) * Np * (H / a.res) * (W / a.res) * a.C * es;
  };
  for (const Op& op : pl.ops) {
    for (int k = 0; k < parts; ++k) {
      cudaStream_t sk = st[k];
      if (op.kind == OP_CONV) {
        const ConvL& c = pl.convs[op.layer];
        void* const* pp = params + op.layer * kParamsPerConv;
        const Act& ao = pl.acts[c.out];
        const int h = H / ao.res, w = W / ao.res;
        void* yraw = act_part(L.yraw[op.layer], ao, k);
        const long long Pg = static_cast<long long>(N / G) * h * w;
        if (c.kind == KIND_CT) {   // plain linear layer: no bias, no BatchNorm, no activation
          const void* x0 = act_part(L.act_data[c.in0], pl.acts[c.in0], k);
          void* y = act_part(L.act_data[c.out], ao, k);
          const void* wf = base + L.wf[op.layer];
          rc = dt == PP_BF16 ? conv3x3_tc(x0, c.cin0, nullptr, 0, wf, nullptr, y, c.cout, 0, nullptr, 0, 0, Np, h, w, 1, sk)
                             : conv3x3_simt(dt, x0, c.cin0, nullptr, 0, wf, nullptr, y, c.cout, 0, nullptr, 0, 0, Np, h, w,
                                            1, sk);
          if (rc) return rc;
          continue;
        }
        if (fuse_eval) {   // conv + running-statistics BatchNorm + LeakyReLU in one kernel: writes the activation
          const float* cf = reinterpret_cast<const float*>(base + L.coef[op.layer]);
          const ConvAffine af{cf, cf + c.cout, 0.01f};
          void* a_out = act_part(L.act_data[c.out], ao, k);
          if (c.in0 < 0) {
            rc = first_conv_fwd(dt, x + static_cast<long long>(k) * Np * pl.input_ch * H * W,
                                static_cast<const float*>(pp[0]), nullptr, a_out, Np, h, w, c.cout, sk, pl.input_ch, &af);
          } else {
            const void* x0 = act_part(L.act_data[c.in0], pl.acts[c.in0], k);
            const void* x1 = c.in1 >= 0 ? act_part(L.act_data[c.in1], pl.acts[c.in1], k) : nullptr;
            rc = conv3x3_tc(x0, c.cin0, x1, c.cin1, base + L.wf[op.layer], nullptr, a_out, c.cout, 0, nullptr, 0, 0, Np,
                            h, w, c.dil, sk, nullptr, Gp, &af);
          }
          if (rc) return rc;
          continue;
        }
        bool fused_stats = false;
        if (dt == PP_BF16 && c.in0 >= 0) {
          // batch statistics ride in the conv epilogue when a pixel tile never straddles two statistics groups
          int bw_ = 1, bh_ = 1;
          while (bw_ < w && bw_ < 128) bw_ <<= 1;
          while (bh_ < h && bw_ * bh_ < 128) bh_ <<= 1;
          const int bn_ = 128 / (bw_ * bh_);
          fused_stats = training && (bn_ == 1 || (N / G) % bn_ == 0);
        }
        const int reps = fused_stats ? kStatReplicas : 1;
        // sums: [replica][group][C][2] per part; coef: [group][4][C]
        double* sums = reinterpret_cast<double*>(base + L.sums[op.layer]) + static_cast<long long>(k) * Gp * reps * c.cout * 2;
        float* coef = reinterpret_cast<float*>(base + L.coef[op.layer]) + static_cast<long long>(k) * Gp * 4 * c.cout;
        if (c.in0 < 0) {
          rc = first_conv_fwd(dt, x + static_cast<long long>(k) * Np * pl.input_ch * H * W,
                              static_cast<const float*>(pp[0]), static_cast<const float*>(pp[1]), yraw, Np, h, w, c.cout,
                              sk, pl.input_ch);
        } else {
          void* wf = base + L.wf[op.layer];
          const void* x0 = act_part(L.act_data[c.in0], pl.acts[c.in0], k);
          const void* x1 = c.in1 >= 0 ? act_part(L.act_data[c.in1], pl.acts[c.in1], k) : nullptr;
          if (dt == PP_BF16) {
            rc = conv3x3_tc(x0, c.cin0, x1, c.cin1, wf, static_cast<const float*>(pp[1]), yraw, c.cout, 0, nullptr, 0,
                            0, Np, h, w, c.dil, sk, fused_stats ? sums : nullptr, Gp);
          } else {
            rc = conv3x3_simt(dt, x0, c.cin0, x1, c.cin1, wf, static_cast<const float*>(pp[1]), yraw, c.cout, 0,
                              nullptr, 0, 0, Np, h, w, c.dil, sk);
          }
        }
        if (rc) return rc;
        if (training && !fused_stats) {
          rc = bn_stats(dt, yraw, sums, Gp, Pg, c.cout, sk);
          if (rc) return rc;
        }
        if (k > 0) PP_CHECK_CUDA(cudaStreamWaitEvent(sk, pl.stat_order[k - 1], 0));   // running stats: group order
        rc = bn_finalize(sums, static_cast<const float*>(pp[2]), static_cast<const float*>(pp[3]),
                         static_cast<float*>(pp[4]), static_cast<float*>(pp[5]), static_cast<long long*>(pp[6]), coef,
                         Gp, Pg, c.cout, training, 1e-5f, 0.1f, sk, reps);
        if (rc) return rc;
        if (parts > 1 && k + 1 < parts) PP_CHECK_CUDA(cudaEventRecord(pl.stat_order[k], sk));
        rc = bn_apply(dt, yraw, coef, act_part(L.act_data[c.out], ao, k), Gp, Pg, c.cout, 0.01f, sk);
        if (rc) return rc;
      } else if (op.kind == OP_POOL) {
        const Act& as = pl.acts[op.src];
        rc = maxpool_fwd(dt, act_part(L.act_data[op.src], as, k), act_part(L.act_data[op.dst], pl.acts[op.dst], k), Np,
                         H / as.res, W / as.res, as.C, sk);
        if (rc) return rc;
      } else if (op.kind == OP_S2D) {
        const Act& as = pl.acts[op.src];
        const Act& ad = pl.acts[op.dst];
        rc = space_to_depth(dt, act_part(L.act_data[op.src], as, k), act_part(L.act_data[op.dst], ad, k), Np, H / ad.res,
                            W / ad.res, as.C, sk);
        if (rc) return rc;
      } else if (op.kind == OP_D2S) {
        const Act& as = pl.acts[op.src];
        const Act& ad = pl.acts[op.dst];
        rc = depth_to_space(dt, act_part(L.act_data[op.src], as, k), act_part(L.act_data[op.dst], ad, k), Np, H / as.res,
                            W / as.res, ad.C, 0, sk);
        if (rc) return rc;
      } else {
        const Act& as = pl.acts[op.src];
        const Act& ad = pl.acts[op.dst];
        rc = upsample_nhwc_fwd(dt, act_part(L.act_data[op.src], as, k), act_part(L.act_data[op.dst], ad, k), Np,
                               H / as.res, W / as.res, H / ad.res, W / ad.res, as.C, sk);
        if (rc) return rc;
      }
    }
  }
  void* const* hp = params + pl.convs.size() * kParamsPerConv;
  const Act& ah = pl.acts[pl.head_in];
  const int hh = H / ah.res, hw = W / ah.res;
  for (int k = 0; k < parts; ++k) {
    rc = head_fwd(dt, act_part(L.act_data[pl.head_in], ah, k), static_cast<const float*>(hp[0]),
                  static_cast<const float*>(hp[1]), logits + static_cast<long long>(k) * Np * pl.num_classes * hh * hw,
                  static_cast<long long>(Np) * hh * hw, hh * hw, ah.C, pl.num_classes, st[k]);
    if (rc) return rc;
  }
  for (int k = 1; k < parts; ++k) {   // the caller's stream owns the result
    PP_CHECK_CUDA(cudaEventRecord(pl.part_done[k], st[k]));
    PP_CHECK_CUDA(cudaStreamWaitEvent(s, pl.part_done[k], 0));
  }
  return PP_OK;
}

// dfeat: optional external gradients w.r.t. named activations (NHWC, activation dtype), e.g. the
// aux path's gradient into encoder/stage5 and encoder/stage6. dfeat_act[i] = activation id.
static int unet_backward_eager(const UNetPlan& pl, const float* x, void* const* params, void* ws, int N, int H, int W,
                               int G, int training, const float* dlogits, int n_dfeat, const int* dfeat_act,
                               const void* const* dfeat, float* const* grads, cudaStream_t user_stream) {
  int rc = check_shape(pl, N, H, W, G);
  if (rc) return rc;
  rc = overlap_init(pl);
  if (rc) return rc;
  // With overlap on, the critical chain (BatchNorm backward -> dgrad -> pool / upsample backward) runs on an internal
  // HIGH-priority stream forked from the caller's stream, the weight gradients on an internal default-priority stream:
  // when both have blocks pending, the block scheduler serves the chain first, so the HBM-bound chain kernels are not
  // starved by the co-running wgrad CTAs. The caller's stream joins both at the end.
  cudaStream_t s = user_stream;
  if (pl.overlap == 1 && pl.hi != nullptr) {
    PP_CHECK_CUDA(cudaEventRecord(pl.fork, user_stream));
    PP_CHECK_CUDA(cudaStreamWaitEvent(pl.hi, pl.fork, 0));
    s = pl.hi;
  }
  const UNetLayout L = make_layout(pl, N, H, W, G);
  char* base = static_cast<char*>(ws);
  const int dt = pl.dtype;
  const long long es = dt == PP_BF16 ? 2 : 4;
  std::vector<char> written(pl.acts.size(), 0);
  auto act_bytes = [&](const Act& a) { return static_cast<long long>(N) * (H / a.res) * (W / a.res) * a.C * es; };

  // head
  {
    void* const* hp = params + pl.convs.size() * kParamsPerConv;
    float* const* hg = grads + pl.convs.size() * kGradsPerConv;
    const Act& ah = pl.acts[pl.head_in];
    const int hh = H / ah.res, hw = W / ah.res;
    rc = head_bwd(dt, dlogits, base + L.act_data[pl.head_in], static_cast<const float*>(hp[0]),
                  base + L.act_grad[pl.head_in], hg[0], hg[1], static_cast<long long>(N) * hh * hw, hh * hw, ah.C,
                  pl.num_classes, s);
    if (rc) return rc;
    written[pl.head_in] = 1;
  }
  for (int i = 0; i < n_dfeat; ++i) {
    const int a = dfeat_act[i];
    PP_REQUIRE(a >= 0 && a < int(pl.acts.size()) && !written[a], "unet_backward: bad external gradient target %d", a);
    PP_CHECK_CUDA(cudaMemcpyAsync(base + L.act_grad[a], dfeat[i], act_bytes(pl.acts[a]), cudaMemcpyDeviceToDevice, s));
    written[a] = 1;
  }

  const bool ov = pl.overlap == 1;
  const bool fuse_eval = eval_fusion(pl, training);
  cudaStream_t ws_ = ov ? pl.side : s;       // stream of the weight-gradient kernels
  PP_CHECK_CUDA(cudaMemsetAsync(base + L.bsums_begin, 0, L.bsums_bytes, s));   // BatchNorm-backward sums of all layers
  bool buf_used[UNetPlan::kDyBufs] = {false, false, false};
  int nbuf = 0;

  for (int oi = int(pl.ops.size()) - 1; oi >= 0; --oi) {
    const Op& op = pl.ops[oi];
    if (op.kind == OP_CONV) {
      const ConvL& c = pl.convs[op.layer];
      float* const* gg = grads + op.layer * kGradsPerConv;
      const Act& ao = pl.acts[c.out];
      const int h = H / ao.res, w = W / ao.res;
      const long long Pg = static_cast<long long>(N / G) * h * w;
      if (!written[c.out]) {  // no consumer produced a gradient: it is zero
        PP_CHECK_CUDA(cudaMemsetAsync(base + L.act_grad[c.out], 0, act_bytes(ao), s));
        written[c.out] = 1;
      }
      const int kb = ov ? (nbuf++ % UNetPlan::kDyBufs) : 0;
      void* dy = base + L.dy_scratch + kb * L.dy_stride;
      if (ov && buf_used[kb]) PP_CHECK_CUDA(cudaStreamWaitEvent(s, pl.buf_free[kb], 0));   // its last reader is done
      if (c.kind == KIND_CT) {
        dy = base + L.act_grad[c.out];   // plain layer: the activation gradient IS the conv-output gradient
      } else if (fuse_eval) {   // one pass from the saved activation (no pre-BN tensor exists in this mode)
        rc = bn_bwd_eval(dt, base + L.act_grad[c.out], base + L.act_data[c.out],
                         reinterpret_cast<const float*>(base + L.coef[op.layer]),
                         reinterpret_cast<double*>(base + L.bsums_layer[op.layer]), gg[2], gg[3], gg[1], dy,
                         static_cast<long long>(N) * h * w, c.cout, 0.01f, s, bn_bwd_replicas(c.cout));
        if (rc) return rc;
      } else {
        rc = bn_bwd(dt, base + L.act_grad[c.out], base + L.yraw[op.layer],
                    reinterpret_cast<const float*>(base + L.coef[op.layer]),
                    reinterpret_cast<double*>(base + L.bsums_layer[op.layer]), reinterpret_cast<float*>(base + L.bcoef),
                    gg[2], gg[3], gg[1], dy, G, Pg, c.cout, training, 0.01f, s, /*sums_zeroed=*/true);
        if (rc) return rc;
      }
      if (ov) {
        PP_CHECK_CUDA(cudaEventRecord(pl.dy_ready[kb], s));
        PP_CHECK_CUDA(cudaStreamWaitEvent(ws_, pl.dy_ready[kb], 0));
      }
      auto grads_ready = [&]() -> int {   // all parameter gradients of layers >= op.layer are enqueued
        for (size_t e = 0; e < pl.ev_layer.size(); ++e)
          if (pl.ev_layer[e] == op.layer) PP_CHECK_CUDA(cudaEventRecord(pl.ev[e], ws_));   // (never inside a capture)
        if (ov) {
          PP_CHECK_CUDA(cudaEventRecord(pl.buf_free[kb], ws_));
          buf_used[kb] = true;
        }
        return PP_OK;
      };
      if (c.in0 < 0) {
        rc = first_conv_wgrad(dt, dy, x, gg[0], N, h, w, c.cout, ws_, pl.input_ch);
        if (rc) return rc;
        rc = grads_ready();
        if (rc) return rc;
        continue;
      }
      const int ctot = c.cin0 + c.cin1;
      float* dwp = reinterpret_cast<float*>(base + L.dwp_scratch);
      const void* x0 = base + L.act_data[c.in0];
      const void* x1 = c.in1 >= 0 ? base + L.act_data[c.in1] : nullptr;
      // embedded layers: the OIHW gradient of the zero-embedded kernel lands in a scratch, then is gathered back
      float* g_oihw = gg[0];
      if (c.kind != KIND_CONV) {
        g_oihw = reinterpret_cast<float*>(base + L.dwe_scratch);
        PP_CHECK_CUDA(cudaMemsetAsync(g_oihw, 0, sizeof(float) * 9 * c.cout * ctot, ws_));
      }
      if (dt == PP_BF16) {
        // wide sources accumulate straight into the OIHW gradient; narrow ones go through the packed scratch
        if (conv3x3_wgrad_tc_uses_scratch(c.cout, c.cin0, c.cin1, w, c.dil))
          PP_CHECK_CUDA(cudaMemsetAsync(dwp, 0, sizeof(float) * 9 * c.cout * ctot, ws_));
        rc = conv3x3_wgrad_tc(dy, c.cout, x0, c.cin0, x1, c.cin1, dwp, g_oihw, N, h, w, c.dil, ws_,
                              reinterpret_cast<float*>(base + L.dws_scratch), L.dws_floats);
        if (rc) return rc;
      } else {
        PP_CHECK_CUDA(cudaMemsetAsync(dwp, 0, sizeof(float) * 9 * c.cout * ctot, ws_));
        rc = conv3x3_wgrad_simt(dt, dy, c.cout, x0, c.cin0, x1, c.cin1, dwp, N, h, w, c.dil, ws_);
        if (rc) return rc;
        rc = unpack_wgrad(dwp, g_oihw, c.cout, ctot, 1, ws_);
        if (rc) return rc;
      }
      if (c.kind == KIND_S2) rc = collapse_s2_wgrad(g_oihw, gg[0], c.cout, c.cin0 / 4, ws_);
      else if (c.kind == KIND_CT) rc = collapse_ct_wgrad(g_oihw, gg[0], c.cin0, c.cout / (c.S * c.S), c.S, ws_);
      if (rc) return rc;
      rc = grads_ready();
      if (rc) return rc;
      // dgrad: forward kernel on the flipped/transposed pack, scattered to the two sources
      void* g0 = base + L.act_grad[c.in0];
      void* g1 = c.in1 >= 0 ? base + L.act_grad[c.in1] : nullptr;
      const int acc0 = written[c.in0], acc1 = c.in1 >= 0 ? written[c.in1] : 0;
      const void* wd = base + L.wd[op.layer];
      if (dt == PP_BF16)
        rc = conv3x3_tc(dy, c.cout, nullptr, 0, wd, nullptr, g0, c.cin0, acc0, g1, c.cin1, acc1, N, h, w, c.dil, s);
      else
        rc = conv3x3_simt(dt, dy, c.cout, nullptr, 0, wd, nullptr, g0, c.cin0, acc0, g1, c.cin1, acc1, N, h, w, c.dil,
                          s);
      if (rc) return rc;
      written[c.in0] = 1;
      if (c.in1 >= 0) written[c.in1] = 1;
    } else if (op.kind == OP_POOL) {
      const Act& as = pl.acts[op.src];
      if (!written[op.dst]) {
        PP_CHECK_CUDA(cudaMemsetAsync(base + L.act_grad[op.dst], 0, act_bytes(pl.acts[op.dst]), s));
        written[op.dst] = 1;
      }
      rc = maxpool_bwd(dt, base + L.act_data[op.src], base + L.act_grad[op.dst], base + L.act_grad[op.src], N,
                       H / as.res, W / as.res, as.C, written[op.src], s);
      if (rc) return rc;
      written[op.src] = 1;
    } else if (op.kind == OP_S2D || op.kind == OP_D2S) {
      const Act& as = pl.acts[op.src];
      const Act& ad = pl.acts[op.dst];
      if (!written[op.dst]) {
        PP_CHECK_CUDA(cudaMemsetAsync(base + L.act_grad[op.dst], 0, act_bytes(ad), s));
        written[op.dst] = 1;
      }
      if (op.kind == OP_S2D)   // forward gathered big -> small: the gradient scatters small -> big
        rc = depth_to_space(dt, base + L.act_grad[op.dst], base + L.act_grad[op.src], N, H / ad.res, W / ad.res, as.C,
                            written[op.src], s);
      else {
        PP_REQUIRE(!written[op.src], "unet_backward: transposed-conv output has a second consumer");
        rc = space_to_depth(dt, base + L.act_grad[op.dst], base + L.act_grad[op.src], N, H / as.res, W / as.res, ad.C, s);
      }
      if (rc) return rc;
      written[op.src] = 1;
    } else {
      const Act& as = pl.acts[op.src];
      const Act& ad = pl.acts[op.dst];
      if (!written[op.dst]) {
        PP_CHECK_CUDA(cudaMemsetAsync(base + L.act_grad[op.dst], 0, act_bytes(ad), s));
        written[op.dst] = 1;
      }
      rc = upsample_nhwc_bwd(dt, base + L.act_grad[op.dst], base + L.act_grad[op.src], N, H / as.res, W / as.res,
                             H / ad.res, W / ad.res, as.C, written[op.src], s);
      if (rc) return rc;
      written[op.src] = 1;
    }
  }
  if (ov) {   // the caller's stream owns the result: every gradient is complete when its work is
    PP_CHECK_CUDA(cudaEventRecord(pl.join, pl.side));
    PP_CHECK_CUDA(cudaStreamWaitEvent(user_stream, pl.join, 0));
    if (s != user_stream) {
      PP_CHECK_CUDA(cudaEventRecord(pl.hi_done, s));
      PP_CHECK_CUDA(cudaStreamWaitEvent(user_stream, pl.hi_done, 0));
    }
  }
  return PP_OK;
}

// ---- CUDA-graph replay of whole passes (see UNetPlan::GraphEntry) ---------------------------------------------------
static std::atomic<long long> g_graph_replays{0};
long long unet_graph_replays() { return g_graph_replays.load(); }

static bool graphs_enabled() {
  static const int on = [] { const char* e = getenv("PP_GRAPHS"); return (e && e[0] == '0') ? 0 : 1; }();
  return on != 0;
}

template <typename Fn>
static int run_with_graph(const UNetPlan& pl, std::vector<unsigned long long>&& key, cudaStream_t s, Fn&& eager) {
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (!graphs_enabled() || prof_enabled() || cudaStreamIsCapturing(s, &cap) != cudaSuccess ||
      cap != cudaStreamCaptureStatusNone)
    return eager(s);
  cudaGraphExec_t exec = nullptr;
  long long launches = 0;
  bool capture = false;
  {
    std::lock_guard<std::mutex> lk(pl.graph_mu);
    UNetPlan::GraphEntry* hit = nullptr;
    for (auto& g : pl.graphs)
      if (g.key == key) { hit = &g; break; }
    if (hit == nullptr) {
      if (pl.graphs.size() >= 12) {   // evict the least recently used entry
        size_t lru = 0;
        for (size_t i = 1; i < pl.graphs.size(); ++i)
          if (pl.graphs[i].last_use < pl.graphs[lru].last_use) lru = i;
        if (pl.graphs[lru].exec != nullptr) cudaGraphExecDestroy(pl.graphs[lru].exec);
        pl.graphs.erase(pl.graphs.begin() + lru);
      }
      UNetPlan::GraphEntry g;
      g.key = key;
      g.seen = 1;
      g.last_use = ++pl.graph_clock;
      pl.graphs.push_back(std::move(g));
    } else {
      hit->last_use = ++pl.graph_clock;
      if (hit->exec != nullptr) { exec = hit->exec; launches = hit->launches; }
      else if (hit->seen >= 1) capture = true;   // seen once eagerly (kernel choices are measured by now): capture this one
      ++hit->seen;
    }
  }
  static const int debug = [] { const char* e = getenv("PP_GRAPHS_DEBUG"); return (e && e[0] == '1') ? 1 : 0; }();
  if (debug) fprintf(stderr, "pp graph pass %llu: %s\n", key[0], exec != nullptr ? "replay" : (capture ? "capture" : "eager"));
  if (exec != nullptr) {
    PP_CHECK_CUDA(cudaGraphLaunch(exec, s));
    count_launches(static_cast<int>(launches));
    g_graph_replays.fetch_add(1);
    return PP_OK;
  }
  if (!capture) return eager(s);
  const long long l0 = launch_count();
  {
    std::lock_guard<std::mutex> lk(pl.graph_mu);
    if (pl.cap_stream == nullptr) PP_CHECK_CUDA(cudaStreamCreateWithFlags(&pl.cap_stream, cudaStreamNonBlocking));
  }
  if (cudaStreamBeginCapture(pl.cap_stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
    cudaGetLastError();
    return eager(s);
  }
  const int rc = eager(pl.cap_stream);
  cudaGraph_t graph = nullptr;
  const cudaError_t ce = cudaStreamEndCapture(pl.cap_stream, &graph);
  if (rc != PP_OK || ce != cudaSuccess || graph == nullptr) {
    if (graph != nullptr) cudaGraphDestroy(graph);
    if (rc != PP_OK) return rc;
    cudaGetLastError();
    return eager(s);   // not capturable here: run it the ordinary way
  }
  cudaGraphExec_t ex = nullptr;
  if (cudaGraphInstantiate(&ex, graph, 0) != cudaSuccess || ex == nullptr) {
    cudaGraphDestroy(graph);
    cudaGetLastError();
    return eager(s);
  }
  cudaGraphDestroy(graph);
  launches = launch_count() - l0;
  {
    std::lock_guard<std::mutex> lk(pl.graph_mu);
    bool stored = false;
    for (auto& g : pl.graphs)
      if (g.key == key && g.exec == nullptr) { g.exec = ex; g.launches = launches; stored = true; break; }
    if (!stored) { cudaGraphExecDestroy(ex); return eager(s); }
  }
  PP_CHECK_CUDA(cudaGraphLaunch(ex, s));   // the captured launches were recorded, not run
  return PP_OK;
}

static void key_push_ptrs(std::vector<unsigned long long>* key, void* const* p, size_t n) {
  unsigned long long h = 1469598103934665603ULL;   // FNV-1a over the pointer table
  for (size_t i = 0; i < n; ++i) {
    h ^= reinterpret_cast<unsigned long long>(p[i]);
    h *= 1099511628211ULL;
  }
  key->push_back(h);
  key->push_back(n);
}

int unet_forward(const UNetPlan& pl, const float* x, void* const* params, void* ws, int N, int H, int W, int G,
                 int training, float* logits, cudaStream_t s) {
  std::vector<unsigned long long> key = {1ULL, reinterpret_cast<unsigned long long>(x), reinterpret_cast<unsigned long long>(ws),
                                         reinterpret_cast<unsigned long long>(logits), static_cast<unsigned long long>(N),
                                         static_cast<unsigned long long>(H), static_cast<unsigned long long>(W),
                                         static_cast<unsigned long long>(G), static_cast<unsigned long long>(training)};
  key_push_ptrs(&key, params, pl.convs.size() * kParamsPerConv + 2);
  return run_with_graph(pl, std::move(key), s, [&](cudaStream_t st) {
    return unet_forward_eager(pl, x, params, ws, N, H, W, G, training, logits, st);
  });
}

int unet_backward(const UNetPlan& pl, const float* x, void* const* params, void* ws, int N, int H, int W, int G,
                  int training, const float* dlogits, int n_dfeat, const int* dfeat_act, const void* const* dfeat,
                  float* const* grads, cudaStream_t user_stream) {
  auto eager = [&](cudaStream_t st) {
    return unet_backward_eager(pl, x, params, ws, N, H, W, G, training, dlogits, n_dfeat, dfeat_act, dfeat, grads, st);
  };
  // the data-parallel gradient events are recorded for streams OUTSIDE this call: keep such passes out of graphs
  if (!pl.ev.empty()) return eager(user_stream);
  std::vector<unsigned long long> key = {2ULL, reinterpret_cast<unsigned long long>(x), reinterpret_cast<unsigned long long>(ws),
                                         reinterpret_cast<unsigned long long>(dlogits), static_cast<unsigned long long>(N),
                                         static_cast<unsigned long long>(H), static_cast<unsigned long long>(W),
                                         static_cast<unsigned long long>(G), static_cast<unsigned long long>(training),
                                         static_cast<unsigned long long>(n_dfeat)};
  for (int i = 0; i < n_dfeat; ++i) {
    key.push_back(static_cast<unsigned long long>(dfeat_act[i]));
    key.push_back(reinterpret_cast<unsigned long long>(dfeat[i]));
  }
  key_push_ptrs(&key, params, pl.convs.size() * kParamsPerConv + 2);
  key_push_ptrs(&key, reinterpret_cast<void* const*>(grads), pl.convs.size() * kGradsPerConv + 2);
  return run_with_graph(pl, std::move(key), user_stream, eager);
}

UNetPlan* unet_create(int input_ch, int init_ch, int max_ch, int num_classes, int output_stride, int dtype,
                      int strided) {
  if (input_ch < 1 || input_ch > 16) { set_error("unet: input_ch=%d unsupported (1..16)", input_ch); return nullptr; }
  if (init_ch % 32 != 0 || max_ch % 32 != 0 || init_ch <= 0 || max_ch < init_ch) {
    set_error("unet: init_ch=%d / max_ch=%d must be positive multiples of 32", init_ch, max_ch);
    return nullptr;
  }
  if (init_ch != 32 && init_ch != 64) { set_error("unet: init_ch=%d unsupported (1x1 head handles 32 or 64)", init_ch); return nullptr; }
  if (output_stride != 8 && output_stride != 16 && output_stride != 32) { set_error("unet: output_stride=%d", output_stride); return nullptr; }
  if (num_classes < 1 || num_classes > 8) { set_error("unet: num_classes=%d unsupported (1..8)", num_classes); return nullptr; }
  if (dtype != PP_F32 && dtype != PP_BF16) { set_error("unet: bad dtype %d", dtype); return nullptr; }
  UNetPlan* pl = new UNetPlan();
  pl->input_ch = input_ch; pl->init_ch = init_ch; pl->max_ch = max_ch; pl->num_classes = num_classes;
  pl->output_stride = output_stride; pl->dtype = dtype;
  pl->strided = strided ? 1 : 0;
  pl->build();
  return pl;
}
void unet_destroy(UNetPlan* pl) {
  if (pl) {
    for (auto& g : pl->graphs)
      if (g.exec != nullptr) cudaGraphExecDestroy(g.exec);
    if (pl->cap_stream) cudaStreamDestroy(pl->cap_stream);
    for (cudaEvent_t e : pl->ev) cudaEventDestroy(e);
    for (int k = 0; k < UNetPlan::kDyBufs; ++k) {
      if (pl->dy_ready[k]) cudaEventDestroy(pl->dy_ready[k]);
      if (pl->buf_free[k]) cudaEventDestroy(pl->buf_free[k]);
    }
    if (pl->join) cudaEventDestroy(pl->join);
    if (pl->hi_done) cudaEventDestroy(pl->hi_done);
    if (pl->hi) cudaStreamDestroy(pl->hi);
    if (pl->side) cudaStreamDestroy(pl->side);
    if (pl->fork) cudaEventDestroy(pl->fork);
    for (int k = 0; k < UNetPlan::kMaxParts; ++k) {
      if (pl->stat_order[k]) cudaEventDestroy(pl->stat_order[k]);
      if (pl->part_done[k]) cudaEventDestroy(pl->part_done[k]);
      if (pl->part_stream[k]) cudaStreamDestroy(pl->part_stream[k]);
    }
  }
  delete pl;
}
int unet_set_grad_events(UNetPlan* pl, int n, const int* layers) {
  PP_REQUIRE(n >= 0 && (n == 0 || layers != nullptr), "unet_set_grad_events: bad arguments");
  for (int i = 0; i < n; ++i)
    PP_REQUIRE(layers[i] >= 0 && layers[i] < int(pl->convs.size()), "unet_set_grad_events: bad layer %d", layers[i]);
  for (cudaEvent_t e : pl->ev) cudaEventDestroy(e);
  pl->ev.clear();
  pl->ev_layer.assign(layers, layers + n);
  for (int i = 0; i < n; ++i) {
    cudaEvent_t e;
    PP_CHECK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    pl->ev.push_back(e);
  }
  return PP_OK;
}
int unet_wait_grad_event(const UNetPlan* pl, int i, cudaStream_t s) {
  PP_REQUIRE(i >= 0 && i < int(pl->ev.size()), "unet_wait_grad_event: bad event index %d", i);
  PP_CHECK_CUDA(cudaStreamWaitEvent(s, pl->ev[i], 0));
  return PP_OK;
}
int unet_num_convs(const UNetPlan* pl) { return int(pl->convs.size()); }
int unet_conv_info(const UNetPlan* pl, int layer, int* cin, int* cout, int* dil, const char** name) {
  PP_REQUIRE(layer >= 0 && layer < int(pl->convs.size()), "unet_conv_info: bad layer %d", layer);
  const ConvL& c = pl->convs[layer];
  // the PARAMETER's channels: a stride-2 layer reads 4 C space-to-depth channels, a transposed conv writes S*S*Cout
  *cin = c.in0 < 0 ? pl->input_ch : (c.kind == KIND_S2 ? c.cin0 / 4 : c.cin0 + c.cin1);
  *cout = c.kind == KIND_CT ? c.cout / (c.S * c.S) : c.cout;
  *dil = c.dil; *name = c.name.c_str();
  return PP_OK;
}
int unet_conv_kind(const UNetPlan* pl, int layer, int* kind, int* scale) {
  PP_REQUIRE(layer >= 0 && layer < int(pl->convs.size()), "unet_conv_kind: bad layer %d", layer);
  *kind = pl->convs[layer].kind;
  *scale = pl->convs[layer].kind == KIND_S2 ? 2 : pl->convs[layer].S;
  return PP_OK;
}
long long unet_workspace_bytes(const UNetPlan* pl, int N, int H, int W, int G) {
  if (check_shape(*pl, N, H, W, G)) return -1;
  return make_layout(*pl, N, H, W, G).total;
}
int unet_activation(const UNetPlan* pl, const char* name, int N, int H, int W, int G, int* act_id, long long* offset,
                    int* C, int* h, int* w) {
  auto it = pl->named.find(name);
  PP_REQUIRE(it != pl->named.end(), "unet_activation: unknown end point '%s'", name);
  const UNetLayout L = make_layout(*pl, N, H, W, G);
  const Act& a = pl->acts[it->second];
  *act_id = it->second; *offset = L.act_data[it->second]; *C = a.C; *h = H / a.res; *w = W / a.res;
  return PP_OK;
}

}  // namespace pp
