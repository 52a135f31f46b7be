// loss.cu — the loss side of the PacingPseudo step as fused, HBM-bound passes over fp32 NCHW logits:
//   * one-hot -> uint8 index map (torch.argmax(scribble, 1), consistency_reglur_memory.py:31)
//   * ONE kernel for partial CE + entropy + branch consistency (+ aux partial CE), forward and backward
//     (losses.py:9-24, 35-43, 45-62, 64-116; call sites consistency_reglur_memory.py:32,41,56-63,81)
//   * Dice (losses.py:147-162), memory-bank update (aux_path_memory.py:68-116, sample 0 only),
//     bank classification loss (consistency_reglur_memory.py:94).
// Thread = pixel; per-class planes are read with unit stride across the warp (coalesced).
#include "pp_common.cuh"
#include <stdlib.h>

namespace pp {

constexpr int kMaxC = 8;

static inline int grid_for_px(long long work, int block) {
  long long g = ceil_div_ll(work, block);
  long long cap = static_cast<long long>(sm_count()) * 8;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

// ---------------------------------------------------------------------------------------------
// argmax over K channels of a one-hot / soft fp32 NCHW tensor -> uint8 (first maximum wins)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) onehot_argmax_kernel(const float* __restrict__ x, uint8_t* __restrict__ out,
                                                            long long P, int HW, int K) {
  for (long long p = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; p < P;
       p += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long n = p / HW, hw = p % HW;
    const float* b = x + n * K * HW + hw;
    float best = b[0];
    int bi = 0;
    for (int k = 1; k < K; ++k) {
      const float v = b[static_cast<long long>(k) * HW];
      if (v > best) { best = v; bi = k; }
    }
    out[p] = static_cast<uint8_t>(bi);
  }
}
int onehot_argmax(const float* x, uint8_t* out, int N, int K, int HW, cudaStream_t s) {
  PP_REQUIRE(K >= 1 && K <= 255, "onehot_argmax: K=%d unsupported", K);
  const long long P = static_cast<long long>(N) * HW;
  onehot_argmax_kernel<<<grid_for_px(P, 256), 256, 0, s>>>(x, out, P, HW, K);
  PP_LAUNCH_CHECK();
  return PP_OK;
}

// ---------------------------------------------------------------------------------------------
// fused scribble loss
// ---------------------------------------------------------------------------------------------
enum : int { CR_NONE = 0, CR_CE = 1, CR_L1 = 2, CR_L2 = 3, CR_KL = 4 };
// accumulator slots (double)
enum : int { ACC_PCE = 0, ACC_NLAB = 1, ACC_ENT = 2, ACC_MASK = 3, ACC_CR = 4, ACC_AUX = 5, ACC_SLOTS = 8 };

// NC = compile-time class capacity of the per-pixel arrays: the fused scribble-loss kernels are instantiated for
// NC in {2, 4, 5, 8} so that C = 5 (CHAOS) does not carry 8-wide register arrays (162 -> ~100 registers, 2x occupancy)
template <int NC>
struct SoftmaxT {
  float p[NC], lp[NC];
};
using Softmax = SoftmaxT<kMaxC>;
template <int NC>
__device__ __forceinline__ void softmax_of(const float (&v)[NC], int C, SoftmaxT<NC>& s) {
  float mx = -INFINITY;
#pragma unroll
  for (int c = 0; c < NC; ++c)
    if (c < C) mx = fmaxf(mx, v[c]);
  float sum = 0.f;
#pragma unroll
  for (int c = 0; c < NC; ++c)
    if (c < C) { s.p[c] = __expf(v[c] - mx); sum += s.p[c]; }
  const float inv = 1.f / sum, lse = mx + __logf(sum);
#pragma unroll
  for (int c = 0; c < NC; ++c)
    if (c < C) { s.p[c] *= inv; s.lp[c] = v[c] - lse; }
    else { s.p[c] = 0.f; s.lp[c] = 0.f; }
}
__device__ __forceinline__ void load_softmax(const float* __restrict__ z, long long HW, int C, Softmax& s) {
  float v[kMaxC];
#pragma unroll
  for (int c = 0; c < kMaxC; ++c) v[c] = (c < C) ? z[c * HW] : 0.f;
  softmax_of(v, C, s);
}
// V consecutive pixels of every class plane: one 16-byte load per plane when V == 4 (HW % 4 == 0 keeps the group
// inside one image and 16-byte aligned), scalar loads when V == 1.
template <int V, int NC>
__device__ __forceinline__ void load_planes(const float* __restrict__ z, int HW, int C, float (&v)[V][NC]) {
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    if (c < C) {
      if constexpr (V == 4) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(z + static_cast<size_t>(c) * HW));
        v[0][c] = q.x; v[1][c] = q.y; v[2][c] = q.z; v[3][c] = q.w;
      } else {
        v[0][c] = __ldg(z + static_cast<size_t>(c) * HW);
      }
    } else {
#pragma unroll
      for (int j = 0; j < V; ++j) v[j][c] = 0.f;
    }
  }
}
template <int V, int NC>
__device__ __forceinline__ void store_planes(float* __restrict__ z, int HW, int C, const float (&v)[V][NC]) {
#pragma unroll
  for (int c = 0; c < NC; ++c)
    if (c < C) {
      if constexpr (V == 4)
        *reinterpret_cast<float4*>(z + static_cast<size_t>(c) * HW) = make_float4(v[0][c], v[1][c], v[2][c], v[3][c]);
      else
        z[static_cast<size_t>(c) * HW] = v[0][c];
    }
}
template <int V>
__device__ __forceinline__ void load_target_mask(const uint8_t* __restrict__ target, const float* __restrict__ mask,
                                                 int p, int ignore_index, int (&t)[V], float (&m)[V]) {
  if constexpr (V == 4) {
    if (target) {
      const uint32_t q = __ldg(reinterpret_cast<const uint32_t*>(target + p));
      t[0] = q & 0xff; t[1] = (q >> 8) & 0xff; t[2] = (q >> 16) & 0xff; t[3] = q >> 24;
    } else { t[0] = t[1] = t[2] = t[3] = ignore_index; }
    if (mask) {
      const float4 q = __ldg(reinterpret_cast<const float4*>(mask + p));
      m[0] = q.x; m[1] = q.y; m[2] = q.z; m[3] = q.w;
    } else { m[0] = m[1] = m[2] = m[3] = 1.f; }
  } else {
    t[0] = target ? target[p] : ignore_index;
    m[0] = mask ? mask[p] : 1.f;
  }
}
__device__ __forceinline__ float sgn(float v) { return (v > 0.f) - (v < 0.f); }

template <int NC>
__device__ __forceinline__ float cr_pixel(int variant, int C, const SoftmaxT<NC>& w, const SoftmaxT<NC>& s) {
  float L = 0.f;
#pragma unroll
  for (int c = 0; c < NC; ++c)
    if (c < C) {
      if (variant == CR_CE) L -= w.p[c] * s.lp[c];
      else if (variant == CR_L1) L += fabsf(s.p[c] - w.p[c]);
      else if (variant == CR_L2) { const float d = s.p[c] - w.p[c]; L = fmaf(d, d, L); }
      else if (variant == CR_KL) L += w.p[c] * (w.lp[c] - s.lp[c]);
    }
  return L;
}

__device__ __forceinline__ void block_accumulate(float (&v)[6], double* __restrict__ acc) {
  __shared__ float red[6][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const float r = warp_sum(v[k]);
    if (lane == 0) red[k][warp] = r;
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    double t = 0.0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) t += static_cast<double>(red[threadIdx.x][w]);
    atomicAdd(acc + threadIdx.x, t);
  }
}

// Aux logits. The reference up-samples fc_cls's [N][C][h][w] output to the label size (F.interpolate, bilinear,
// align_corners=True: aux_path_memory.py:52) and takes the partial CE of that tensor. Only LABELLED pixels (~1 % of a
// scribble map) ever read it, so with AuxLow.h > 0 the kernels take the small tensor itself and interpolate the C logits
// of a labelled pixel in place (4 taps per class out of L1/L2, same expression as upsample_planes_fwd), and the backward
// pass scatters the pixel's gradient to the same 4 taps with fp32 atomics into the pre-zeroed [N][C][h][w] gradient.
// No full-resolution aux tensor or gradient is written or read. AuxLow.h == 0: za / dza are full-resolution planes.
struct AuxLow { int h, w, W; float sh, sw; };
template <int NC>
__device__ __forceinline__ void aux_logits_at(const float* __restrict__ za, const AuxLow& al, int n, int hw, int HW, int C,
                                              float (&v)[NC]) {
  if (al.h == 0) {
    const float* z = za + static_cast<size_t>(n) * C * HW + hw;
#pragma unroll
    for (int c = 0; c < NC; ++c) v[c] = (c < C) ? __ldg(z + static_cast<size_t>(c) * HW) : 0.f;
  } else {
    const int Y = hw / al.W, X = hw - Y * al.W;
    const Lerp ly = lerp_src(Y, al.h, al.sh), lx = lerp_src(X, al.w, al.sw);
    const float* z = za + static_cast<size_t>(n) * C * al.h * al.w;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      if (c < C) {
        const float* r0 = z + (static_cast<size_t>(c) * al.h + ly.i0) * al.w;
        const float* r1 = z + (static_cast<size_t>(c) * al.h + ly.i1) * al.w;
        v[c] = ly.w0 * (lx.w0 * __ldg(r0 + lx.i0) + lx.w1 * __ldg(r0 + lx.i1)) +
               ly.w1 * (lx.w0 * __ldg(r1 + lx.i0) + lx.w1 * __ldg(r1 + lx.i1));
      } else {
        v[c] = 0.f;
      }
    }
  }
}
// The scatter accumulates in 2^-44 fixed point with INTEGER atomics: integer addition is associative, so the low-resolution
// gradient is bit-reproducible whatever order the blocks arrive in. (fp32 atomics here would be the only order-dependent
// sum on the DATA-gradient path: their 1e-7 noise flips bf16 roundings further down the backward chain and two passes over
// the same batch then differ by 1e-4 in the early layers' gradients - tests/run_dp_check.py relies on repeatability.)
// Resolution 5.7e-14, range +-5.2e5 per element.
constexpr double kAuxFixScale = 17592186044416.0;   // 2^44
__device__ __forceinline__ void aux_fix_add(unsigned long long* p, float v) {
  atomicAdd(p, static_cast<unsigned long long>(__double2ll_rn(static_cast<double>(v) * kAuxFixScale)));
}
template <int NC>
__device__ __forceinline__ void aux_grad_scatter(unsigned long long* __restrict__ dzq, const AuxLow& al, int n, int hw,
                                                 int C, const float (&g)[NC]) {
  const int Y = hw / al.W, X = hw - Y * al.W;
  const Lerp ly = lerp_src(Y, al.h, al.sh), lx = lerp_src(X, al.w, al.sw);
  unsigned long long* z = dzq + static_cast<size_t>(n) * C * al.h * al.w;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    if (c < C) {
      unsigned long long* r0 = z + (static_cast<size_t>(c) * al.h + ly.i0) * al.w;
      unsigned long long* r1 = z + (static_cast<size_t>(c) * al.h + ly.i1) * al.w;
      aux_fix_add(r0 + lx.i0, ly.w0 * lx.w0 * g[c]);
      aux_fix_add(r0 + lx.i1, ly.w0 * lx.w1 * g[c]);
      aux_fix_add(r1 + lx.i0, ly.w1 * lx.w0 * g[c]);
      aux_fix_add(r1 + lx.i1, ly.w1 * lx.w1 * g[c]);
    }
  }
}
__global__ void __launch_bounds__(256) aux_fix_to_float_kernel(const long long* __restrict__ q, float* __restrict__ out,
                                                               int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = static_cast<float>(static_cast<double>(q[i]) * (1.0 / kAuxFixScale));
}

template <int V, int NC>
__global__ void __launch_bounds__(256)
scribble_loss_fwd_kernel(const float* __restrict__ zw, const float* __restrict__ zs, const float* __restrict__ za,
                         const uint8_t* __restrict__ target, const float* __restrict__ mask, double* __restrict__ acc,
                         int P, int HW, int C, int ignore_index, int do_ent, int cr_variant, const AuxLow al) {
  float part[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const int groups = P / V;
  for (int gi = blockIdx.x * blockDim.x + threadIdx.x; gi < groups; gi += gridDim.x * blockDim.x) {
    const int p = gi * V;
    const int n = p / HW, hw = p - n * HW;
    const size_t off = static_cast<size_t>(n) * C * HW + hw;
    float vw[V][NC], vs[V][NC];
    int tv[V];
    float mv[V];
    load_planes<V, NC>(zw + off, HW, C, vw);
    if (cr_variant != CR_NONE) load_planes<V, NC>(zs + off, HW, C, vs);
    load_target_mask<V>(target, mask, p, ignore_index, tv, mv);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      SoftmaxT<NC> w;
      softmax_of(vw[j], C, w);
      const int t = tv[j];
      const bool lab = (target != nullptr) && (t != ignore_index) && (t < C);
      const float m = mv[j];
      if (lab) {
        float lpt = 0.f;
#pragma unroll
        for (int c = 0; c < NC; ++c) lpt = (c == t) ? w.lp[c] : lpt;
        part[ACC_PCE] -= lpt;
        part[ACC_NLAB] += 1.f;
      }
      part[ACC_MASK] += m;
      if (do_ent) {
        float H = 0.f;
#pragma unroll
        for (int c = 0; c < NC; ++c)
          if (c < C) H -= w.p[c] * w.lp[c];
        part[ACC_ENT] += m * H;
      }
      if (cr_variant != CR_NONE) {
        SoftmaxT<NC> sx;
        softmax_of(vs[j], C, sx);
        part[ACC_CR] += m * cr_pixel(cr_variant, C, w, sx);
      }
      if (za != nullptr && lab) {   // aux logits matter on labelled pixels only
        float va[NC];
        aux_logits_at<NC>(za, al, n, hw + j, HW, C, va);
        SoftmaxT<NC> a;
        softmax_of(va, C, a);
        float lpt = 0.f;
#pragma unroll
        for (int c = 0; c < NC; ++c) lpt = (c == t) ? a.lp[c] : lpt;
        part[ACC_AUX] -= lpt;
      }
    }
  }
  block_accumulate(part, acc);
}

// Normalisers (losses.py:21-23,58-61,75-78): with a mask  sum / max(sum(mask), 1e-8); without, the
// element mean: N*C*H*W elements for entropy / ce / kl, N*H*W for l1 / l2 (keepdim channel sum).
__device__ __forceinline__ double masked_denom(const double* acc, int has_mask, double nelem) {
  return has_mask ? fmax(acc[ACC_MASK], 1e-8) : nelem;
}

__global__ void scribble_loss_finalize_kernel(const double* __restrict__ acc, float* loss_pce, float* loss_ent,
                                              float* loss_cr, float* loss_aux, long long P, int C, int has_mask,
                                              int cr_variant) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double nlab = acc[ACC_NLAB];
  if (loss_pce) *loss_pce = static_cast<float>(acc[ACC_PCE] / nlab);  // 0/0 -> NaN like torch (all ignored)
  if (loss_aux) *loss_aux = static_cast<float>(acc[ACC_AUX] / nlab);
  if (loss_ent) *loss_ent = static_cast<float>(acc[ACC_ENT] / masked_denom(acc, has_mask, double(P) * C));
  if (loss_cr) {
    const double ne = (cr_variant == CR_L1 || cr_variant == CR_L2) ? double(P) : double(P) * C;
    *loss_cr = static_cast<float>(acc[ACC_CR] / masked_denom(acc, has_mask, ne));
  }
}

// ---------------------------------------------------------------------------------------------
// Lean instantiations of the fused scribble loss: class count C and consistency variant CR are compile-time, four
// pixels per thread, softmax in the log2 domain on MUFU.EX2 / LG2 / RCP. Per pixel and logits tensor:
//   d_c = z_c - max,  e_c = 2^(d_c*log2e),  S = sum e,  lS = ln S,  p_c = e_c / S,  log p_c = d_c - lS
// so that  H = lS - (sum e_c d_c)/S  and  CE(p_w, s) = lS_s - (sum e_w,c d_s,c)/S_w  need neither p nor log p in the
// forward pass (both are sums of non-negative terms: no cancellation). ~70 instructions per pixel forward instead of
// ~430 in the generic kernel (runtime C / variant predicates), which made that kernel issue- rather than HBM-bound.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float fast_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
constexpr float kLog2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f;

template <int C>
struct LeanSoftmax {
  float d[C], e[C], inv, lS;   // shifted logits, unnormalised exponentials, 1/S, ln S
  __device__ __forceinline__ void of(const float (&v)[C]) {
    float mx = v[0];
#pragma unroll
    for (int c = 1; c < C; ++c) mx = fmaxf(mx, v[c]);
    float S = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) { d[c] = v[c] - mx; e[c] = fast_ex2(d[c] * kLog2e); S += e[c]; }
    inv = fast_rcp(S);
    lS = fast_lg2(S) * kLn2;
  }
  __device__ __forceinline__ float p(int c) const { return e[c] * inv; }
  __device__ __forceinline__ float lp(int c) const { return d[c] - lS; }
};

template <int C, int V>
__device__ __forceinline__ void load_planes4(const float* __restrict__ z, int HW, float (&v)[V][C]) {
#pragma unroll
  for (int c = 0; c < C; ++c)
#pragma unroll
    for (int h = 0; h < V; h += 4) {
      const float4 q = __ldg(reinterpret_cast<const float4*>(z + static_cast<size_t>(c) * HW + h));
      v[h][c] = q.x; v[h + 1][c] = q.y; v[h + 2][c] = q.z; v[h + 3][c] = q.w;
    }
}
template <int C, int V>
__device__ __forceinline__ void store_planes4(float* __restrict__ z, int HW, const float (&v)[V][C]) {
#pragma unroll
  for (int c = 0; c < C; ++c)
#pragma unroll
    for (int h = 0; h < V; h += 4)
      *reinterpret_cast<float4*>(z + static_cast<size_t>(c) * HW + h) =
          make_float4(v[h][c], v[h + 1][c], v[h + 2][c], v[h + 3][c]);
}
template <int V>
__device__ __forceinline__ void load_target_mask_v(const uint8_t* __restrict__ target, const float* __restrict__ mask,
                                                   int p, int ignore_index, int (&t)[V], float (&m)[V]) {
#pragma unroll
  for (int h = 0; h < V; h += 4) {
    int t4[4];
    float m4[4];
    load_target_mask<4>(target, mask, p + h, ignore_index, t4, m4);
#pragma unroll
    for (int j = 0; j < 4; ++j) { t[h + j] = t4[j]; m[h + j] = m4[j]; }
  }
}
template <int C>
__device__ __forceinline__ float pick(const float (&a)[C], int t) {
  float r = a[0];
#pragma unroll
  for (int c = 1; c < C; ++c) r = (c == t) ? a[c] : r;
  return r;
}

// per-pixel consistency term from two softmaxes (w = weak = target side, s = strong)
template <int C, int CR>
__device__ __forceinline__ float lean_cr_pixel(const LeanSoftmax<C>& w, const LeanSoftmax<C>& s) {
  float acc = 0.f;
  if constexpr (CR == CR_CE) {
#pragma unroll
    for (int c = 0; c < C; ++c) acc = fmaf(w.e[c], s.d[c], acc);
    return s.lS - acc * w.inv;
  } else if constexpr (CR == CR_KL) {
#pragma unroll
    for (int c = 0; c < C; ++c) acc = fmaf(w.e[c], w.d[c] - s.d[c], acc);
    return (s.lS - w.lS) + acc * w.inv;
  } else if constexpr (CR == CR_L1) {
#pragma unroll
    for (int c = 0; c < C; ++c) acc += fabsf(s.p(c) - w.p(c));
    return acc;
  } else {
#pragma unroll
    for (int c = 0; c < C; ++c) { const float df = s.p(c) - w.p(c); acc = fmaf(df, df, acc); }
    return acc;
  }
}

// Balanced contiguous spans instead of a grid-stride loop: block b owns groups [G*b/nb, G*(b+1)/nb) rounded to
// 8 groups (128 bytes per class plane), so with ~1.7 groups per thread (12 pairs of 256^2 on 148 x 3 blocks) every
// block finishes at the same time instead of 73 % of the blocks running a second full iteration.
template <int V>
__device__ __forceinline__ void lean_span(int groups, int& begin, int& end) {
  constexpr int U = 32 / V;                     // groups per 128-byte line of one class plane
  const long long units = (groups + U - 1) / U;
  const int u0 = static_cast<int>(units * blockIdx.x / gridDim.x), u1 = static_cast<int>(units * (blockIdx.x + 1) / gridDim.x);
  begin = u0 * U + threadIdx.x;
  end = min(u1 * U, groups);
}

template <int C, int CR, int V>
__global__ void __launch_bounds__(256, 3)
scribble_loss_fwd_lean_kernel(const float* __restrict__ zw, const float* __restrict__ zs, const float* __restrict__ za,
                              const uint8_t* __restrict__ target, const float* __restrict__ mask,
                              double* __restrict__ acc, int P, int HW, int ignore_index, int do_ent, const AuxLow al) {
  float part[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  int gi, g_end;
  lean_span<V>(P / V, gi, g_end);
  for (; gi < g_end; gi += blockDim.x) {
    const int p = gi * V;
    const int n = p / HW, hw = p - n * HW;
    const size_t off = static_cast<size_t>(n) * C * HW + hw;
    float vw[V][C], vs[V][C];
    int tv[V];
    float mv[V];
    load_planes4<C, V>(zw + off, HW, vw);
    if constexpr (CR != CR_NONE) load_planes4<C, V>(zs + off, HW, vs);
    load_target_mask_v<V>(target, mask, p, ignore_index, tv, mv);
    bool any_lab = false;
#pragma unroll
    for (int j = 0; j < V; ++j) any_lab |= (tv[j] != ignore_index) && (tv[j] < C);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      LeanSoftmax<C> w;
      w.of(vw[j]);
      const float m = mv[j];
      part[ACC_MASK] += m;
      if (do_ent) {
        float ed = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) ed = fmaf(w.e[c], w.d[c], ed);
        part[ACC_ENT] = fmaf(m, w.lS - ed * w.inv, part[ACC_ENT]);
      }
      if constexpr (CR != CR_NONE) {
        LeanSoftmax<C> s;
        s.of(vs[j]);
        part[ACC_CR] = fmaf(m, lean_cr_pixel<C, CR>(w, s), part[ACC_CR]);
      }
      if (any_lab) {                                   // scribble pixels are ~1 % of the image
        const int t = tv[j];
        if ((t != ignore_index) && (t < C)) {
          part[ACC_PCE] += w.lS - pick<C>(w.d, t);
          part[ACC_NLAB] += 1.f;
        }
      }
    }
    if (za != nullptr && any_lab) {                    // aux logits matter on labelled pixels only
#pragma unroll
      for (int j = 0; j < V; ++j) {
        const int t = tv[j];
        if ((t != ignore_index) && (t < C)) {
          float va[C];
          aux_logits_at<C>(za, al, n, hw + j, HW, C, va);
          LeanSoftmax<C> a;
          a.of(va);
          part[ACC_AUX] += a.lS - pick<C>(a.d, t);
        }
      }
    }
  }
  block_accumulate(part, acc);
}

template <int C, int CR, int V>
__global__ void __launch_bounds__(256, 3)
scribble_loss_bwd_lean_kernel(const float* __restrict__ zw, const float* __restrict__ zs, const float* __restrict__ za,
                              const uint8_t* __restrict__ target, const float* __restrict__ mask,
                              const double* __restrict__ acc, const float* __restrict__ g_pce,
                              const float* __restrict__ g_ent, const float* __restrict__ g_cr,
                              const float* __restrict__ g_aux, float* __restrict__ dzw, float* __restrict__ dzs,
                              float* __restrict__ dza, int P, int HW, int ignore_index, int do_ent, int detach_weak,
                              const AuxLow al, unsigned long long* __restrict__ dzq) {
  const float gp = g_pce ? *g_pce : 0.f, ge = (do_ent && g_ent) ? *g_ent : 0.f;
  const float gc = (CR != CR_NONE && g_cr) ? *g_cr : 0.f, ga = (za && g_aux) ? *g_aux : 0.f;
  const int has_mask = mask != nullptr;
  const double nlab = acc[ACC_NLAB];
  const float inv_lab = nlab > 0.0 ? static_cast<float>(1.0 / nlab) : 0.f;
  const float k_ent = ge * static_cast<float>(1.0 / masked_denom(acc, has_mask, double(P) * C));
  const double ne = (CR == CR_L1 || CR == CR_L2) ? double(P) : double(P) * C;
  const float k_cr = gc * static_cast<float>(1.0 / masked_denom(acc, has_mask, ne));
  const float k_pce = gp * inv_lab, k_aux = ga * inv_lab;
  const bool weak_gets_cr = (CR == CR_KL) || !detach_weak;
  const bool do_aux = za != nullptr && (al.h > 0 ? dzq != nullptr : dza != nullptr);

  int gi, g_end;
  lean_span<V>(P / V, gi, g_end);
  for (; gi < g_end; gi += blockDim.x) {
    const int p = gi * V;
    const int n = p / HW, hw = p - n * HW;
    const size_t off = static_cast<size_t>(n) * C * HW + hw;
    float vw[V][C], vs[V][C];
    int tv[V];
    float mv[V];
    load_planes4<C, V>(zw + off, HW, vw);
    if constexpr (CR != CR_NONE) load_planes4<C, V>(zs + off, HW, vs);
    load_target_mask_v<V>(target, mask, p, ignore_index, tv, mv);
    bool any_lab = false;
#pragma unroll
    for (int j = 0; j < V; ++j) any_lab |= (tv[j] != ignore_index) && (tv[j] < C);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      LeanSoftmax<C> w;
      w.of(vw[j]);
      const float m = mv[j];
      float pw[C], coef[C];                        // dz_w = p_w * coef (+ the pCE term on labelled pixels)
#pragma unroll
      for (int c = 0; c < C; ++c) { pw[c] = w.p(c); coef[c] = 0.f; }
      if (do_ent) {                                // -k p (log p + H)
        float ed = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) ed = fmaf(pw[c], w.d[c], ed);
        const float k = k_ent * m;                 // log p + H = d - lS + lS - sum p d = d - sum p d
#pragma unroll
        for (int c = 0; c < C; ++c) coef[c] = k * (ed - w.d[c]);
      }
      if constexpr (CR != CR_NONE) {
        LeanSoftmax<C> s;
        s.of(vs[j]);
        const float k = k_cr * m;
        if constexpr (CR == CR_CE || CR == CR_KL) {
          // strong: k (p_s - p_w). weak, ce: -k p_w (log p_s + L) with log p_s + L = d_s - sum p_w d_s;
          // kl: k p_w ((log p_w - log p_s) - L) with that bracket = (d_w - d_s) - sum p_w (d_w - d_s)
          float q[C], qs = 0.f;
#pragma unroll
          for (int c = 0; c < C; ++c) {
            q[c] = (CR == CR_CE) ? s.d[c] : (s.d[c] - w.d[c]);
            qs = fmaf(pw[c], q[c], qs);
          }
          if (weak_gets_cr) {
#pragma unroll
            for (int c = 0; c < C; ++c) coef[c] = fmaf(k, qs - q[c], coef[c]);
          }
#pragma unroll
          for (int c = 0; c < C; ++c) vs[j][c] = k * (s.p(c) - pw[c]);
        } else {
          float e[C], es = 0.f, ew = 0.f;
#pragma unroll
          for (int c = 0; c < C; ++c) {
            const float df = s.p(c) - pw[c];
            e[c] = (CR == CR_L1) ? sgn(df) : 2.f * df;
            es = fmaf(e[c], s.p(c), es);
            ew = fmaf(e[c], pw[c], ew);
          }
#pragma unroll
          for (int c = 0; c < C; ++c) {
            vs[j][c] = k * s.p(c) * (e[c] - es);
            if (weak_gets_cr) coef[c] = fmaf(k, ew - e[c], coef[c]);
          }
        }
      }
#pragma unroll
      for (int c = 0; c < C; ++c) vw[j][c] = pw[c] * coef[c];
      if (any_lab) {
        const int t = tv[j];
        if ((t != ignore_index) && (t < C)) {
#pragma unroll
          for (int c = 0; c < C; ++c) vw[j][c] = fmaf(k_pce, pw[c] - (c == t ? 1.f : 0.f), vw[j][c]);
        }
      }
    }
    if constexpr (CR != CR_NONE) {
      if (dzs != nullptr) store_planes4<C, V>(dzs + off, HW, vs);
    }
    if (dzw != nullptr) store_planes4<C, V>(dzw + off, HW, vw);
    if (do_aux && al.h > 0) {                          // low-resolution aux logits: labelled pixels scatter to 4 taps
      if (any_lab) {
#pragma unroll
        for (int j = 0; j < V; ++j) {
          const int t = tv[j];
          if ((t != ignore_index) && (t < C)) {
            float va[C];
            aux_logits_at<C>(za, al, n, hw + j, HW, C, va);
            LeanSoftmax<C> a;
            a.of(va);
#pragma unroll
            for (int c = 0; c < C; ++c) va[c] = k_aux * (a.p(c) - (c == t ? 1.f : 0.f));
            aux_grad_scatter<C>(dzq, al, n, hw + j, C, va);
          }
        }
      }
    } else if (do_aux) {
      float va[V][C];
      if (any_lab) {
        load_planes4<C, V>(za + off, HW, va);
#pragma unroll
        for (int j = 0; j < V; ++j) {
          const int t = tv[j];
          const bool lab = (t != ignore_index) && (t < C);
          LeanSoftmax<C> a;
          a.of(va[j]);
#pragma unroll
          for (int c = 0; c < C; ++c) va[j][c] = lab ? k_aux * (a.p(c) - (c == t ? 1.f : 0.f)) : 0.f;
        }
      } else {
#pragma unroll
        for (int j = 0; j < V; ++j)
#pragma unroll
          for (int c = 0; c < C; ++c) va[j][c] = 0.f;
      }
      store_planes4<C, V>(dza + off, HW, va);
    }
  }
}

// persistent-style grid for the lean kernels: three 256-thread blocks per SM (launch bounds), grid-stride loop
static inline int lean_grid(long long groups) {
  long long g = ceil_div_ll(groups, 256);
  const long long cap = static_cast<long long>(sm_count()) * 3;
  if (g > cap) g = cap;
  return static_cast<int>(g < 1 ? 1 : g);
}
static inline bool lean_classes(int C) { return C >= 2 && C <= 5; }
// (eight pixels per thread were measured too: no gain at C = 2, spills at C = 5; V stays a template parameter)
// PP_LOSS_GENERIC=1 routes every call to the generic (runtime C / variant) kernels: A/B timing and cross-checks
static bool generic_loss_forced() {   // read per call so a test can toggle it inside one process
  const char* e = getenv("PP_LOSS_GENERIC");
  return e != nullptr && e[0] == '1';
}

// 4 pixels per thread need HW % 4 == 0 (a group never straddles two images) and 16 / 4-byte aligned base pointers
static bool loss_vec4_ok(int HW, const void* a, const void* b, const void* c, const void* t, const void* m,
                         const void* d = nullptr, const void* e = nullptr, const void* f = nullptr) {
  if (HW % 4 != 0) return false;
  const void* f4[] = {a, b, c, m, d, e, f};
  for (const void* q : f4)
    if (q != nullptr && (reinterpret_cast<uintptr_t>(q) & 15) != 0) return false;
  return t == nullptr || (reinterpret_cast<uintptr_t>(t) & 3) == 0;
}

int scribble_loss_fwd(const float* zw, const float* zs, const float* za, const uint8_t* target, const float* mask,
                      double* acc, float* loss_pce, float* loss_ent, float* loss_cr, float* loss_aux, int N, int C,
                      int HW, int ignore_index, int do_ent, int cr_variant, cudaStream_t s, int aux_h, int aux_w,
                      int W) {
  PP_REQUIRE(C >= 1 && C <= kMaxC, "scribble_loss: num_classes=%d unsupported (max %d)", C, kMaxC);
  AuxLow al{0, 0, 0, 0.f, 0.f};
  if (za != nullptr && aux_h > 0) {   // za is the low-resolution tensor [N][C][aux_h][aux_w]
    PP_REQUIRE(aux_w > 0 && W > 0 && HW % W == 0, "scribble_loss: bad aux / image size (%d x %d, W=%d, HW=%d)", aux_h, aux_w, W, HW);
    al = AuxLow{aux_h, aux_w, W, ac_scale(aux_h, HW / W), ac_scale(aux_w, W)};
  }
  const bool aux_full = za != nullptr && al.h == 0;
  PP_REQUIRE(cr_variant >= CR_NONE && cr_variant <= CR_KL, "scribble_loss: bad consistency variant %d", cr_variant);
  PP_REQUIRE((cr_variant == CR_NONE) == (zs == nullptr), "scribble_loss: strong logits / variant mismatch");
  const long long P = static_cast<long long>(N) * HW;
  PP_REQUIRE(P * C < (1LL << 31), "scribble_loss: %lld logits exceed the 32-bit index range", P * C);
  PP_CHECK_CUDA(cudaMemsetAsync(acc, 0, sizeof(double) * ACC_SLOTS, s));
  // algorithmic bytes: every logits tensor read once, u8 target + fp32 mask read once
  // (a low-resolution aux tensor adds N*C*h*w*4 bytes: 0.2 MB for 12 x 5 x 32 x 32)
  const double bytes = static_cast<double>(P) * ((1 + (zs != nullptr) + aux_full) * 4.0 * C + (target != nullptr) +
                                                 4.0 * (mask != nullptr)) +
                       (al.h > 0 ? 4.0 * N * C * al.h * al.w : 0.0);
  const int slot = prof_begin(PROF_LOSS, bytes, s);
#define PP_LOSS_FWD(V_, NC_, P_)                                                                              \
  scribble_loss_fwd_kernel<V_, NC_><<<grid_for_px(P_, 256), 256, 0, s>>>(zw, zs, za, target, mask, acc, int(P), HW, C, \
                                                                         ignore_index, do_ent, cr_variant, al)
#define PP_LEAN_FWD(C_, CR_)                                                                                       \
  scribble_loss_fwd_lean_kernel<C_, CR_, 4><<<lean_grid(P / 4), 256, 0, s>>>(zw, zs, za, target, mask, acc, int(P), HW, \
                                                                            ignore_index, do_ent, al)
#define PP_LEAN_FWD_C(C_)                                                                                           \
  switch (cr_variant) {                                                                                             \
    case CR_NONE: PP_LEAN_FWD(C_, CR_NONE); break;                                                                  \
    case CR_CE: PP_LEAN_FWD(C_, CR_CE); break;                                                                      \
    case CR_L1: PP_LEAN_FWD(C_, CR_L1); break;                                                                      \
    case CR_L2: PP_LEAN_FWD(C_, CR_L2); break;                                                                      \
    default: PP_LEAN_FWD(C_, CR_KL); break;                                                                         \
  }
  const bool vec4 = loss_vec4_ok(HW, zw, zs, nullptr, target, mask);   // (the aux logits are read with scalar loads)
  if (vec4 && target != nullptr && lean_classes(C) && !generic_loss_forced()) {
    if (C == 2) { PP_LEAN_FWD_C(2); } else if (C == 3) { PP_LEAN_FWD_C(3); }
    else if (C == 4) { PP_LEAN_FWD_C(4); } else { PP_LEAN_FWD_C(5); }
  } else if (vec4) {
    if (C <= 2) PP_LOSS_FWD(4, 2, P / 4); else if (C <= 4) PP_LOSS_FWD(4, 4, P / 4);
    else if (C == 5) PP_LOSS_FWD(4, 5, P / 4); else PP_LOSS_FWD(4, 8, P / 4);
  } else {
    PP_LOSS_FWD(1, 8, P);
  }
#undef PP_LOSS_FWD
#undef PP_LEAN_FWD_C
#undef PP_LEAN_FWD
  prof_end(slot, s);
  scribble_loss_finalize_kernel<<<1, 32, 0, s>>>(acc, loss_pce, do_ent ? loss_ent : nullptr,
                                                 cr_variant != CR_NONE ? loss_cr : nullptr,
                                                 za != nullptr ? loss_aux : nullptr, P, C, mask != nullptr, cr_variant);
  PP_LAUNCH_CHECK_N(2);
  return PP_OK;
}

template <int V, int NC>
__global__ void __launch_bounds__(256)
scribble_loss_bwd_kernel(const float* __restrict__ zw, const float* __restrict__ zs, const float* __restrict__ za,
                         const uint8_t* __restrict__ target, const float* __restrict__ mask,
                         const double* __restrict__ acc, const float* __restrict__ g_pce,
                         const float* __restrict__ g_ent, const float* __restrict__ g_cr,
                         const float* __restrict__ g_aux, float* __restrict__ dzw, float* __restrict__ dzs,
                         float* __restrict__ dza, int P, int HW, int C, int ignore_index, int do_ent,
                         int cr_variant, int detach_weak, const AuxLow al, unsigned long long* __restrict__ dzq) {
  const float gp = g_pce ? *g_pce : 0.f, ge = (do_ent && g_ent) ? *g_ent : 0.f;
  const float gc = (cr_variant != CR_NONE && g_cr) ? *g_cr : 0.f, ga = (za && g_aux) ? *g_aux : 0.f;
  const int has_mask = mask != nullptr;
  const double nlab = acc[ACC_NLAB];
  const float inv_lab = nlab > 0.0 ? static_cast<float>(1.0 / nlab) : 0.f;
  const float inv_ent = static_cast<float>(1.0 / masked_denom(acc, has_mask, double(P) * C));
  const double ne = (cr_variant == CR_L1 || cr_variant == CR_L2) ? double(P) : double(P) * C;
  const float inv_cr = static_cast<float>(1.0 / masked_denom(acc, has_mask, ne));
  const bool weak_gets_cr = (cr_variant == CR_KL) || !detach_weak;

  const int groups = P / V;
  for (int gi = blockIdx.x * blockDim.x + threadIdx.x; gi < groups; gi += gridDim.x * blockDim.x) {
    const int p = gi * V;
    const int n = p / HW, hw = p - n * HW;
    const size_t off = static_cast<size_t>(n) * C * HW + hw;
    float vw[V][NC], vs[V][NC], va[V][NC];
    int tv[V];
    float mv[V];
    load_planes<V, NC>(zw + off, HW, C, vw);
    if (cr_variant != CR_NONE) load_planes<V, NC>(zs + off, HW, C, vs);
    load_target_mask<V>(target, mask, p, ignore_index, tv, mv);
    bool any_lab = false;
#pragma unroll
    for (int j = 0; j < V; ++j) any_lab |= (target != nullptr) && (tv[j] != ignore_index) && (tv[j] < C);
    const bool aux_low = al.h > 0;
    const bool do_aux = za != nullptr && (aux_low ? dzq != nullptr : dza != nullptr);
    if (do_aux && any_lab && !aux_low) load_planes<V, NC>(za + off, HW, C, va);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      SoftmaxT<NC> w;
      softmax_of(vw[j], C, w);
      const int t = tv[j];
      const bool lab = (target != nullptr) && (t != ignore_index) && (t < C);
      const float m = mv[j];
      float d[NC];
#pragma unroll
      for (int c = 0; c < NC; ++c) d[c] = lab ? gp * inv_lab * (w.p[c] - (c == t ? 1.f : 0.f)) : 0.f;
      if (do_ent) {
        float H = 0.f;
#pragma unroll
        for (int c = 0; c < NC; ++c)
          if (c < C) H -= w.p[c] * w.lp[c];
        const float k = ge * m * inv_ent;
#pragma unroll
        for (int c = 0; c < NC; ++c) d[c] -= k * w.p[c] * (w.lp[c] + H);
      }
      if (cr_variant != CR_NONE) {
        SoftmaxT<NC> sx;
        softmax_of(vs[j], C, sx);
        const float k = gc * m * inv_cr;
        float ds[NC];
        if (cr_variant == CR_CE || cr_variant == CR_KL) {
          const float L = cr_pixel(cr_variant, C, w, sx);
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            ds[c] = k * (sx.p[c] - w.p[c]);
            if (weak_gets_cr) {
              if (cr_variant == CR_CE) d[c] -= k * w.p[c] * (sx.lp[c] + L);
              else d[c] += k * w.p[c] * ((w.lp[c] - sx.lp[c]) - L);
            }
          }
        } else {
          float e[NC], es = 0.f, ew = 0.f;
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            const float df = sx.p[c] - w.p[c];
            e[c] = (c < C) ? (cr_variant == CR_L1 ? sgn(df) : 2.f * df) : 0.f;
            es = fmaf(e[c], sx.p[c], es);
            ew = fmaf(e[c], w.p[c], ew);
          }
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            ds[c] = k * sx.p[c] * (e[c] - es);
            if (weak_gets_cr) d[c] += k * w.p[c] * (ew - e[c]);
          }
        }
#pragma unroll
        for (int c = 0; c < NC; ++c) vs[j][c] = ds[c];     // the strong logits are consumed: reuse as output
      }
#pragma unroll
      for (int c = 0; c < NC; ++c) vw[j][c] = d[c];
      if (do_aux && aux_low) {
        if (lab) {   // labelled pixels scatter their gradient to the 4 taps of the low-resolution tensor
          float g[NC];
          aux_logits_at<NC>(za, al, n, hw + j, HW, C, g);
          SoftmaxT<NC> a;
          softmax_of(g, C, a);
#pragma unroll
          for (int c = 0; c < NC; ++c) g[c] = ga * inv_lab * (a.p[c] - (c == t ? 1.f : 0.f));
          aux_grad_scatter<NC>(dzq, al, n, hw + j, C, g);
        }
      } else if (do_aux) {
        if (lab) {
          SoftmaxT<NC> a;
          softmax_of(va[j], C, a);
#pragma unroll
          for (int c = 0; c < NC; ++c) va[j][c] = ga * inv_lab * (a.p[c] - (c == t ? 1.f : 0.f));
        } else {
#pragma unroll
          for (int c = 0; c < NC; ++c) va[j][c] = 0.f;
        }
      }
    }
    if (cr_variant != CR_NONE && dzs != nullptr) store_planes<V, NC>(dzs + off, HW, C, vs);
    if (dzw != nullptr) store_planes<V, NC>(dzw + off, HW, C, vw);
    if (do_aux && !aux_low) store_planes<V, NC>(dza + off, HW, C, va);
  }
}

int scribble_loss_bwd(const float* zw, const float* zs, const float* za, const uint8_t* target, const float* mask,
                      const double* acc, const float* g_pce, const float* g_ent, const float* g_cr, const float* g_aux,
                      float* dzw, float* dzs, float* dza, int N, int C, int HW, int ignore_index, int do_ent,
                      int cr_variant, int detach_weak, cudaStream_t s, int aux_h, int aux_w, int W,
                      long long* aux_scratch) {
  PP_REQUIRE(C >= 1 && C <= kMaxC, "scribble_loss_bwd: num_classes=%d unsupported", C);
  AuxLow al{0, 0, 0, 0.f, 0.f};
  unsigned long long* dzq = nullptr;
  const int n_low = N * C * aux_h * aux_w;
  if (za != nullptr && aux_h > 0) {   // za / dza are the low-resolution tensors [N][C][aux_h][aux_w]
    PP_REQUIRE(aux_w > 0 && W > 0 && HW % W == 0, "scribble_loss_bwd: bad aux / image size (%d x %d, W=%d, HW=%d)", aux_h, aux_w, W, HW);
    al = AuxLow{aux_h, aux_w, W, ac_scale(aux_h, HW / W), ac_scale(aux_w, W)};
    if (dza != nullptr) {
      PP_REQUIRE(aux_scratch != nullptr, "scribble_loss_bwd: the low-resolution aux gradient needs its fixed-point scratch");
      dzq = reinterpret_cast<unsigned long long*>(aux_scratch);
      PP_CHECK_CUDA(cudaMemsetAsync(dzq, 0, sizeof(long long) * n_low, s));
    }
  }
  const bool aux_full = za != nullptr && al.h == 0;
  const long long P = static_cast<long long>(N) * HW;
  PP_REQUIRE(P * C < (1LL << 31), "scribble_loss_bwd: %lld logits exceed the 32-bit index range", P * C);
  // algorithmic bytes: forward reads + one write per gradient tensor
  const double bytes = static_cast<double>(P) * ((1 + (zs != nullptr) + aux_full) * 4.0 * C + (target != nullptr) +
                                                 4.0 * (mask != nullptr) +
                                                 ((dzw != nullptr) + (dzs != nullptr) + (aux_full && dza != nullptr)) * 4.0 * C) +
                       (al.h > 0 ? 8.0 * N * C * al.h * al.w : 0.0);
  const int slot = prof_begin(PROF_LOSS, bytes, s);
#define PP_LOSS_BWD(V_, NC_, P_)                                                                                    \
  scribble_loss_bwd_kernel<V_, NC_><<<grid_for_px(P_, 256), 256, 0, s>>>(zw, zs, za, target, mask, acc, g_pce, g_ent, g_cr, \
                                                                         g_aux, dzw, dzs, dza, int(P), HW, C,          \
                                                                         ignore_index, do_ent, cr_variant, detach_weak, al, \
                                                                         dzq)
#define PP_LEAN_BWD(C_, CR_)                                                                                       \
  scribble_loss_bwd_lean_kernel<C_, CR_, 4><<<lean_grid(P / 4), 256, 0, s>>>(                                        \
      zw, zs, za, target, mask, acc, g_pce, g_ent, g_cr, g_aux, dzw, dzs, dza, int(P), HW, ignore_index, do_ent,     \
      detach_weak, al, dzq)
#define PP_LEAN_BWD_C(C_)                                                                                           \
  switch (cr_variant) {                                                                                             \
    case CR_NONE: PP_LEAN_BWD(C_, CR_NONE); break;                                                                  \
    case CR_CE: PP_LEAN_BWD(C_, CR_CE); break;                                                                      \
    case CR_L1: PP_LEAN_BWD(C_, CR_L1); break;                                                                      \
    case CR_L2: PP_LEAN_BWD(C_, CR_L2); break;                                                                      \
    default: PP_LEAN_BWD(C_, CR_KL); break;                                                                         \
  }
  const bool vec4 = loss_vec4_ok(HW, zw, zs, aux_full ? za : nullptr, target, mask, dzw, dzs, aux_full ? dza : nullptr);
  if (vec4 && target != nullptr && lean_classes(C) && !generic_loss_forced()) {
    if (C == 2) { PP_LEAN_BWD_C(2); } else if (C == 3) { PP_LEAN_BWD_C(3); }
    else if (C == 4) { PP_LEAN_BWD_C(4); } else { PP_LEAN_BWD_C(5); }
  } else if (vec4) {
    if (C <= 2) PP_LOSS_BWD(4, 2, P / 4); else if (C <= 4) PP_LOSS_BWD(4, 4, P / 4);
    else if (C == 5) PP_LOSS_BWD(4, 5, P / 4); else PP_LOSS_BWD(4, 8, P / 4);
  } else {
    PP_LOSS_BWD(1, 8, P);
  }
#undef PP_LOSS_BWD
#undef PP_LEAN_BWD_C
#undef PP_LEAN_BWD
  prof_end(slot, s);
  PP_LAUNCH_CHECK();
  if (dzq != nullptr) {
    aux_fix_to_float_kernel<<<ceil_div(n_low, 256), 256, 0, s>>>(reinterpret_cast<const long long*>(dzq), dza, n_low);
    PP_LAUNCH_CHECK();
  }
  return PP_OK;
}

// ---------------------------------------------------------------------------------------------
// Stand-alone pair losses on probability tensors, for the reference's free functions
// soft_label_cross_entropy_loss (a = logits, b = target probabilities; losses.py:45-62),
// l1_loss / l2_loss (a, b = probabilities; losses.py:64-96). pacc: [0] = sum m*L, [1] = sum m.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double pair_denom(const double* pacc, int has_mask, long long P, int C, int variant) {
  if (has_mask) return fmax(pacc[1], 1e-8);
  return variant == CR_CE ? double(P) * C : double(P);
}
__global__ void __launch_bounds__(256) pair_loss_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                            const float* __restrict__ mask, double* __restrict__ pacc,
                                                            long long P, int HW, int C, int variant) {
  float part[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (long long p = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; p < P;
       p += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long n = p / HW, hw = p % HW;
    const long long off = n * C * HW + hw;
    const float m = mask ? mask[p] : 1.f;
    float L = 0.f;
    if (variant == CR_CE) {
      Softmax s;
      load_softmax(a + off, HW, C, s);
#pragma unroll
      for (int c = 0; c < kMaxC; ++c)
        if (c < C) L -= b[off + c * static_cast<long long>(HW)] * s.lp[c];
    } else {
#pragma unroll
      for (int c = 0; c < kMaxC; ++c)
        if (c < C) {
          const float d = a[off + c * static_cast<long long>(HW)] - b[off + c * static_cast<long long>(HW)];
          L += variant == CR_L1 ? fabsf(d) : d * d;
        }
    }
    part[0] += m * L;
    part[1] += m;
  }
  block_accumulate(part, pacc);
}
__global__ void pair_loss_finalize_kernel(const double* __restrict__ pacc, float* loss, long long P, int C,
                                          int has_mask, int variant) {
  if (threadIdx.x == 0 && blockIdx.x == 0) *loss = static_cast<float>(pacc[0] / pair_denom(pacc, has_mask, P, C, variant));
}
__global__ void __launch_bounds__(256) pair_loss_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                            const float* __restrict__ mask,
                                                            const double* __restrict__ pacc, const float* __restrict__ g,
                                                            float* __restrict__ da, float* __restrict__ db, long long P,
                                                            int HW, int C, int variant) {
  const float k0 = (*g) * static_cast<float>(1.0 / pair_denom(pacc, mask != nullptr, P, C, variant));
  for (long long p = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; p < P;
       p += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long n = p / HW, hw = p % HW;
    const long long off = n * C * HW + hw;
    const float k = k0 * (mask ? mask[p] : 1.f);
    if (variant == CR_CE) {
      Softmax s;
      load_softmax(a + off, HW, C, s);
      float bs = 0.f;
#pragma unroll
      for (int c = 0; c < kMaxC; ++c)
        if (c < C) bs += b[off + c * static_cast<long long>(HW)];
#pragma unroll
      for (int c = 0; c < kMaxC; ++c)
        if (c < C) {
          const long long o = off + c * static_cast<long long>(HW);
          if (da) da[o] = k * (s.p[c] * bs - b[o]);
          if (db) db[o] = -k * s.lp[c];
        }
    } else {
#pragma unroll
      for (int c = 0; c < kMaxC; ++c)
        if (c < C) {
          const long long o = off + c * static_cast<long long>(HW);
          const float d = a[o] - b[o];
          const float e = k * (variant == CR_L1 ? sgn(d) : 2.f * d);
          if (da) da[o] = e;
          if (db) db[o] = -e;
        }
    }
  }
}
int pair_loss_fwd(const float* a, const float* b, const float* mask, double* pacc, float* loss, int N, int C, int HW,
                  int variant, cudaStream_t s) {
  PP_REQUIRE(C >= 1 && C <= kMaxC, "pair_loss: num_classes=%d unsupported", C);
  PP_REQUIRE(variant == CR_CE || variant == CR_L1 || variant == CR_L2, "pair_loss: bad variant %d", variant);
  const long long P = static_cast<long long>(N) * HW;
  PP_CHECK_CUDA(cudaMemsetAsync(pacc, 0, sizeof(double) * ACC_SLOTS, s));
  pair_loss_fwd_kernel<<<grid_for_px(P, 256), 256, 0, s>>>(a, b, mask, pacc, P, HW, C, variant);
  pair_loss_finalize_kernel<<<1, 32, 0, s>>>(pacc, loss, P, C, mask != nullptr, variant);
  PP_LAUNCH_CHECK_N(2);
  return PP_OK;
}
int pair_loss_bwd(const float* a, const float* b, const float* mask, const double* pacc, const float* g, float* da,
                  float* db, int N, int C, int HW, int variant, cudaStream_t s) {
  PP_REQUIRE(C >= 1 && C <= kMaxC, "pair_loss: num_classes=%d unsupported", C);
  const long long P = static_cast<long long>(N) * HW;
  pair_loss_bwd_kernel<<<grid_for_px(P, 256), 256, 0, s>>>(a, b, mask, pacc, g, da, db, P, HW, C, variant);
  PP_LAUNCH_CHECK();
  return PP_OK;
}

// ---------------------------------------------------------------------------------------------
// Dice loss (losses.py:147-162): p = softmax(z); per (n,c): I = sum p*t, Pn = sum p, Tn = sum t;
// loss = -mean_{n,c} 2I / (Pn + Tn + 1e-5). sums layout [N][C][3] double.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dice_fwd_kernel(const float* __restrict__ z, const float* __restrict__ label,
                                                       double* __restrict__ sums, int HW, int C) {
  __shared__ float red[3 * kMaxC][8];
  const long long n = blockIdx.y;
  float I[kMaxC], Ps[kMaxC], Ts[kMaxC];
#pragma unroll
  for (int c = 0; c < kMaxC; ++c) { I[c] = 0.f; Ps[c] = 0.f; Ts[c] = 0.f; }
  for (int hw = blockIdx.x * blockDim.x + threadIdx.x; hw < HW; hw += gridDim.x * blockDim.x) {
    const long long off = n * C * HW + hw;
    Softmax s;
    load_softmax(z + off, HW, C, s);
#pragma unroll
    for (int c = 0; c < kMaxC; ++c)
      if (c < C) {
        const float t = label[off + c * static_cast<long long>(HW)];
        I[c] = fmaf(s.p[c], t, I[c]);
        Ps[c] += s.p[c];
        Ts[c] += t;
      }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < kMaxC; ++c) {
    const float a = warp_sum(I[c]), b = warp_sum(Ps[c]), d = warp_sum(Ts[c]);
    if (lane == 0) { red[c][warp] = a; red[kMaxC + c][warp] = b; red[2 * kMaxC + c][warp] = d; }
  }
  __syncthreads();
  if (threadIdx.x < 3 * kMaxC) {
    const int k = threadIdx.x / kMaxC, c = threadIdx.x % kMaxC;
    if (c < C) {
      double t = 0.0;
      for (int w = 0; w < 8; ++w) t += static_cast<double>(red[threadIdx.x][w]);
      atomicAdd(sums + (n * C + c) * 3 + k, t);
    }
  }
}
// coef [N][C][2]: gp_c = A*t_c + B with A = -2/(D*N*C), B = 2I/(D^2*N*C)
__global__ void dice_finalize_kernel(const double* __restrict__ sums, float* __restrict__ loss, float* __restrict__ coef,
                                     int N, int C) {
  __shared__ double tot;
  if (threadIdx.x == 0) tot = 0.0;
  __syncthreads();
  double local = 0.0;
  const double nc = static_cast<double>(N) * C;
  for (int i = threadIdx.x; i < N * C; i += blockDim.x) {
    const double I = sums[i * 3], D = sums[i * 3 + 1] + sums[i * 3 + 2] + 1e-5;
    local += 2.0 * I / D;
    coef[i * 2] = static_cast<float>(-2.0 / (D * nc));
    coef[i * 2 + 1] = static_cast<float>(2.0 * I / (D * D * nc));
  }
  atomicAdd(&tot, local);
  __syncthreads();
  if (threadIdx.x == 0) *loss = static_cast<float>(-tot / nc);
}
__global__ void __launch_bounds__(256) dice_bwd_kernel(const float* __restrict__ z, const float* __restrict__ label,
                                                       const float* __restrict__ coef, const float* __restrict__ g,
                                                       float* __restrict__ dz, int HW, int C, int accumulate) {
  const long long n = blockIdx.y;
  const float gv = *g;
  for (int hw = blockIdx.x * blockDim.x + threadIdx.x; hw < HW; hw += gridDim.x * blockDim.x) {
    const long long off = n * C * HW + hw;
    Softmax s;
    load_softmax(z + off, HW, C, s);
    float gp[kMaxC], dot = 0.f;
#pragma unroll
    for (int c = 0; c < kMaxC; ++c) {
      gp[c] = 0.f;
      if (c < C) {
        const float t = label[off + c * static_cast<long long>(HW)];
        gp[c] = fmaf(coef[(n * C + c) * 2], t, coef[(n * C + c) * 2 + 1]);
        dot = fmaf(gp[c], s.p[c], dot);
      }
    }
#pragma unroll
    for (int c = 0; c < kMaxC; ++c)
      if (c < C) {
        const float v = gv * s.p[c] * (gp[c] - dot);
        float* o = dz + off + c * static_cast<long long>(HW);
        *o = accumulate ? *o + v : v;
      }
  }
}
int dice_fwd(const float* z, const float* label, double* sums, float* coef, float* loss, int N, int C, int HW,
             cudaStream_t s) {
  PP_REQUIRE(C >= 1 && C <= kMaxC, "dice: num_classes=%d unsupported", C);
  PP_CHECK_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 3 * N * C, s));
  int bx = ceil_div(HW, 256 * 4);
  if (bx < 1) bx = 1;
  dice_fwd_kernel<<<dim3(bx, N), 256, 0, s>>>(z, label, sums, HW, C);
  dice_finalize_kernel<<<1, 64, 0, s>>>(sums, loss, coef, N, C);
  PP_LAUNCH_CHECK_N(2);
  return PP_OK;
}
int dice_bwd(const float* z, const float* label, const float* coef, const float* g, float* dz, int N, int C, int HW,
             int accumulate, cudaStream_t s) {
  PP_REQUIRE(C >= 1 && C <= kMaxC, "dice: num_classes=%d unsupported", C);
  int bx = ceil_div(HW, 256 * 2);
  if (bx < 1) bx = 1;
  dice_bwd_kernel<<<dim3(bx, N), 256, 0, s>>>(z, label, coef, g, dz, HW, C, accumulate);
  PP_LAUNCH_CHECK();
  return PP_OK;
}

// ---------------------------------------------------------------------------------------------
// Validation Dice metric (utils/metrics.py:7-34 compute_dice, called per sample by train_chaos.py:386-390 on softmax
// values copied to the host): for every sample n and class c, with pred = argmax_c score (first maximum wins, as
// np.argmax) and the one-hot label t,   dice = 2 * sum[pred == c] * t_c / (sum[pred == c] + sum t_c + 1e-5),
// NaN when the class is absent from both prediction and label. One pass over scores + label for the whole batch.
// counts: [N][C][3] (intersection, predicted, labelled) as fp64 sums of the label VALUES (the reference multiplies by
// the label array, so non-binary labels behave the same), plus [N][C] flags "label plane has a non-zero entry".
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dice_metric_count_kernel(const float* __restrict__ scores,
                                                                const float* __restrict__ label,
                                                                double* __restrict__ counts,
                                                                unsigned int* __restrict__ nonzero, int HW, int C) {
  const int n = blockIdx.y;
  const float* sc = scores + static_cast<size_t>(n) * C * HW;
  const float* lb = label + static_cast<size_t>(n) * C * HW;
  float inter[kMaxC], predc[kMaxC], labc[kMaxC];
  unsigned int nz = 0;
#pragma unroll
  for (int c = 0; c < kMaxC; ++c) { inter[c] = 0.f; predc[c] = 0.f; labc[c] = 0.f; }
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += gridDim.x * blockDim.x) {
    int best = 0;
    float bv = sc[p];
#pragma unroll
    for (int c = 1; c < kMaxC; ++c)
      if (c < C) {
        const float v = sc[static_cast<size_t>(c) * HW + p];
        if (v > bv) { bv = v; best = c; }
      }
#pragma unroll
    for (int c = 0; c < kMaxC; ++c)
      if (c < C) {
        const float t = lb[static_cast<size_t>(c) * HW + p];
        const float pr = (c == best) ? 1.f : 0.f;
        inter[c] = fmaf(pr, t, inter[c]);
        predc[c] += pr;
        labc[c] += t;
        if (t != 0.f) nz |= 1u << c;
      }
  }
  __shared__ float red[3 * kMaxC][8];
  __shared__ unsigned int red_nz[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < kMaxC; ++c) {
    const float a = warp_sum(inter[c]), b = warp_sum(predc[c]), d = warp_sum(labc[c]);
    if (lane == 0) { red[c][warp] = a; red[kMaxC + c][warp] = b; red[2 * kMaxC + c][warp] = d; }
  }
  nz = __reduce_or_sync(0xffffffffu, nz);
  if (lane == 0) red_nz[warp] = nz;
  __syncthreads();
  if (threadIdx.x < 3 * kMaxC) {
    const int k = threadIdx.x / kMaxC, c = threadIdx.x % kMaxC;
    if (c < C) {
      double t = 0.0;
      for (int w = 0; w < 8; ++w) t += static_cast<double>(red[threadIdx.x][w]);
      atomicAdd(counts + (static_cast<size_t>(n) * C + c) * 3 + k, t);
    }
  }
  if (threadIdx.x == 0) {
    unsigned int all = 0;
    for (int w = 0; w < 8; ++w) all |= red_nz[w];
    for (int c = 0; c < C; ++c)
      if (all & (1u << c)) atomicOr(nonzero + n * C + c, 1u);
  }
}
__global__ void dice_metric_finalize_kernel(const double* __restrict__ counts, const unsigned int* __restrict__ nonzero,
                                            float* __restrict__ dice, int NC) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= NC) return;
  const double inter = counts[i * 3], pred = counts[i * 3 + 1], lab = counts[i * 3 + 2];
  if (pred == 0.0 && nonzero[i] == 0u) dice[i] = nanf("");                       // absent from both: skipped by the caller
  else dice[i] = static_cast<float>(2.0 * inter / (pred + lab + 1e-5));
}
// scratch: 3*N*C doubles + N*C uint32 (zeroed here)
int dice_metric(const float* scores, const float* label, float* dice, void* scratch, int N, int C, int HW,
                cudaStream_t s) {
  PP_REQUIRE(C >= 1 && C <= kMaxC, "dice_metric: num_classes=%d unsupported (max %d)", C, kMaxC);
  PP_REQUIRE(N >= 1 && N <= 65535 && HW >= 1, "dice_metric: bad shape N=%d HW=%d", N, HW);
  double* counts = static_cast<double*>(scratch);
  unsigned int* nonzero = reinterpret_cast<unsigned int*>(counts + 3LL * N * C);
  PP_CHECK_CUDA(cudaMemsetAsync(scratch, 0, sizeof(double) * 3 * N * C + sizeof(unsigned int) * N * C, s));
  int bx = ceil_div(HW, 256 * 4);
  if (bx > 64) bx = 64;
  dice_metric_count_kernel<<<dim3(bx, N), 256, 0, s>>>(scores, label, counts, nonzero, HW, C);
  dice_metric_finalize_kernel<<<ceil_div(N * C, 128), 128, 0, s>>>(counts, nonzero, dice, N * C);
  PP_LAUNCH_CHECK_N(2);
  return PP_OK;
}

// ---------------------------------------------------------------------------------------------
// Memory-bank update (aux_path_memory.py:68-116). Only sample 0 of the batch is visited (the
// reference returns from inside its per-sample loop). One block per class; a warp owns one
// labelled pixel at a time: lanes split the hidden channels, the 8x bilinear upsample of the
// 32x32 feature map is evaluated on the fly at that pixel only.
// ---------------------------------------------------------------------------------------------
constexpr int kMaxHidPerLane = 8;  // hid_ch <= 256
constexpr int kMemSlices = 32;     // blocks per class in the gather phase

// phase 1: grid (kMemSlices, C). Each block scans a slice of sample 0's scribble plane of its class and
// writes its partial sums [U(hid) | E(hid) | S | count] to scratch[cls][slice][2*hid+2].
template <typename T>
__global__ void __launch_bounds__(256)
memory_gather_kernel(const T* __restrict__ feat /*[N][h][w][hid], sample 0 used*/, const float* __restrict__ scribble
                     /*[N][K][H][W], sample 0 used; or nullptr*/, const uint8_t* __restrict__ scribble_idx
                     /*[N][H][W] class index map, sample 0 used*/, const float* __restrict__ bank /*[C][hid]*/,
                     float* __restrict__ scratch, int h, int w, int H, int W, int hid, float sh, float sw) {
  __shared__ float s_rhat[256];
  __shared__ float s_U[8][256], s_E[8][256];
  __shared__ float s_S[8], s_cnt[8], s_red[8];
  const int cls = blockIdx.y, slice = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int R = hid / 32;
  float nrm_part = 0.f;
  for (int k = threadIdx.x; k < hid; k += blockDim.x) {
    const float v = bank[cls * hid + k];
    s_rhat[k] = v;
    nrm_part = fmaf(v, v, nrm_part);
  }
  nrm_part = warp_sum(nrm_part);
  if (lane == 0) s_red[warp] = nrm_part;
  __syncthreads();
  float nrm2 = 0.f;
  for (int i = 0; i < 8; ++i) nrm2 += s_red[i];
  const float rinv = 1.f / (sqrtf(nrm2) + 1e-8f);
  __syncthreads();
  for (int k = threadIdx.x; k < hid; k += blockDim.x) s_rhat[k] *= rinv;
  __syncthreads();

  float U[kMaxHidPerLane], E[kMaxHidPerLane];
#pragma unroll
  for (int r = 0; r < kMaxHidPerLane; ++r) { U[r] = 0.f; E[r] = 0.f; }
  float S = 0.f, cnt = 0.f;
  const float* plane = scribble ? scribble + static_cast<long long>(cls) * H * W : nullptr;  // sample 0, channel cls
  const int HWp = H * W;
  const int per = (HWp + kMemSlices - 1) / kMemSlices;
  const int pbeg = slice * per, pend = min(pbeg + per, HWp);
  for (int base = pbeg + warp * 32; base < pend; base += 8 * 32) {
    const int pix = base + lane;
    const bool hit = (pix < pend) && (plane ? plane[pix] == 1.f : scribble_idx[pix] == cls);
    unsigned bal = __ballot_sync(0xffffffffu, hit);
    while (bal) {
      const int b = __ffs(bal) - 1;
      bal &= bal - 1;
      const int q = base + b;
      const int Y = q / W, X = q % W;
      // bilinear sample (align_corners=True) of the hid-dim feature at (Y, X)
      const float ry = sh * Y, rx = sw * X;
      const int y0 = static_cast<int>(ry), x0 = static_cast<int>(rx);
      const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
      const float wy1 = ry - y0, wy0 = 1.f - wy1, wx1 = rx - x0, wx0 = 1.f - wx1;
      float e[kMaxHidPerLane];
      float n2 = 0.f;
#pragma unroll
      for (int r = 0; r < kMaxHidPerLane; ++r)
        if (r < R) {
          const int k = r * 32 + lane;
          const float f00 = to_f32(feat[(static_cast<long long>(y0) * w + x0) * hid + k]);
          const float f01 = to_f32(feat[(static_cast<long long>(y0) * w + x1) * hid + k]);
          const float f10 = to_f32(feat[(static_cast<long long>(y1) * w + x0) * hid + k]);
          const float f11 = to_f32(feat[(static_cast<long long>(y1) * w + x1) * hid + k]);
          e[r] = wy0 * (wx0 * f00 + wx1 * f01) + wy1 * (wx0 * f10 + wx1 * f11);
          n2 = fmaf(e[r], e[r], n2);
          E[r] += e[r];
        }
      n2 = warp_sum(n2);
      const float einv = 1.f / (sqrtf(n2) + 1e-8f);
      float dot = 0.f;
#pragma unroll
      for (int r = 0; r < kMaxHidPerLane; ++r)
        if (r < R) { e[r] *= einv; dot = fmaf(e[r], s_rhat[r * 32 + lane], dot); }
      dot = warp_sum(dot);
      const float wgt = 1.f - dot;
      S += wgt;
      cnt += 1.f;
#pragma unroll
      for (int r = 0; r < kMaxHidPerLane; ++r)
        if (r < R) U[r] = fmaf(e[r], wgt, U[r]);
    }
  }
#pragma unroll
  for (int r = 0; r < kMaxHidPerLane; ++r)
    if (r < R) { s_U[warp][r * 32 + lane] = U[r]; s_E[warp][r * 32 + lane] = E[r]; }
  if (lane == 0) { s_S[warp] = S; s_cnt[warp] = cnt; }
  __syncthreads();
  float* out = scratch + (static_cast<long long>(cls) * kMemSlices + slice) * (2 * hid + 2);
  for (int k = threadIdx.x; k < hid; k += blockDim.x) {
    float Ut = 0.f, Et = 0.f;
    for (int i = 0; i < 8; ++i) { Ut += s_U[i][k]; Et += s_E[i][k]; }
    out[k] = Ut;
    out[hid + k] = Et;
  }
  if (threadIdx.x == 0) {
    float St = 0.f, ct = 0.f;
    for (int i = 0; i < 8; ++i) { St += s_S[i]; ct += s_cnt[i]; }
    out[2 * hid] = St;
    out[2 * hid + 1] = ct;
  }
}

// phase 2: grid C. Fixed-order reduction over the slices, then the update rule of aux_path_memory.py:84-114.
__global__ void memory_apply_kernel(const float* __restrict__ scratch, float* __restrict__ bank, int hid, int cosine_mode,
                                    float m, float one_minus_m) {
  __shared__ float s_red[8];
  __shared__ int s_nonzero;
  const int cls = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_nonzero = 0;
  __syncthreads();
  const int k = threadIdx.x;  // blockDim.x == 256 >= hid
  const float row = k < hid ? bank[cls * hid + k] : 0.f;
  if (row != 0.f) s_nonzero = 1;
  float part = warp_sum(row * row);
  if (lane == 0) s_red[warp] = part;
  __syncthreads();
  float nrm2 = 0.f;
  for (int i = 0; i < 8; ++i) nrm2 += s_red[i];
  const float rhat = row / (sqrtf(nrm2) + 1e-8f);
  const float* in = scratch + static_cast<long long>(cls) * kMemSlices * (2 * hid + 2);
  float St = 0.f, ct = 0.f, Ut = 0.f, Et = 0.f;
  for (int sl = 0; sl < kMemSlices; ++sl) {
    const float* q = in + static_cast<long long>(sl) * (2 * hid + 2);
    St += q[2 * hid];
    ct += q[2 * hid + 1];
    if (k < hid) { Ut += q[k]; Et += q[hid + k]; }
  }
  if (ct == 0.f || k >= hid) return;  // class absent in sample 0: row untouched
  float out;
  if (!s_nonzero) out = Et / ct;                                             // first touch: plain mean
  else if (cosine_mode) out = one_minus_m * rhat + m * (Ut / (St + 1e-8f));
  else out = one_minus_m * row + m * (Et / ct);
  bank[cls * hid + k] = out;
}

int memory_update_scratch_floats(int C, int hid) { return C * kMemSlices * (2 * hid + 2); }

int memory_update(int dtype, const void* feat, const float* scribble, float* bank, float* scratch, int C, int h, int w,
                  int H, int W, int hid, int cosine_mode, float m, float one_minus_m, cudaStream_t s,
                  const uint8_t* scribble_idx) {
  PP_REQUIRE(hid % 32 == 0 && hid <= 256, "memory_update: hid_ch=%d unsupported (multiple of 32, <= 256)", hid);
  PP_REQUIRE(scratch != nullptr, "memory_update: scratch buffer required");
  PP_REQUIRE((scribble != nullptr) != (scribble_idx != nullptr), "memory_update: one-hot OR index-map scribble");
  const float sh = H > 1 ? static_cast<float>(h - 1) / static_cast<float>(H - 1) : 0.f;
  const float sw = W > 1 ? static_cast<float>(w - 1) / static_cast<float>(W - 1) : 0.f;
  dim3 grid(kMemSlices, C);
  if (dtype == PP_F32)
    memory_gather_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float*>(feat), scribble, scribble_idx, bank,
                                                     scratch, h, w, H, W, hid, sh, sw);
  else
    memory_gather_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(feat), scribble,
                                                             scribble_idx, bank, scratch, h, w, H, W, hid, sh, sw);
  memory_apply_kernel<<<C, 256, 0, s>>>(scratch, bank, hid, cosine_mode, m, one_minus_m);
  PP_LAUNCH_CHECK_N(2);
  return PP_OK;
}

// ---------------------------------------------------------------------------------------------
// Bank classification loss: logits = bank (C x hid) . Wfc^T (C x hid), CE against arange(C).
// (aux_path_memory.py:61 fc_cls(memory_bank); consistency_reglur_memory.py:94.)
// Forward writes the loss and the softmax (C x C) for backward.
// ---------------------------------------------------------------------------------------------
__global__ void memory_loss_fwd_kernel(const float* __restrict__ bank, const float* __restrict__ wfc,
                                       float* __restrict__ loss, float* __restrict__ probs, int C, int hid) {
  __shared__ float lg[kMaxC][kMaxC];
  const int t = threadIdx.x;
  if (t < C * C) {
    const int i = t / C, j = t % C;
    float a = 0.f;
    for (int k = 0; k < hid; ++k) a = fmaf(bank[i * hid + k], wfc[j * hid + k], a);
    lg[i][j] = a;
  }
  __syncthreads();
  if (t == 0) {
    float tot = 0.f;
    for (int i = 0; i < C; ++i) {
      float mx = -INFINITY;
      for (int j = 0; j < C; ++j) mx = fmaxf(mx, lg[i][j]);
      float sum = 0.f;
      for (int j = 0; j < C; ++j) sum += expf(lg[i][j] - mx);
      const float lse = mx + logf(sum);
      for (int j = 0; j < C; ++j) probs[i * C + j] = expf(lg[i][j] - lse);
      tot += lse - lg[i][i];
    }
    *loss = tot / C;
  }
}
// dWfc[j][k] += g/C * sum_i (probs[i][j] - [i==j]) * bank[i][k]
__global__ void memory_loss_bwd_kernel(const float* __restrict__ bank, const float* __restrict__ probs,
                                       const float* __restrict__ g, float* __restrict__ dwfc, int C, int hid) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= C * hid) return;
  const int j = idx / hid, k = idx % hid;
  float a = 0.f;
  for (int i = 0; i < C; ++i) a = fmaf(probs[i * C + j] - (i == j ? 1.f : 0.f), bank[i * hid + k], a);
  dwfc[idx] += (*g) * a / C;
}
int memory_loss_fwd(const float* bank, const float* wfc, float* loss, float* probs, int C, int hid, cudaStream_t s) {
  PP_REQUIRE(C >= 1 && C <= kMaxC, "memory_loss: num_classes=%d unsupported", C);
  memory_loss_fwd_kernel<<<1, 64, 0, s>>>(bank, wfc, loss, probs, C, hid);
  PP_LAUNCH_CHECK();
  return PP_OK;
}
int memory_loss_bwd(const float* bank, const float* probs, const float* g, float* dwfc, int C, int hid,
                    cudaStream_t s) {
  memory_loss_bwd_kernel<<<ceil_div(C * hid, 128), 128, 0, s>>>(bank, probs, g, dwfc, C, hid);
  PP_LAUNCH_CHECK();
  return PP_OK;
}

}  // namespace pp
