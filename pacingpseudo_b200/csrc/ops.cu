// ops.cu — the HBM-bound operators around the convolutions, all on NHWC activations of type T
// (bf16 by default, fp32 in the fp32 precision mode):
//   weight (un)packing, first conv (Cin = 1), 1x1 heads, BatchNorm statistics / finalize / apply /
//   backward (train-mode batch statistics per statistics group, or eval-mode running statistics),
//   LeakyReLU(0.01), MaxPool2d(2,2), bilinear align_corners=True upsampling, NCHW<->NHWC converts,
//   Adam. Reference call sites: /root/reference/models/unet.py:60,109,144,188-193,
//   /root/reference/models/aux_path_memory.py:22-33,52 and /root/reference/train_chaos.py:219.
#include "pp_common.cuh"
#include "pp_ops.h"
#include <stdlib.h>

namespace pp {

static inline int grid_for(long long work, int block, int max_blocks_per_sm = 16) {
  long long g = ceil_div_ll(work, block);
  long long cap = static_cast<long long>(sm_count()) * max_blocks_per_sm;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

// Block size of the BatchNorm apply / backward kernels. These HBM-bound kernels run CONCURRENTLY with the conv CTAs of
// another stream (unet_plan.cu); a conv CTA holds ~36K of an SM's 64K registers, so small blocks (128 threads x <=128
// registers = 16K) are what still fits beside it.
static constexpr int kBnThreads = 128;

// Hot loops index with 32-bit integers (a 64-bit div/mod costs ~100 instructions per element on the GPU); the
// launchers check that the element counts fit.
#define PP_REQUIRE_INT32(v, what) \
  PP_REQUIRE(static_cast<long long>(v) < (1LL << 31), "%s: %lld elements exceed the 32-bit index range", what, \
             static_cast<long long>(v))

#define PP_DISPATCH_T(dtype, ...)                        \
  do {                                                   \
    if ((dtype) == PP_F32) { using T = float; __VA_ARGS__ } \
    else { using T = __nv_bfloat16; __VA_ARGS__ }        \
  } while (0)

// ==============================================================================================
// Weight packing: OIHW fp32 [Cout][Cin][3][3]  ->  wf [tap][Cout][Cin]  and  wd [8-tap][Cin][Cout]
// (wd = spatially flipped + transposed weights: dgrad runs through the forward kernel).
// ==============================================================================================
template <typename T>
__global__ void __launch_bounds__(256) pack_weights_kernel(const float* __restrict__ w, T* __restrict__ wf,
                                                           T* __restrict__ wd, int Cout, int Cin) {
  __shared__ float tile[32][32 * 9 + 1];
  const int co0 = blockIdx.y * 32, ci0 = blockIdx.x * 32;
  for (int i = threadIdx.x; i < 32 * 288; i += 256) {
    const int r = i / 288, c = i % 288;  // r: co in tile, c: (ci, tap) contiguous in OIHW
    tile[r][c] = w[(static_cast<long long>(co0 + r) * Cin + ci0) * 9 + c];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 9 * 32 * 32; i += 256) {
    const int tap = i / 1024, a = (i / 32) % 32, b = i % 32;
    // wf: rows = co (a), contiguous ci (b)
    wf[(static_cast<long long>(tap) * Cout + co0 + a) * Cin + ci0 + b] = from_f32<T>(tile[a][b * 9 + tap]);
    // wd: rows = ci (a), contiguous co (b)
    wd[(static_cast<long long>(8 - tap) * Cin + ci0 + a) * Cout + co0 + b] = from_f32<T>(tile[b][a * 9 + tap]);
  }
}

int pack_weights(int dtype, const float* w, void* wf, void* wd, int Cout, int Cin, cudaStream_t s) {
  PP_REQUIRE(Cout % 32 == 0 && Cin % 32 == 0, "pack_weights: Cout=%d Cin=%d must be multiples of 32", Cout, Cin);
  PP_DISPATCH_T(dtype, pack_weights_kernel<T><<<dim3(Cin / 32, Cout / 32), 256, 0, s>>>(
                           w, static_cast<T*>(wf), static_cast<T*>(wd), Cout, Cin););
  PP_LAUNCH_CHECK();
  return PP_OK;
}

// All conv layers of a network in ONE launch (the per-layer kernels are launch-latency bound: 22 launches of a few
// microseconds each per forward pass). The layer table travels as a kernel parameter.
constexpr int kMaxPackLayers = 32;
struct PackTable {
  const float* w[kMaxPackLayers];
  void* wf[kMaxPackLayers];
  void* wd[kMaxPackLayers];
  int cout[kMaxPackLayers], cin[kMaxPackLayers];
  int tile_begin[kMaxPackLayers + 1];   // prefix sum of 32x32 tiles per layer
  int n;
};
template <typename T>
__global__ void __launch_bounds__(256) pack_weights_multi_kernel(const __grid_constant__ PackTable tb) {
  __shared__ float tile[32][32 * 9 + 1];
  int l = 0;
  while (l + 1 < tb.n && static_cast<int>(blockIdx.x) >= tb.tile_begin[l + 1]) ++l;
  const int Cout = tb.cout[l], Cin = tb.cin[l];
  const int tidx = blockIdx.x - tb.tile_begin[l];
  const int tiles_ci = Cin / 32;
  const int co0 = (tidx / tiles_ci) * 32, ci0 = (tidx % tiles_ci) * 32;
  const float* __restrict__ w = tb.w[l];
  T* __restrict__ wf = static_cast<T*>(tb.wf[l]);
  T* __restrict__ wd = static_cast<T*>(tb.wd[l]);
  for (int i = threadIdx.x; i < 32 * 288; i += 256) {
    const int r = i / 288, c = i % 288;
    tile[r][c] = w[(static_cast<long long>(co0 + r) * Cin + ci0) * 9 + c];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 9 * 32 * 32; i += 256) {
    const int tap = i / 1024, a = (i / 32) % 32, b = i % 32;
    wf[(static_cast<long long>(tap) * Cout + co0 + a) * Cin + ci0 + b] = from_f32<T>(tile[a][b * 9 + tap]);
    wd[(static_cast<long long>(8 - tap) * Cin + ci0 + a) * Cout + co0 + b] = from_f32<T>(tile[b][a * 9 + tap]);
  }
}

int pack_weights_multi(int dtype, int n, const float* const* w, void* const* wf, void* const* wd, const int* cout,
                       const int* cin, cudaStream_t s) {
  for (int base = 0; base < n; base += kMaxPackLayers) {
    PackTable tb{};
    tb.n = n - base < kMaxPackLayers ? n - base : kMaxPackLayers;
    tb.tile_begin[0] = 0;
    for (int i = 0; i < tb.n; ++i) {
      const int k = base + i;
      PP_REQUIRE(cout[k] % 32 == 0 && cin[k] % 32 == 0, "pack_weights: Cout=%d Cin=%d must be multiples of 32", cout[k],
                 cin[k]);
      tb.w[i] = w[k]; tb.wf[i] = wf[k]; tb.wd[i] = wd[k]; tb.cout[i] = cout[k]; tb.cin[i] = cin[k];
      tb.tile_begin[i + 1] = tb.tile_begin[i] + (cout[k] / 32) * (cin[k] / 32);
    }
    if (tb.tile_begin[tb.n] == 0) continue;
    PP_DISPATCH_T(dtype, pack_weights_multi_kernel<T><<<tb.tile_begin[tb.n], 256, 0, s>>>(tb););
    PP_LAUNCH_CHECK();
  }
  return PP_OK;
}

// dwp [tap][Cout][Cin] fp32 -> OIHW grad [Cout][Cin][3][3] (accumulate ? += : =), for input channels
// [ci_begin, ci_begin + ci_count) only (multiples of 32).
__global__ void __launch_bounds__(256) unpack_wgrad_kernel(const float* __restrict__ dwp, float* __restrict__ g,
                                                           int Cout, int Cin, int ci_begin, int accumulate) {
  __shared__ float tile[32][32 * 9 + 1];
  const int co0 = blockIdx.y * 32, ci0 = ci_begin + blockIdx.x * 32;
  for (int i = threadIdx.x; i < 9 * 32 * 32; i += 256) {
    const int tap = i / 1024, a = (i / 32) % 32, b = i % 32;
    tile[a][b * 9 + tap] = dwp[(static_cast<long long>(tap) * Cout + co0 + a) * Cin + ci0 + b];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 32 * 288; i += 256) {
    const int r = i / 288, c = i % 288;
    float* o = g + (static_cast<long long>(co0 + r) * Cin + ci0) * 9 + c;
    *o = accumulate ? (*o + tile[r][c]) : tile[r][c];
  }
}

// small ranges (the narrow layers): one thread per OIHW element, strided gather from the packed layout
__global__ void __launch_bounds__(256) unpack_wgrad_small_kernel(const float* __restrict__ dwp, float* __restrict__ g,
                                                                 int Cout, int Cin, int ci_begin, int ci_count,
                                                                 int accumulate) {
  const int total = Cout * ci_count * 9;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int tap = i % 9, ci = ci_begin + (i / 9) % ci_count, co = i / (9 * ci_count);
  const float v = dwp[(static_cast<size_t>(tap) * Cout + co) * Cin + ci];
  float* o = g + (static_cast<size_t>(co) * Cin + ci) * 9 + tap;
  *o = accumulate ? *o + v : v;
}

int unpack_wgrad_range(const float* dwp, float* g, int Cout, int Cin, int ci_begin, int ci_count, int accumulate,
                       cudaStream_t s) {
  PP_REQUIRE(Cout % 32 == 0 && ci_begin % 32 == 0 && ci_count % 32 == 0 && ci_begin + ci_count <= Cin,
             "unpack_wgrad: Cout=%d Cin=%d range [%d,+%d) must be multiples of 32", Cout, Cin, ci_begin, ci_count);
  if (Cout * ci_count <= 64 * 64) {
    unpack_wgrad_small_kernel<<<ceil_div(Cout * ci_count * 9, 256), 256, 0, s>>>(dwp, g, Cout, Cin, ci_begin, ci_count,
                                                                                accumulate);
    PP_LAUNCH_CHECK();
    return PP_OK;
  }
  unpack_wgrad_kernel<<<dim3(ci_count / 32, Cout / 32), 256, 0, s>>>(dwp, g, Cout, Cin, ci_begin, accumulate);
  PP_LAUNCH_CHECK();
  return PP_OK;
}

int unpack_wgrad(const float* dwp, float* g, int Cout, int Cin, int accumulate, cudaStream_t s) {
  return unpack_wgrad_range(dwp, g, Cout, Cin, 0, Cin, accumulate, s);
}

// ==============================================================================================
// First conv: Cin = 1, 3x3, pad 1 (unet.py:28 enc_block1.conv_layer1). Direct convolution.
// Thread = (pixel, group of 8 output channels); the group is fixed per thread, so its 72 weights + 8 biases
// live in registers and the inner loop is 9 broadcast loads of x, 72 FMAs and one 16-byte store.
// ==============================================================================================
template <typename T>
__global__ void __launch_bounds__(256) first_conv_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                             const float* __restrict__ bias, T* __restrict__ y, int N,
                                                             int H, int W, int Cout, const ConvAffine af) {
  const int vecs = Cout / 8;                 // power of two, <= 32 (checked by the launcher)
  const int v = threadIdx.x % vecs;
  const int ppb = 256 / vecs;                // pixels per block iteration
  float wr[8][9], br[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {   // eval-mode BatchNorm folded in: the per-channel scale goes into the weights
    const float sc = af.scale ? af.scale[v * 8 + j] : 1.f;
    br[j] = af.scale ? af.shift[v * 8 + j] : (bias ? bias[v * 8 + j] : 0.f);
#pragma unroll
    for (int t = 0; t < 9; ++t) wr[j][t] = w[(v * 8 + j) * 9 + t] * sc;
  }
  const int P = N * H * W;
  for (int p = blockIdx.x * ppb + threadIdx.x / vecs; p < P; p += gridDim.x * ppb) {
    const int px = p % W, row = p / W, py = row % H;
    const float* xr = x + static_cast<size_t>(row) * W + px;   // row = img * H + py
    float xin[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int dy = t / 3 - 1, dx = t % 3 - 1;
      const bool ok = (py + dy >= 0) && (py + dy < H) && (px + dx >= 0) && (px + dx < W);
      xin[t] = ok ? __ldg(xr + dy * W + dx) : 0.f;
    }
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float a = br[j];
#pragma unroll
      for (int t = 0; t < 9; ++t) a = fmaf(xin[t], wr[j][t], a);
      o[j] = af.scale ? lrelu(a, af.slope) : a;
    }
    Vec8<T> pk;
    pk.set(o);
    pk.store(y + static_cast<size_t>(p) * Cout + v * 8);
  }
}

// Run variant (W % 4 == 0): four consecutive pixels of a row per thread share one 3 x 6 input window (18 loads issued
// together instead of 4 x 9 in four dependent rounds); the per-pixel kernel above is latency-bound (41 us for 50 MB).
template <typename T>
__global__ void __launch_bounds__(256, 2) first_conv_fwd_run4_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                  const float* __restrict__ bias, T* __restrict__ y,
                                                                  int N, int H, int W, int Cout, const ConvAffine af) {
  const int vecs = Cout / 8;
  const int v = threadIdx.x % vecs;
  const int rpb = 256 / vecs;                // pixel runs per block iteration
  float wr[8][9], br[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float sc = af.scale ? af.scale[v * 8 + j] : 1.f;
    br[j] = af.scale ? af.shift[v * 8 + j] : (bias ? bias[v * 8 + j] : 0.f);
#pragma unroll
    for (int t = 0; t < 9; ++t) wr[j][t] = w[(v * 8 + j) * 9 + t] * sc;
  }
  const int runs = (N * H * W) >> 2, runs_per_row = W >> 2;
  for (int r = blockIdx.x * rpb + threadIdx.x / vecs; r < runs; r += gridDim.x * rpb) {
    const int row = r / runs_per_row, px = (r - row * runs_per_row) << 2, py = row % H;
    const float* xr = x + static_cast<size_t>(row) * W + px;
    float xin[3][6];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const bool rok = (py + a - 1 >= 0) && (py + a - 1 < H);
#pragma unroll
      for (int b = 0; b < 6; ++b) {
        const bool ok = rok && (px + b - 1 >= 0) && (px + b - 1 < W);
        xin[a][b] = ok ? __ldg(xr + (a - 1) * W + (b - 1)) : 0.f;
      }
    }
    T* yo = y + (static_cast<size_t>(row) * W + px) * Cout + v * 8;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float a = br[j];
#pragma unroll
        for (int t = 0; t < 9; ++t) a = fmaf(xin[t / 3][u + t % 3], wr[j][t], a);
        o[j] = af.scale ? lrelu(a, af.slope) : a;
      }
      Vec8<T> pk;
      pk.set(o);
      pk.store(yo + static_cast<size_t>(u) * Cout);
    }
  }
}

// Multi-channel input (input_ch > 1, unet.py:28 with --input_ch): x is NCHW fp32 [N][Cin][H][W], w OIHW [Cout][Cin][3][3].
// Same thread mapping as the single-channel kernels, the input channels are walked in a loop with the weights in shared
// memory (forward) / one grid.y slice per input channel (weight gradient). Not a tuned path: every shipped dataset is
// single-channel.
template <typename T>
__global__ void __launch_bounds__(256) first_conv_fwd_cin_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                 const float* __restrict__ bias, T* __restrict__ y, int N,
                                                                 int H, int W, int Cout, int Cin, const ConvAffine af) {
  extern __shared__ float s_wc[];             // [Cout][Cin][9]
  for (int i = threadIdx.x; i < Cout * Cin * 9; i += blockDim.x) s_wc[i] = w[i];
  __syncthreads();
  const int vecs = Cout / 8;
  const int v = threadIdx.x % vecs;
  const int ppb = 256 / vecs;
  const int P = N * H * W, HW = H * W;
  for (int p = blockIdx.x * ppb + threadIdx.x / vecs; p < P; p += gridDim.x * ppb) {
    const int n = p / HW, hw = p - n * HW, py = hw / W, px = hw - py * W;
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = bias ? bias[v * 8 + j] : 0.f;
    for (int ci = 0; ci < Cin; ++ci) {
      const float* xr = x + (static_cast<size_t>(n) * Cin + ci) * HW + hw;
      float xin[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int dy = t / 3 - 1, dx = t % 3 - 1;
        const bool ok = (py + dy >= 0) && (py + dy < H) && (px + dx >= 0) && (px + dx < W);
        xin[t] = ok ? __ldg(xr + dy * W + dx) : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float* wj = s_wc + (static_cast<size_t>(v * 8 + j) * Cin + ci) * 9;
#pragma unroll
        for (int t = 0; t < 9; ++t) o[j] = fmaf(xin[t], wj[t], o[j]);
      }
    }
    if (af.scale) {   // shift already holds bias * scale (bias is null in this mode)
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = lrelu(fmaf(o[j], af.scale[v * 8 + j], af.shift[v * 8 + j]), af.slope);
    }
    Vec8<T> pk;
    pk.set(o);
    pk.store(y + static_cast<size_t>(p) * Cout + v * 8);
  }
}
// grid.y = input channel
template <typename T>
__global__ void __launch_bounds__(256) first_conv_wgrad_cin_kernel(const T* __restrict__ dy, const float* __restrict__ x,
                                                                   float* __restrict__ dw, int N, int H, int W, int Cout,
                                                                   int Cin) {
  extern __shared__ float sacc[];  // [Cout*9]
  for (int i = threadIdx.x; i < Cout * 9; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  const int ci = blockIdx.y;
  const int vecs = Cout / 8;
  const int v = threadIdx.x % vecs;
  const int ppb = 256 / vecs;
  float acc[8][9];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[j][t] = 0.f;
  const int P = N * H * W, HW = H * W;
  for (int p = blockIdx.x * ppb + threadIdx.x / vecs; p < P; p += gridDim.x * ppb) {
    const int n = p / HW, hw = p - n * HW, py = hw / W, px = hw - py * W;
    Vec8<T> g;
    g.load(dy + static_cast<size_t>(p) * Cout + v * 8);
    float f[8];
    g.get(f);
    const float* xr = x + (static_cast<size_t>(n) * Cin + ci) * HW + hw;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int ddy = t / 3 - 1, ddx = t % 3 - 1;
      const bool ok = (py + ddy >= 0) && (py + ddy < H) && (px + ddx >= 0) && (px + ddx < W);
      const float xv = ok ? __ldg(xr + ddy * W + ddx) : 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j][t] = fmaf(f[j], xv, acc[j][t]);
    }
  }
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      float a = acc[j][t];
      for (int o = vecs; o < 32; o <<= 1) a += __shfl_xor_sync(0xffffffffu, a, o);   // lanes with the same v
      if (lane < vecs) atomicAdd(&sacc[(v * 8 + j) * 9 + t], a);
    }
  __syncthreads();
  for (int i = threadIdx.x; i < Cout * 9; i += blockDim.x)
    atomicAdd(dw + (static_cast<size_t>(i / 9) * Cin + ci) * 9 + i % 9, sacc[i]);
}

static bool pow2_vecs(int C) { return C % 8 == 0 && C <= 256 && ((C / 8) & (C / 8 - 1)) == 0; }

int first_conv_fwd(int dtype, const float* x, const float* w, const float* bias, void* y, int N, int H, int W, int Cout,
                   cudaStream_t s, int Cin, const ConvAffine* affine) {
  PP_REQUIRE(pow2_vecs(Cout), "first_conv_fwd: Cout=%d unsupported (8,16,...,256)", Cout);
  ConvAffine af{nullptr, nullptr, 0.f};
  if (affine != nullptr && affine->scale != nullptr) {
    PP_REQUIRE(affine->shift != nullptr && bias == nullptr, "first_conv_fwd: affine epilogue needs shift and no bias");
    af = *affine;
  }
  PP_REQUIRE(Cin >= 1 && Cin <= 16, "first_conv_fwd: input_ch=%d unsupported (1..16)", Cin);
  const long long P = static_cast<long long>(N) * H * W;
  PP_REQUIRE_INT32(P * Cout, "first_conv_fwd");
  const int ppb = 256 / (Cout / 8);
  if (Cin > 1) {
    PP_DISPATCH_T(dtype, (first_conv_fwd_cin_kernel<T><<<grid_for(ceil_div_ll(P, ppb) * 256, 256, 8), 256,
                                                        sizeof(float) * Cout * Cin * 9, s>>>(x, w, bias, static_cast<T*>(y),
                                                                                             N, H, W, Cout, Cin, af)););
    PP_LAUNCH_CHECK();
    return PP_OK;
  }
  static const int run4_on = [] { const char* e = getenv("PP_FIRST_CONV_RUN4"); return (e && e[0] == '0') ? 0 : 1; }();
  if (run4_on && W % 4 == 0) {
    PP_DISPATCH_T(dtype, first_conv_fwd_run4_kernel<T><<<grid_for(ceil_div_ll(P / 4, ppb) * 256, 256, 8), 256, 0, s>>>(
                             x, w, bias, static_cast<T*>(y), N, H, W, Cout, af););
    PP_LAUNCH_CHECK();
    return PP_OK;
  }
  PP_DISPATCH_T(dtype, first_conv_fwd_kernel<T><<<grid_for(ceil_div_ll(P, ppb) * 256, 256, 8), 256, 0, s>>>(
                           x, w, bias, static_cast<T*>(y), N, H, W, Cout, af););
  PP_LAUNCH_CHECK();
  return PP_OK;
}

// dW[co][tap] += sum_p dy[p][co] * x[p + off(tap)]   (dw is the OIHW fp32 grad [Cout][1][3][3])
// Thread = (pixel, group of 8 channels): one 16-byte load of dy per pixel, 72 register accumulators; partial sums
// are combined by warp shuffles (lanes holding the same channel group), then shared memory, then one atomic per
// (block, weight).
template <typename T>
__global__ void __launch_bounds__(256) first_conv_wgrad_kernel(const T* __restrict__ dy, const float* __restrict__ x,
                                                               float* __restrict__ dw, int N, int H, int W, int Cout) {
  extern __shared__ float sacc[];  // [Cout*9]
  for (int i = threadIdx.x; i < Cout * 9; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  const int vecs = Cout / 8;
  const int v = threadIdx.x % vecs;
  const int ppb = 256 / vecs;
  float acc[8][9];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[j][t] = 0.f;
  const int P = N * H * W;
  constexpr int U = 2;   // pixels in flight per thread
  for (int p0 = blockIdx.x * ppb + threadIdx.x / vecs; p0 < P; p0 += gridDim.x * ppb * U) {
    Vec8<T> g[U];
    float xin[U][9];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int p = p0 + u * gridDim.x * ppb;
      if (p < P) {
        g[u].load(dy + static_cast<size_t>(p) * Cout + v * 8);
        const int px = p % W, row = p / W, py = row % H;
        const float* xr = x + static_cast<size_t>(row) * W + px;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const int ddy = t / 3 - 1, ddx = t % 3 - 1;
          const bool ok = (py + ddy >= 0) && (py + ddy < H) && (px + ddx >= 0) && (px + ddx < W);
          xin[u][t] = ok ? __ldg(xr + ddy * W + ddx) : 0.f;
        }
      } else {
        g[u].zero();
#pragma unroll
        for (int t = 0; t < 9; ++t) xin[u][t] = 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float f[8];
      g[u].get(f);
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int t = 0; t < 9; ++t) acc[j][t] = fmaf(f[j], xin[u][t], acc[j][t]);
    }
  }
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      float a = acc[j][t];
      for (int o = vecs; o < 32; o <<= 1) a += __shfl_xor_sync(0xffffffffu, a, o);   // lanes with the same v
      if (lane < vecs) atomicAdd(&sacc[(v * 8 + j) * 9 + t], a);
    }
  __syncthreads();
  for (int i = threadIdx.x; i < Cout * 9; i += blockDim.x) atomicAdd(dw + i, sacc[i]);
}
// Run variant (W % 4 == 0): a thread owns FOUR consecutive pixels of a row for its 8 channels, so the 3 x 6 input
// window is loaded once for the four pixels (18 instead of 36 scalar loads) and four 16-byte dy loads are in flight
// per thread. The kernel above keeps 16 KB of dy in flight per SM and is latency-bound (86 us for 100 MB); this is the
// last kernel of the backward pass (nothing overlaps it), so its duration is fully exposed.
template <typename T>
__global__ void __launch_bounds__(256) first_conv_wgrad_run4_kernel(const T* __restrict__ dy, const float* __restrict__ x,
                                                                    float* __restrict__ dw, int N, int H, int W,
                                                                    int Cout) {
  extern __shared__ float sacc[];  // [Cout*9]
  for (int i = threadIdx.x; i < Cout * 9; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  const int vecs = Cout / 8;
  const int v = threadIdx.x % vecs;
  const int rpb = 256 / vecs;                      // pixel runs per block iteration
  float acc[8][9];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[j][t] = 0.f;
  const int runs = (N * H * W) >> 2;
  const int runs_per_row = W >> 2;
  for (int r = blockIdx.x * rpb + threadIdx.x / vecs; r < runs; r += gridDim.x * rpb) {
    const int row = r / runs_per_row, px = (r - row * runs_per_row) << 2, py = row % H;
    Vec8<T> g[4];
    const T* gp = dy + (static_cast<size_t>(row) * W + px) * Cout + v * 8;
#pragma unroll
    for (int u = 0; u < 4; ++u) g[u].load(gp + static_cast<size_t>(u) * Cout);
    float xin[3][6];                               // rows py-1..py+1, columns px-1..px+4 (zero padding)
    const float* xr = x + static_cast<size_t>(row) * W + px;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const bool rok = (py + a - 1 >= 0) && (py + a - 1 < H);
#pragma unroll
      for (int b = 0; b < 6; ++b) {
        const bool ok = rok && (px + b - 1 >= 0) && (px + b - 1 < W);
        xin[a][b] = ok ? __ldg(xr + (a - 1) * W + (b - 1)) : 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float f[8];
      g[u].get(f);
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int t = 0; t < 9; ++t) acc[j][t] = fmaf(f[j], xin[t / 3][u + t % 3], acc[j][t]);
    }
  }
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      float a = acc[j][t];
      for (int o = vecs; o < 32; o <<= 1) a += __shfl_xor_sync(0xffffffffu, a, o);   // lanes with the same v
      if (lane < vecs) atomicAdd(&sacc[(v * 8 + j) * 9 + t], a);
    }
  __syncthreads();
  for (int i = threadIdx.x; i < Cout * 9; i += blockDim.x) atomicAdd(dw + i, sacc[i]);
}
int first_conv_wgrad(int dtype, const void* dy, const float* x, float* dw, int N, int H, int W, int Cout,
                     cudaStream_t s, int Cin) {
  PP_REQUIRE(pow2_vecs(Cout), "first_conv_wgrad: Cout=%d unsupported (8,16,...,256)", Cout);
  PP_REQUIRE(Cin >= 1 && Cin <= 16, "first_conv_wgrad: input_ch=%d unsupported (1..16)", Cin);
  PP_REQUIRE_INT32(static_cast<long long>(N) * H * W * Cout, "first_conv_wgrad");
  if (Cin > 1) {
    PP_DISPATCH_T(dtype, (first_conv_wgrad_cin_kernel<T><<<dim3(sm_count(), Cin), 256, Cout * 9 * sizeof(float), s>>>(
                             static_cast<const T*>(dy), x, dw, N, H, W, Cout, Cin)););
    PP_LAUNCH_CHECK();
    return PP_OK;
  }
  static const int run4_on = [] { const char* e = getenv("PP_FIRST_WGRAD_RUN4"); return (e && e[0] == '0') ? 0 : 1; }();
  if (run4_on && W % 4 == 0) {
    PP_DISPATCH_T(dtype, first_conv_wgrad_run4_kernel<T><<<sm_count() * 2, 256, Cout * 9 * sizeof(float), s>>>(
                             static_cast<const T*>(dy), x, dw, N, H, W, Cout););
    PP_LAUNCH_CHECK();
    return PP_OK;
  }
  const int blocks = sm_count() * 2;
  PP_DISPATCH_T(dtype, first_conv_wgrad_kernel<T><<<blocks, 256, Cout * 9 * sizeof(float), s>>>(
                           static_cast<const T*>(dy), x, dw, N, H, W, Cout););
  PP_LAUNCH_CHECK();
  return PP_OK;
}

// ==============================================================================================
// 1x1 heads (unet.py:60 final_conv with bias; aux_path_memory.py:32 fc_cls without bias).
// Input NHWC T [P][Cin], output logits NCHW fp32 [N][C][HW]. C <= 8.
// Thread = (pixel, group of 8 input channels): fully coalesced 16-byte loads, the class sums are combined across the
// CIN/8 lanes of a pixel with xor-shuffles; weights of the thread's channel group live in registers.
// ==============================================================================================
constexpr int kMaxClasses = 8;

template <typename T, int CIN, int NC>
__global__ void __launch_bounds__(256) head_fwd_kernel(const T* __restrict__ a, const float* __restrict__ w,
                                                       const float* __restrict__ bias, float* __restrict__ logits,
                                                       int P, int HW, int C) {
  constexpr int VECS = CIN / 8, PPB = 256 / VECS;
  const int v = threadIdx.x % VECS;
  float wr[NC][8], br[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    br[c] = (c < C && bias) ? bias[c] : 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) wr[c][j] = c < C ? w[c * CIN + v * 8 + j] : 0.f;
  }
  constexpr int U = 4;
  for (int p0 = blockIdx.x * PPB + threadIdx.x / VECS; p0 < P; p0 += gridDim.x * PPB * U) {
    Vec8<T> pk[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int p = p0 + u * gridDim.x * PPB;
      if (p < P) pk[u].load(a + static_cast<size_t>(p) * CIN + v * 8);
      else pk[u].zero();
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int p = p0 + u * gridDim.x * PPB;
      float f[8], acc[NC];
      pk[u].get(f);
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        float t = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) t = fmaf(f[j], wr[c][j], t);
#pragma unroll
        for (int o = 1; o < VECS; o <<= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        acc[c] = t + br[c];
      }
      if (p < P) {
        const int n = p / HW, hw = p % HW;
#pragma unroll
        for (int c = 0; c < NC; ++c)   // the VECS lanes of a pixel share the stores: lane v writes classes v, v+VECS, ..
          if (c < C && (c % VECS) == v) logits[(static_cast<size_t>(n) * C + c) * HW + hw] = acc[c];
      }
    }
  }
}

// Backward of the 1x1 head in ONE pass over dlogits and the activations:
//   da[p][ci] = sum_c dl[c][p] * w[c][ci]            (skipped when da == nullptr)
//   dW[c][ci] += sum_p dl[c][p] * a[p][ci];  db[c] += sum_p dl[c][p]
// The weights of the thread's channel group are read from shared memory (two broadcast LDS.128 per class) instead of
// 8*NC registers: with them in registers the kernel needed ~130 registers, ONE 256-thread block per SM, and kept only
// 18 KB of loads in flight per SM (89 us for 231 MB, the first and fully exposed kernel of the backward pass). Now two
// blocks per SM with four pixels in flight per thread.
template <typename T, int CIN, int NC>
__global__ void __launch_bounds__(256, (NC <= 5 ? 2 : 1)) head_bwd_kernel(const float* __restrict__ dlogits, const T* __restrict__ a,
                                                          const float* __restrict__ w, T* __restrict__ da,
                                                          float* __restrict__ dw, float* __restrict__ db, int P, int HW,
                                                          int C) {
  constexpr int VECS = CIN / 8, PPB = 256 / VECS;
  __shared__ float sacc[NC * CIN + NC];
  __shared__ __align__(16) float s_w[NC * CIN];
  for (int i = threadIdx.x; i < NC * CIN + NC; i += blockDim.x) sacc[i] = 0.f;
  for (int i = threadIdx.x; i < NC * CIN; i += blockDim.x) s_w[i] = i < C * CIN ? w[i] : 0.f;
  __syncthreads();
  const int v = threadIdx.x % VECS;
  float accw[NC][8], accb[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    accb[c] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) accw[c][j] = 0.f;
  }
  constexpr int U = 3;
  for (int p0 = blockIdx.x * PPB + threadIdx.x / VECS; p0 < P; p0 += gridDim.x * PPB * U) {
    Vec8<T> pk[U];
    float dl[U][NC];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int p = p0 + u * gridDim.x * PPB;
      if (p < P) {
        pk[u].load(a + static_cast<size_t>(p) * CIN + v * 8);
        const int n = p / HW, hw = p % HW;
#pragma unroll
        for (int c = 0; c < NC; ++c) dl[u][c] = c < C ? __ldg(dlogits + (static_cast<size_t>(n) * C + c) * HW + hw) : 0.f;
      } else {
        pk[u].zero();
#pragma unroll
        for (int c = 0; c < NC; ++c) dl[u][c] = 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int p = p0 + u * gridDim.x * PPB;
      float f[8], o[8];
      pk[u].get(f);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = 0.f;
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        accb[c] += dl[u][c];
        const float4 w0 = *reinterpret_cast<const float4*>(&s_w[c * CIN + v * 8]);
        const float4 w1 = *reinterpret_cast<const float4*>(&s_w[c * CIN + v * 8 + 4]);
        const float wr[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          accw[c][j] = fmaf(dl[u][c], f[j], accw[c][j]);
          o[j] = fmaf(dl[u][c], wr[j], o[j]);
        }
      }
      if (da != nullptr && p < P) {
        Vec8<T> q;
        q.set(o);
        q.store(da + static_cast<size_t>(p) * CIN + v * 8);
      }
    }
  }
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float t = accw[c][j];
#pragma unroll
      for (int o = VECS; o < 32; o <<= 1) t += __shfl_xor_sync(0xffffffffu, t, o);   // lanes with the same v
      if (lane < VECS && c < C) atomicAdd(&sacc[c * CIN + v * 8 + j], t);
    }
    float t = (v == 0) ? accb[c] : 0.f;   // every lane of a pixel saw the same dl: count it once
    t = warp_sum(t);
    if (lane == 0 && c < C) atomicAdd(&sacc[NC * CIN + c], t);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * CIN; i += blockDim.x) atomicAdd(dw + i, sacc[i]);
  if (db != nullptr)
    for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(db + i, sacc[NC * CIN + i]);
}

#define PP_HEAD_DISPATCH(MACRO)                                                  \
  do {                                                                           \
    if (C <= 2)      { if (Cin == 32) MACRO(32, 2); else if (Cin == 64) MACRO(64, 2); else MACRO(128, 2); } \
    else if (C <= 4) { if (Cin == 32) MACRO(32, 4); else if (Cin == 64) MACRO(64, 4); else MACRO(128, 4); } \
    else if (C == 5) { if (Cin == 32) MACRO(32, 5); else if (Cin == 64) MACRO(64, 5); else MACRO(128, 5); } \
    else             { if (Cin == 32) MACRO(32, 8); else if (Cin == 64) MACRO(64, 8); else MACRO(128, 8); } \
  } while (0)

int head_fwd(int dtype, const void* a, const float* w, const float* bias, float* logits, long long P, int HW, int Cin,
             int C, cudaStream_t s) {
  PP_REQUIRE(C >= 1 && C <= kMaxClasses, "head_fwd: num_classes=%d unsupported (max %d)", C, kMaxClasses);
  PP_REQUIRE(Cin == 32 || Cin == 64 || Cin == 128, "head_fwd: Cin=%d unsupported (32/64/128)", Cin);
  PP_REQUIRE_INT32(P * Cin, "head_fwd");
  const int grid = grid_for(ceil_div_ll(P, 4 * (256 / (Cin / 8))) * 256, 256, 8);
#define PP_HEAD_FWD(CIN_, NC_) \
  head_fwd_kernel<T, CIN_, NC_><<<grid, 256, 0, s>>>(static_cast<const T*>(a), w, bias, logits, int(P), HW, C)
  PP_DISPATCH_T(dtype, PP_HEAD_DISPATCH(PP_HEAD_FWD););
#undef PP_HEAD_FWD
  PP_LAUNCH_CHECK();
  return PP_OK;
}

int head_bwd(int dtype, const float* dlogits, const void* a, const float* w, void* da, float* dw, float* db, long long P,
             int HW, int Cin, int C, cudaStream_t s) {
  PP_REQUIRE(C >= 1 && C <= kMaxClasses, "head_bwd: num_classes=%d unsupported", C);
  PP_REQUIRE(Cin == 32 || Cin == 64 || Cin == 128, "head_bwd: Cin=%d unsupported (32/64/128)", Cin);
  PP_REQUIRE_INT32(P * Cin, "head_bwd");
  int grid = sm_count() * 2;
  const long long need = ceil_div_ll(P, 3 * (256 / (Cin / 8)));
  if (grid > need) grid = static_cast<int>(need < 1 ? 1 : need);
#define PP_HEAD_BWD(CIN_, NC_)                                                                                     \
  head_bwd_kernel<T, CIN_, NC_><<<grid, 256, 0, s>>>(dlogits, static_cast<const T*>(a), w, static_cast<T*>(da), dw, db, \
                                                     int(P), HW, C)
  PP_DISPATCH_T(dtype, PP_HEAD_DISPATCH(PP_HEAD_BWD););
#undef PP_HEAD_BWD
  PP_LAUNCH_CHECK();
  return PP_OK;
}

// ==============================================================================================
// BatchNorm2d (eps 1e-5, momentum 0.1) + LeakyReLU(0.01) on the conv output y [G*Pg][C].
// Statistics are per "group" g (the weak and the strong branch are batched into one tensor but
// keep separate batch statistics, exactly as two reference forward passes would).
// ==============================================================================================
// sums[g][c][0..1] (double) += sum y, sum y^2
template <typename T>
__global__ void __launch_bounds__(256) bn_stats_kernel(const T* __restrict__ y, double* __restrict__ sums, long long Pg,
                                                       int C, long long chunk, int chunks_per_group) {
  __shared__ float red[256][17];
  const int vecs = C / 8;
  const int pl = 256 / vecs;                 // pixel lanes per block
  const int v = threadIdx.x % vecs, l = threadIdx.x / vecs;
  const int g = blockIdx.x / chunks_per_group;
  const long long p0 = static_cast<long long>(blockIdx.x % chunks_per_group) * chunk;
  const long long p1 = (p0 + chunk < Pg) ? p0 + chunk : Pg;
  float s[8], ss[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s[j] = 0.f; ss[j] = 0.f; }
  if (l < pl) {
    constexpr int U = 4;
    const T* base = y + (static_cast<long long>(g) * Pg) * C + v * 8;
    for (long long pb = p0 + l; pb < p1; pb += static_cast<long long>(pl) * U) {
      Vec8<T> pk[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long p = pb + static_cast<long long>(u) * pl;
        if (p < p1) pk[u].load(base + p * C);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (pb + static_cast<long long>(u) * pl >= p1) break;
        float f[8];
        pk[u].get(f);
#pragma unroll
        for (int j = 0; j < 8; ++j) { s[j] += f[j]; ss[j] = fmaf(f[j], f[j], ss[j]); }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) { red[threadIdx.x][j] = s[j]; red[threadIdx.x][8 + j] = ss[j]; }
  __syncthreads();
  // thread t < C*2 reduces channel (t % C), stat (t / C) over the pixel lanes
  for (int t = threadIdx.x; t < 2 * C; t += 256) {
    const int c = t % C, st = t / C;
    double acc = 0.0;
    for (int k = 0; k < pl; ++k) acc += static_cast<double>(red[k * vecs + c / 8][st * 8 + c % 8]);
    atomicAdd(sums + (static_cast<long long>(g) * C + c) * 2 + st, acc);
  }
}

int bn_stats(int dtype, const void* y, double* sums, int G, long long Pg, int C, cudaStream_t s) {
  PP_REQUIRE(C % 8 == 0 && C <= 2048 && (256 % (C / 8) == 0 || C / 8 > 256), "bn_stats: C=%d unsupported", C);
  PP_REQUIRE(C / 8 <= 256, "bn_stats: C=%d too large", C);
  int cpg = (sm_count() * 4) / G;
  if (cpg < 1) cpg = 1;
  long long chunk = ceil_div_ll(Pg, cpg);
  if (chunk < 256) chunk = 256;
  cpg = static_cast<int>(ceil_div_ll(Pg, chunk));
  PP_DISPATCH_T(dtype, bn_stats_kernel<T><<<G * cpg, 256, 0, s>>>(static_cast<const T*>(y), sums, Pg, C, chunk, cpg););
  PP_LAUNCH_CHECK();
  return PP_OK;
}

// Finalize: per-group scale/shift/mean/rstd; running statistics updated group by group (train).
// coef layout: [G][4][C] floats = scale, shift, mean, rstd.
__global__ void __launch_bounds__(128) bn_finalize_kernel(const double* __restrict__ sums,
                                                          const float* __restrict__ gamma,
                                                          const float* __restrict__ beta,
                                                          float* __restrict__ running_mean,
                                                          float* __restrict__ running_var,
                                                          long long* __restrict__ num_batches_tracked,
                                                          float* __restrict__ coef, int G, long long Pg, int C,
                                                          int training, float eps, float momentum, int replicas) {
  // one warp per channel; the lanes fetch the replicated partial sums in parallel
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (c >= C) return;
  float rm = running_mean[c], rv = running_var[c];
  for (int g = 0; g < G; ++g) {
    float mean, rstd;
    if (training) {
      double s1 = 0.0, s2 = 0.0;
      for (int r = lane; r < replicas; r += 32) {
        const double* sp = sums + ((static_cast<long long>(r) * G + g) * C + c) * 2;
        s1 += sp[0];
        s2 += sp[1];
      }
      s1 = warp_sum_d(s1);
      s2 = warp_sum_d(s2);
      const double m = s1 / static_cast<double>(Pg);
      double var = s2 / static_cast<double>(Pg) - m * m;
      if (var < 0.0) var = 0.0;
      mean = static_cast<float>(m);
      rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
      const double unbiased = Pg > 1 ? var * static_cast<double>(Pg) / static_cast<double>(Pg - 1) : var;
      rm = (1.f - momentum) * rm + momentum * mean;
      rv = (1.f - momentum) * rv + momentum * static_cast<float>(unbiased);
    } else {
      mean = rm;
      rstd = 1.f / sqrtf(rv + eps);
    }
    if (lane == 0) {
      const float sc = gamma[c] * rstd;
      float* cf = coef + static_cast<long long>(g) * 4 * C;
      cf[c] = sc;
      cf[C + c] = beta[c] - mean * sc;
      cf[2 * C + c] = mean;
      cf[3 * C + c] = rstd;
    }
  }
  if (training && lane == 0) {
    running_mean[c] = rm;
    running_var[c] = rv;
    if (c == 0 && num_batches_tracked != nullptr) *num_batches_tracked += G;
  }
}

int bn_finalize(const double* sums, const float* gamma, const float* beta, float* running_mean, float* running_var,
                long long* nbt, float* coef, int G, long long Pg, int C, int training, float eps, float momentum,
                cudaStream_t s, int replicas) {
  bn_finalize_kernel<<<ceil_div(C, 4), 128, 0, s>>>(sums, gamma, beta, running_mean, running_var, nbt, coef, G, Pg, C,
                                                    training, eps, momentum, replicas);
  PP_LAUNCH_CHECK();
  return PP_OK;
}

// a = lrelu(y * scale + shift). Blocks own (group, pixel chunk); a thread keeps one 8-channel vector's
// coefficients in registers and streams pixels with four independent 16-byte loads in flight.
// Wide layers are split into channel SLABS of at most kBnSlabVecs 8-channel vectors (grid.y): with C = 512 one block row
// of 64 vectors would leave only 2 pixel lanes per 128-thread block and a few hundred threads per SM — these kernels sit
// on the critical chain between the convolutions, so their latency at the 32 x 32 stages matters as much as bandwidth.
static constexpr int kBnSlabVecs = 16;
static inline int bn_slab_vecs(int C) { return C / 8 < kBnSlabVecs ? C / 8 : kBnSlabVecs; }
// Replicas of the one-pass eval backward's accumulators (bn_bwd_eval): its ~890 blocks all end in 2 * (slab channels)
// double atomics, and same-cache-line atomics serialise in L2. The atomics per line scale with 1 / (channel slabs), so the
// narrow layers get the most replicas: 8 for C <= 128, 4 at 256, 2 at 512. PP_BN_REPLICAS=n forces a count (1 = none).
int bn_bwd_replicas(int C) {
  static const int forced = [] { const char* e = getenv("PP_BN_REPLICAS"); return e ? atoi(e) : 0; }();
  if (forced > 0) return forced < kBnBwdReplicas ? forced : kBnBwdReplicas;
  if (C % 8 != 0) return 1;
  const int slabs = (C / 8) / bn_slab_vecs(C);
  const int r = kBnBwdReplicas / (slabs > 0 ? slabs : 1);
  return r < 1 ? 1 : r;
}

template <typename T>
__global__ void __launch_bounds__(kBnThreads) bn_apply_kernel(const T* __restrict__ y, const float* __restrict__ coef,
                                                       T* __restrict__ a, int Pg, int C, int chunk,
                                                       int chunks_per_group, float slope, int vecs) {
  constexpr int U = 4;
  const int pl = kBnThreads / vecs;
  const int v = threadIdx.x % vecs, l = threadIdx.x / vecs;
  const int c0 = blockIdx.y * vecs * 8 + v * 8;   // first channel of this thread
  const int g = blockIdx.x / chunks_per_group;
  const int p0 = (blockIdx.x % chunks_per_group) * chunk;
  const int p1 = min(p0 + chunk, Pg);
  if (l >= pl) return;
  float sc[8], sh[8];
  const float* cf = coef + static_cast<size_t>(g) * 4 * C + c0;
#pragma unroll
  for (int j = 0; j < 8; ++j) { sc[j] = cf[j]; sh[j] = cf[C + j]; }
  const size_t base = static_cast<size_t>(g) * Pg * C + c0;
  for (int pb = p0 + l; pb < p1; pb += pl * U) {
    Vec8<T> pk[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (pb + u * pl < p1) pk[u].load(y + base + static_cast<size_t>(pb + u * pl) * C);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (pb + u * pl >= p1) break;
      float f[8];
      pk[u].get(f);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = lrelu(fmaf(f[j], sc[j], sh[j]), slope);
      pk[u].set(f);
      pk[u].store(a + base + static_cast<size_t>(pb + u * pl) * C);
    }
  }
}

static void bn_chunks(int G, long long Pg, int blocks_per_sm, int* chunk, int* cpg, int slabs = 1) {
  int c = (sm_count() * blocks_per_sm) / (G * slabs);
  if (c < 1) c = 1;
  long long ch = ceil_div_ll(Pg, c);
  if (ch < 64) ch = 64;
  *chunk = static_cast<int>(ch);
  *cpg = static_cast<int>(ceil_div_ll(Pg, ch));
}

int bn_apply(int dtype, const void* y, const float* coef, void* a, int G, long long Pg, int C, float slope,
             cudaStream_t s) {
  PP_REQUIRE(C % 8 == 0 && C / 8 <= kBnThreads && kBnThreads % (C / 8) == 0, "bn_apply: C=%d unsupported", C);
  PP_REQUIRE_INT32(Pg * C, "bn_apply");
  int chunk, cpg;
  const int vecs = bn_slab_vecs(C), slabs = (C / 8) / vecs;
  PP_REQUIRE((C / 8) % vecs == 0, "bn_apply: C=%d does not split into channel slabs", C);
  bn_chunks(G, Pg, 16, &chunk, &cpg, slabs);
  PP_DISPATCH_T(dtype, bn_apply_kernel<T><<<dim3(G * cpg, slabs), kBnThreads, 0, s>>>(
                           static_cast<const T*>(y), coef, static_cast<T*>(a), int(Pg), C, chunk, cpg, slope, vecs););
  PP_LAUNCH_CHECK();
  return PP_OK;
}

// Backward reduce: bsums[g][c][0..1] += sum dz, sum dz * xhat   with  z = y*scale+shift,
// dz = da * (z > 0 ? 1 : slope),  xhat = (y - mean) * rstd.
template <typename T>
__global__ void __launch_bounds__(kBnThreads) bn_bwd_reduce_kernel(const T* __restrict__ da, const T* __restrict__ y,
                                                            const float* __restrict__ coef, double* __restrict__ bsums,
                                                            long long Pg, int C, long long chunk, int chunks_per_group,
                                                            float slope, int vecs) {
  __shared__ float red[kBnThreads][17];
  const int pl = kBnThreads / vecs;
  const int v = threadIdx.x % vecs, l = threadIdx.x / vecs;
  const int cs = blockIdx.y * vecs * 8;          // first channel of this block's slab
  const int c0 = cs + v * 8;
  const int g = blockIdx.x / chunks_per_group;
  const long long p0 = static_cast<long long>(blockIdx.x % chunks_per_group) * chunk;
  const long long p1 = (p0 + chunk < Pg) ? p0 + chunk : Pg;
  float s[8], ss[8], sc[8], sh[8], mu[8], rs[8];
  const float* cf = coef + static_cast<long long>(g) * 4 * C + c0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    s[j] = 0.f; ss[j] = 0.f;
    sc[j] = cf[j]; sh[j] = cf[C + j]; mu[j] = cf[2 * C + j]; rs[j] = cf[3 * C + j];
  }
  if (l < pl) {
    constexpr int U = 8;
    const long long base = (static_cast<long long>(g) * Pg) * C + c0;
    for (long long pb = p0 + l; pb < p1; pb += static_cast<long long>(pl) * U) {
      Vec8<T> pa[U], py[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long p = pb + static_cast<long long>(u) * pl;
        if (p < p1) { pa[u].load(da + base + p * C); py[u].load(y + base + p * C); }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long p = pb + static_cast<long long>(u) * pl;
        if (p >= p1) break;
        float fa[8], fy[8];
        pa[u].get(fa);
        py[u].get(fy);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float z = fmaf(fy[j], sc[j], sh[j]);
          const float dz = z > 0.f ? fa[j] : fa[j] * slope;
          s[j] += dz;
          ss[j] = fmaf(dz, (fy[j] - mu[j]) * rs[j], ss[j]);
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) { red[threadIdx.x][j] = s[j]; red[threadIdx.x][8 + j] = ss[j]; }
  __syncthreads();
  const int cb = vecs * 8;                       // channels of this slab
  for (int t = threadIdx.x; t < 2 * cb; t += kBnThreads) {
    const int c = t % cb, st = t / cb;
    double acc = 0.0;
    for (int k = 0; k < pl; ++k) acc += static_cast<double>(red[k * vecs + c / 8][st * 8 + c % 8]);
    atomicAdd(bsums + (static_cast<long long>(g) * C + cs + c) * 2 + st, acc);
  }
}

// dy = scale * (dz - k1 - xhat * k2) with k1 = sum dz / Pg, k2 = sum dz*xhat / Pg (zeros in eval mode); same
// (group, chunk) decomposition, coefficients in registers. The former finalize kernel is folded in: every thread
// derives k1/k2 of its 8 channels from the reduced sums, and block 0 also adds the parameter gradients
// (dgamma += sum_g sum dz*xhat, dbeta += sum_g sum dz, eval mode: dbias += sum_g scale * sum dz).
template <typename T>
__global__ void __launch_bounds__(kBnThreads) bn_bwd_apply_kernel(const T* __restrict__ da, const T* __restrict__ y,
                                                           const float* __restrict__ coef,
                                                           const double* __restrict__ bsums, T* __restrict__ dy,
                                                           float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                           float* __restrict__ dbias, int G, int Pg, int C, int chunk,
                                                           int chunks_per_group, int training, float slope, int vecs) {
  constexpr int U = 4;
  const int pl = kBnThreads / vecs;
  const int v = threadIdx.x % vecs, l = threadIdx.x / vecs;
  const int c0 = blockIdx.y * vecs * 8 + v * 8;   // first channel of this thread (channel slab = blockIdx.y)
  const int g = blockIdx.x / chunks_per_group;
  const int p0 = (blockIdx.x % chunks_per_group) * chunk;
  const int p1 = min(p0 + chunk, Pg);
  if (l >= pl) return;
  if (blockIdx.x == 0 && l == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = c0 + j;
      double dg = 0.0, dbt = 0.0, dbs = 0.0;
      for (int gg = 0; gg < G; ++gg) {
        const double sdz = bsums[(static_cast<size_t>(gg) * C + c) * 2];
        dbt += sdz;
        dg += bsums[(static_cast<size_t>(gg) * C + c) * 2 + 1];
        if (!training) dbs += sdz * static_cast<double>(coef[static_cast<size_t>(gg) * 4 * C + c]);  // scale * sum dz
      }
      dgamma[c] += static_cast<float>(dg);
      dbeta[c] += static_cast<float>(dbt);
      if (dbias != nullptr) dbias[c] += static_cast<float>(dbs);  // exactly 0 under batch statistics
    }
  }
  float sc[8], sh[8], mu[8], rs[8], k1[8], k2[8];
  const float* cf = coef + static_cast<size_t>(g) * 4 * C + c0;
  const double* bs = bsums + (static_cast<size_t>(g) * C + c0) * 2;
  const double inv = 1.0 / static_cast<double>(Pg);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = cf[j]; sh[j] = cf[C + j]; mu[j] = cf[2 * C + j]; rs[j] = cf[3 * C + j];
    k1[j] = training ? static_cast<float>(bs[2 * j] * inv) : 0.f;
    k2[j] = training ? static_cast<float>(bs[2 * j + 1] * inv) : 0.f;
  }
  const size_t base = static_cast<size_t>(g) * Pg * C + c0;
  for (int pb = p0 + l; pb < p1; pb += pl * U) {
    Vec8<T> pa[U], py[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (pb + u * pl < p1) {
        const size_t o = base + static_cast<size_t>(pb + u * pl) * C;
        pa[u].load(da + o);
        py[u].load(y + o);
      }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (pb + u * pl >= p1) break;
      float fa[8], fy[8], o[8];
      pa[u].get(fa);
      py[u].get(fy);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float z = fmaf(fy[j], sc[j], sh[j]);
        const float dz = z > 0.f ? fa[j] : fa[j] * slope;
        const float xhat = (fy[j] - mu[j]) * rs[j];
        o[j] = sc[j] * (dz - k1[j] - xhat * k2[j]);
      }
      pa[u].set(o);
      pa[u].store(dy + base + static_cast<size_t>(pb + u * pl) * C);
    }
  }
}

int bn_bwd(int dtype, const void* da, const void* y, const float* coef, double* bsums, float* bcoef, float* dgamma,
           float* dbeta, float* dbias, void* dy, int G, long long Pg, int C, int training, float slope,
           cudaStream_t s, bool sums_zeroed) {
  (void)bcoef;   // kept in the signature (workspace layout); the apply pass reads the reduced sums directly
  PP_REQUIRE(C % 8 == 0 && C / 8 <= kBnThreads && kBnThreads % (C / 8) == 0, "bn_bwd: C=%d unsupported", C);
  PP_REQUIRE_INT32(Pg * C, "bn_bwd");
  if (!sums_zeroed) PP_CHECK_CUDA(cudaMemsetAsync(bsums, 0, sizeof(double) * 2 * G * C, s));
  static const int red_bps = [] { const char* e = getenv("PP_BN_RED_BPS"); return e ? atoi(e) : 4; }();
  const int vecs = bn_slab_vecs(C), slabs = (C / 8) / vecs;
  PP_REQUIRE((C / 8) % vecs == 0, "bn_bwd: C=%d does not split into channel slabs", C);
  int cpg = (sm_count() * (red_bps > 0 ? red_bps : 4)) / (G * slabs);
  if (cpg < 1) cpg = 1;
  long long chunk = ceil_div_ll(Pg, cpg);
  if (chunk < 64) chunk = 64;
  cpg = static_cast<int>(ceil_div_ll(Pg, chunk));
  int achunk, acpg;
  bn_chunks(G, Pg, 16, &achunk, &acpg, slabs);
  PP_DISPATCH_T(dtype,
                bn_bwd_reduce_kernel<T><<<dim3(G * cpg, slabs), kBnThreads, 0, s>>>(
                    static_cast<const T*>(da), static_cast<const T*>(y), coef, bsums, Pg, C, chunk, cpg, slope, vecs);
                bn_bwd_apply_kernel<T><<<dim3(G * acpg, slabs), kBnThreads, 0, s>>>(
                    static_cast<const T*>(da), static_cast<const T*>(y), coef, bsums, static_cast<T*>(dy), dgamma, dbeta,
                    dbias, G, int(Pg), C, achunk, acpg, training, slope, vecs););
  PP_LAUNCH_CHECK_N(2);
  return PP_OK;
}

// ----------------------------------------------------------------------------------------------
// Eval-mode BatchNorm folded into the convolutions (bf16 path; the reference's steady state: train_chaos.py:370 puts the
// model in .eval() for validation and never leaves it). Forward: the conv epilogues apply
// lrelu(acc * scale + shift) and write the activation directly (no pre-BN tensor, no finalize / apply pass).
// ----------------------------------------------------------------------------------------------
struct EvalCoefTable {
  static constexpr int kMax = 48;
  const float* gamma[kMax];
  const float* beta[kMax];
  const float* rmean[kMax];
  const float* rvar[kMax];
  const float* bias[kMax];
  float* coef[kMax];
  int C[kMax];
  int n;
};
// coef = [scale | shift | beta | 1/gamma], C floats each; grid.x = layer
__global__ void __launch_bounds__(256) bn_eval_coef_multi_kernel(const __grid_constant__ EvalCoefTable tb, float eps) {
  const int l = blockIdx.x;
  const int C = tb.C[l];
  float* cf = tb.coef[l];
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float g = tb.gamma[l][c], b = tb.beta[l][c];
    const float sc = g * (1.f / sqrtf(tb.rvar[l][c] + eps));
    const float bias = tb.bias[l] != nullptr ? tb.bias[l][c] : 0.f;
    cf[c] = sc;
    cf[C + c] = b + (bias - tb.rmean[l][c]) * sc;
    cf[2 * C + c] = b;
    cf[3 * C + c] = g != 0.f ? 1.f / g : 0.f;   // gamma == 0: xhat is not recoverable from the activation (dgamma term 0)
  }
}
int bn_eval_coef_multi(int n, const float* const* gamma, const float* const* beta, const float* const* rmean,
                       const float* const* rvar, const float* const* bias, float* const* coef, const int* C, float eps,
                       cudaStream_t s) {
  for (int i0 = 0; i0 < n; i0 += EvalCoefTable::kMax) {
    EvalCoefTable tb{};
    tb.n = n - i0 < EvalCoefTable::kMax ? n - i0 : EvalCoefTable::kMax;
    for (int i = 0; i < tb.n; ++i) {
      tb.gamma[i] = gamma[i0 + i]; tb.beta[i] = beta[i0 + i]; tb.rmean[i] = rmean[i0 + i]; tb.rvar[i] = rvar[i0 + i];
      tb.bias[i] = bias[i0 + i]; tb.coef[i] = coef[i0 + i]; tb.C[i] = C[i0 + i];
    }
    bn_eval_coef_multi_kernel<<<tb.n, 256, 0, s>>>(tb, eps);
    PP_LAUNCH_CHECK();
  }
  return PP_OK;
}

// Backward of conv -> eval BatchNorm -> LeakyReLU from the saved ACTIVATION: one pass.
//   dz = da * (a > 0 ? 1 : slope);  dy = dz * scale;  z = a > 0 ? a : a / slope;  xhat = (z - beta) / gamma
//   sums[c] += (sum dz, sum dz * xhat); the last block to finish adds dgamma / dbeta / dbias (ticket counter).
template <typename T>
__global__ void __launch_bounds__(kBnThreads) bn_bwd_eval_kernel(const T* __restrict__ da, const T* __restrict__ a,
                                                                 const float* __restrict__ coef, double* sums,
                                                                 T* __restrict__ dy, float* __restrict__ dgamma,
                                                                 float* __restrict__ dbeta, float* __restrict__ dbias,
                                                                 int P, int C, int chunk, float slope, float inv_slope,
                                                                 int vecs, int reps) {
  __shared__ float red[kBnThreads][17];
  __shared__ unsigned int s_last;
  constexpr int U = 4;
  const int pl = kBnThreads / vecs;
  const int v = threadIdx.x % vecs, l = threadIdx.x / vecs;
  const int cs = blockIdx.y * vecs * 8;          // channel slab of this block
  const int p0 = blockIdx.x * chunk;
  const int p1 = min(p0 + chunk, P);
  float s[8], ss[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s[j] = 0.f; ss[j] = 0.f; }
  if (l < pl) {
    float sc[8], bt[8], ig[8];
    const float* cf = coef + cs + v * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) { sc[j] = cf[j]; bt[j] = cf[2 * C + j]; ig[j] = cf[3 * C + j]; }
    const size_t base = static_cast<size_t>(cs) + static_cast<size_t>(v) * 8;
    for (int pb = p0 + l; pb < p1; pb += pl * U) {
      Vec8<T> pa[U], py[U];
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (pb + u * pl < p1) {
          const size_t o = base + static_cast<size_t>(pb + u * pl) * C;
          pa[u].load(da + o);
          py[u].load(a + o);
        }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (pb + u * pl >= p1) break;
        float fa[8], fy[8], o[8];
        pa[u].get(fa);
        py[u].get(fy);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const bool pos = fy[j] > 0.f;
          const float dz = pos ? fa[j] : fa[j] * slope;
          const float z = pos ? fy[j] : fy[j] * inv_slope;
          s[j] += dz;
          ss[j] = fmaf(dz, (z - bt[j]) * ig[j], ss[j]);
          o[j] = dz * sc[j];
        }
        pa[u].set(o);
        pa[u].store(dy + base + static_cast<size_t>(pb + u * pl) * C);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) { red[threadIdx.x][j] = s[j]; red[threadIdx.x][8 + j] = ss[j]; }
  __syncthreads();
  const int cb = vecs * 8;
  for (int t = threadIdx.x; t < 2 * cb; t += kBnThreads) {
    const int c = t % cb, st = t / cb;
    double acc = 0.0;
    for (int k = 0; k < pl; ++k) acc += static_cast<double>(red[k * vecs + c / 8][st * 8 + c % 8]);
    // [replica][C][2]: the blocks go round-robin over the replicas (see bn_bwd_replicas)
    atomicAdd(sums + (static_cast<size_t>((blockIdx.x + blockIdx.y) % reps) * C + cs + c) * 2 + st, acc);
  }
  // last block (of all slabs): parameter gradients from the complete sums
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int* ticket = reinterpret_cast<unsigned int*>(sums + 2 * static_cast<size_t>(C) * reps);
    s_last = (atomicAdd(ticket, 1u) == gridDim.x * gridDim.y - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    for (int c = threadIdx.x; c < C; c += kBnThreads) {
      double2 part[kBnBwdReplicas];   // all replica loads in flight before the first add
#pragma unroll
      for (int r = 0; r < kBnBwdReplicas; ++r)
        part[r] = r < reps ? __ldcg(reinterpret_cast<const double2*>(sums + (static_cast<size_t>(r) * C + c) * 2))
                           : make_double2(0.0, 0.0);
      double sdz = 0.0, sdx = 0.0;
#pragma unroll
      for (int r = 0; r < kBnBwdReplicas; ++r) { sdz += part[r].x; sdx += part[r].y; }
      dbeta[c] += static_cast<float>(sdz);
      dgamma[c] += static_cast<float>(sdx);
      if (dbias != nullptr) dbias[c] += static_cast<float>(sdz * static_cast<double>(coef[c]));
    }
  }
}

int bn_bwd_eval(int dtype, const void* da, const void* a, const float* coef, double* sums, float* dgamma, float* dbeta,
                float* dbias, void* dy, long long P, int C, float slope, cudaStream_t s, int reps) {
  PP_REQUIRE(C % 8 == 0 && C / 8 <= kBnThreads && kBnThreads % (C / 8) == 0, "bn_bwd_eval: C=%d unsupported", C);
  if (reps < 1) reps = 1;
  if (reps > kBnBwdReplicas) reps = kBnBwdReplicas;
  PP_REQUIRE(slope != 0.f, "bn_bwd_eval: a zero slope is not invertible");
  PP_REQUIRE_INT32(P * C, "bn_bwd_eval");
  int chunk, blocks;
  const int vecs = bn_slab_vecs(C), slabs = (C / 8) / vecs;
  PP_REQUIRE((C / 8) % vecs == 0, "bn_bwd_eval: C=%d does not split into channel slabs", C);
  // fewer, fatter blocks than the pure streaming kernels: every block ends with 2 * (slab channels) double atomics on a
  // handful of cache lines, and ~2400 blocks of them cost more than the pass itself (88 us instead of ~30 us at 128^2)
  static const int bps = [] { const char* e = getenv("PP_BN_EVAL_BPS"); return e ? atoi(e) : 6; }();
  bn_chunks(1, P, bps > 0 ? bps : 6, &chunk, &blocks, slabs);
  PP_DISPATCH_T(dtype, bn_bwd_eval_kernel<T><<<dim3(blocks, slabs), kBnThreads, 0, s>>>(
                           static_cast<const T*>(da), static_cast<const T*>(a), coef, sums, static_cast<T*>(dy), dgamma,
                           dbeta, dbias, int(P), C, chunk, slope, 1.f / slope, vecs, reps););
  PP_LAUNCH_CHECK();
  return PP_OK;
}

// ==============================================================================================
// MaxPool2d(2, 2) (unet.py:109). Backward recomputes the arg-max from the saved input (first
// maximum in window scan order, as ATen does) and ACCUMULATES into dx.
// ==============================================================================================
template <typename T>
__global__ void __launch_bounds__(256) maxpool_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H,
                                                          int W, int C) {
  const int vecs = C / 8, Ho = H / 2, Wo = W / 2;
  const int total = N * Ho * Wo * vecs;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int v = i % vecs;
    const int p = i / vecs;
    const int xo = p % Wo, yo = (p / Wo) % Ho;
    const size_t n = p / (Wo * Ho);
    float m[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      Vec8<T> pk;
      pk.load(x + ((n * H + yo * 2 + k / 2) * W + xo * 2 + k % 2) * C + v * 8);
      float f[8];
      pk.get(f);
#pragma unroll
      for (int j = 0; j < 8; ++j) m[j] = (k == 0 || f[j] > m[j]) ? f[j] : m[j];
    }
    Vec8<T> o;
    o.set(m);
    o.store(y + static_cast<size_t>(i) * 8);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const T* __restrict__ x, const T* __restrict__ gy,
                                                          T* __restrict__ gx, int N, int H, int W, int C,
                                                          int accumulate) {
  const int vecs = C / 8, Ho = H / 2, Wo = W / 2;
  const int total = N * Ho * Wo * vecs;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int v = i % vecs;
    const int p = i / vecs;
    const int xo = p % Wo, yo = (p / Wo) % Ho;
    const size_t n = p / (Wo * Ho);
    float m[8];
    int am[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      Vec8<T> pk;
      pk.load(x + ((n * H + yo * 2 + k / 2) * W + xo * 2 + k % 2) * C + v * 8);
      float f[8];
      pk.get(f);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (k == 0 || f[j] > m[j]) { m[j] = f[j]; am[j] = k; }
    }
    Vec8<T> pg;
    pg.load(gy + static_cast<size_t>(i) * 8);
    float g[8];
    pg.get(g);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      T* dst = gx + ((n * H + yo * 2 + k / 2) * W + xo * 2 + k % 2) * C + v * 8;
      float o[8];
      if (accumulate) {
        Vec8<T> old;
        old.load(dst);
        old.get(o);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] += (am[j] == k) ? g[j] : 0.f;
      Vec8<T> w;
      w.set(o);
      w.store(dst);
    }
  }
}

int maxpool_fwd(int dtype, const void* x, void* y, int N, int H, int W, int C, cudaStream_t s) {
  PP_REQUIRE(H % 2 == 0 && W % 2 == 0 && C % 8 == 0, "maxpool_fwd: H=%d W=%d must be even, C=%d multiple of 8", H, W, C);
  PP_REQUIRE_INT32(static_cast<long long>(N) * H * W * C, "maxpool_fwd");
  const long long total = static_cast<long long>(N) * (H / 2) * (W / 2) * (C / 8);
  PP_DISPATCH_T(dtype, maxpool_fwd_kernel<T><<<grid_for(total, 256), 256, 0, s>>>(static_cast<const T*>(x),
                                                                                   static_cast<T*>(y), N, H, W, C););
  PP_LAUNCH_CHECK();
  return PP_OK;
}
int maxpool_bwd(int dtype, const void* x, const void* gy, void* gx, int N, int H, int W, int C, int accumulate,
                cudaStream_t s) {
  PP_REQUIRE(H % 2 == 0 && W % 2 == 0 && C % 8 == 0, "maxpool_bwd: H=%d W=%d must be even, C=%d multiple of 8", H, W, C);
  PP_REQUIRE_INT32(static_cast<long long>(N) * H * W * C, "maxpool_bwd");
  const long long total = static_cast<long long>(N) * (H / 2) * (W / 2) * (C / 8);
  PP_DISPATCH_T(dtype, maxpool_bwd_kernel<T><<<grid_for(total, 256), 256, 0, s>>>(
                           static_cast<const T*>(x), static_cast<const T*>(gy), static_cast<T*>(gx), N, H, W, C,
                           accumulate););
  PP_LAUNCH_CHECK();
  return PP_OK;
}

// ==============================================================================================
// Bilinear upsampling, align_corners=True (unet.py:144 nn.Upsample; aux_path_memory.py:52
// F.interpolate). Source index arithmetic follows ATen: src = dst * (in-1)/(out-1) in fp32.
// ==============================================================================================
// (Lerp / lerp_src / ac_scale live in pp_common.cuh: the fused scribble loss samples the aux logits with them too.)
// weight with which output index `dst` reads input index `i`
__device__ __forceinline__ float lerp_weight(int dst, int i, int in_size, float scale) {
  const Lerp l = lerp_src(dst, in_size, scale);
  float w = 0.f;
  if (l.i0 == i) w += l.w0;
  if (l.i1 == i) w += l.w1;
  return w;
}
// candidate output range that can read input index i
__device__ __forceinline__ void lerp_range(int i, int out_size, float scale, int* lo, int* hi) {
  if (scale <= 0.f) { *lo = 0; *hi = out_size - 1; return; }
  int a = static_cast<int>(floorf(static_cast<float>(i - 1) / scale)) - 1;
  int b = static_cast<int>(ceilf(static_cast<float>(i + 1) / scale)) + 1;
  *lo = a < 0 ? 0 : a;
  *hi = b > out_size - 1 ? out_size - 1 : b;
}

// Forward: one block row per output row (n, Y): the vertical taps are block-uniform, the thread loop walks
// (X, channel vector) with a single integer division per element.
template <typename T>
__global__ void __launch_bounds__(256) upsample_nhwc_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int N,
                                                                int h, int w, int H, int W, int C, float sh, float sw) {
  const int vecs = C / 8;
  const int rows = N * H;
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const int n = row / H, Y = row % H;
    const Lerp ly = lerp_src(Y, h, sh);
    const T* r0 = x + (static_cast<size_t>(n) * h + ly.i0) * w * C;
    const T* r1 = x + (static_cast<size_t>(n) * h + ly.i1) * w * C;
    T* out = y + static_cast<size_t>(row) * W * C;
    for (int t = threadIdx.x; t < W * vecs; t += blockDim.x) {
      const int X = t / vecs, v = t - X * vecs;
      const Lerp lx = lerp_src(X, w, sw);
      Vec8<T> p00, p01, p10, p11;
      p00.load(r0 + lx.i0 * C + v * 8);
      p01.load(r0 + lx.i1 * C + v * 8);
      p10.load(r1 + lx.i0 * C + v * 8);
      p11.load(r1 + lx.i1 * C + v * 8);
      float f00[8], f01[8], f10[8], f11[8], o[8];
      p00.get(f00); p01.get(f01); p10.get(f10); p11.get(f11);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        o[j] = ly.w0 * (lx.w0 * f00[j] + lx.w1 * f01[j]) + ly.w1 * (lx.w0 * f10[j] + lx.w1 * f11[j]);
      Vec8<T> q;
      q.set(o);
      q.store(out + static_cast<size_t>(t) * 8);
    }
  }
}

// Exact inverse of the forward index map. A(i) = first output index whose lower tap lerp_src(.).i0 is >= i
// (A(in_size) = out_size). Input i is read as the LOWER tap (weight w0) by outputs [A(i), A(i+1)) and as the UPPER
// tap (weight w1) by outputs [A(i-1), A(i)); at the last input index the upper tap folds onto the lower one
// (lerp_src: i1 == i0), where w1 == 0 up to rounding and is added all the same, exactly as the forward pass does.
__device__ __forceinline__ int lerp_first_out(int i, int in_size, int out_size, float scale) {
  if (i <= 0) return 0;
  if (i >= in_size || scale <= 0.f) return out_size;
  int a = static_cast<int>(static_cast<float>(i) / scale);
  a = a < 0 ? 0 : (a > out_size ? out_size : a);
  while (a > 0 && static_cast<int>(scale * static_cast<float>(a - 1)) >= i) --a;
  while (a < out_size && static_cast<int>(scale * static_cast<float>(a)) < i) ++a;
  return a;
}
constexpr int kUpWin = 12;   // widest candidate window handled from registers (scale factor <= ~5)

struct UpTaps { int lo, n; float w[kUpWin]; };
// all outputs that read input index i, with their weights (n == -1: window too wide, caller takes the slow path)
__device__ __forceinline__ UpTaps up_taps(int i, int in_size, int out_size, float scale) {
  UpTaps t;
  const int a0 = lerp_first_out(i - 1, in_size, out_size, scale);
  const int a1 = lerp_first_out(i, in_size, out_size, scale);
  const int a2 = lerp_first_out(i + 1, in_size, out_size, scale);
  t.lo = (i == 0) ? a1 : a0;
  t.n = a2 - t.lo;
  if (t.n > kUpWin) { t.n = -1; return t; }
#pragma unroll
  for (int k = 0; k < kUpWin; ++k) {
    const int o = t.lo + k;
    float wv = 0.f;
    if (k < t.n) {
      const Lerp l = lerp_src(o, in_size, scale);
      if (l.i0 == i) wv += l.w0;
      if (l.i1 == i) wv += l.w1;
    }
    t.w[k] = wv;
  }
  return t;
}

// gx[n,i,j,:] = sum over outputs (Y,X) reading (i,j) of wy*wx*gy[n,Y,X,:]  (gather, deterministic).
// One block row per INPUT row (n, yi): the vertical taps are block-uniform; the horizontal taps of every input
// column are computed once per block into shared memory (they do not depend on the row).
template <typename T>
__global__ void __launch_bounds__(256) upsample_nhwc_bwd_kernel(const T* __restrict__ gy, T* __restrict__ gx, int N,
                                                                int h, int w, int H, int W, int C, float sh, float sw,
                                                                int accumulate) {
  extern __shared__ float s_tx[];               // [w][kUpWin] weights, then [w] lo, [w] n (ints)
  int* s_lo = reinterpret_cast<int*>(s_tx + w * kUpWin);
  int* s_n = s_lo + w;
  for (int xi = threadIdx.x; xi < w; xi += blockDim.x) {
    const UpTaps tx = up_taps(xi, w, W, sw);
    s_lo[xi] = tx.lo;
    s_n[xi] = tx.n;
#pragma unroll
    for (int k = 0; k < kUpWin; ++k) s_tx[xi * kUpWin + k] = tx.w[k];
  }
  __syncthreads();
  const int vecs = C / 8;
  const int rows = N * h;
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const int n = row / h, yi = row % h;
    const UpTaps ty = up_taps(yi, h, H, sh);
    T* out = gx + static_cast<size_t>(row) * w * C;
    const T* gimg = gy + static_cast<size_t>(n) * H * W * C;
    for (int t = threadIdx.x; t < w * vecs; t += blockDim.x) {
      const int xi = t / vecs, v = t - xi * vecs;
      const int xlo0 = s_lo[xi], xn = s_n[xi];
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = 0.f;
      if (ty.n >= 0 && xn >= 0) {
        const float* wx = s_tx + xi * kUpWin;
#pragma unroll
        for (int a = 0; a < kUpWin; ++a) {
          if (a >= ty.n) break;
          const T* grow = gimg + (static_cast<size_t>(ty.lo + a) * W + xlo0) * C + v * 8;
          for (int b = 0; b < xn; ++b) {
            Vec8<T> pk;
            pk.load(grow + static_cast<size_t>(b) * C);
            float f[8];
            pk.get(f);
            const float ww = ty.w[a] * wx[b];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = fmaf(ww, f[j], acc[j]);
          }
        }
      } else {   // very large scale factors: scan the candidate window
        int ylo, yhi, xlo, xhi;
        lerp_range(yi, H, sh, &ylo, &yhi);
        lerp_range(xi, W, sw, &xlo, &xhi);
        for (int Y = ylo; Y <= yhi; ++Y) {
          const float wy = lerp_weight(Y, yi, h, sh);
          if (wy == 0.f) continue;
          for (int X = xlo; X <= xhi; ++X) {
            const float wxv = lerp_weight(X, xi, w, sw);
            if (wxv == 0.f) continue;
            Vec8<T> pk;
            pk.load(gimg + (static_cast<size_t>(Y) * W + X) * C + v * 8);
            float f[8];
            pk.get(f);
            const float ww = wy * wxv;
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = fmaf(ww, f[j], acc[j]);
          }
        }
      }
      Vec8<T> q;
      if (accumulate) {
        float old[8];
        q.load(out + static_cast<size_t>(t) * 8);
        q.get(old);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += old[j];
      }
      q.set(acc);
      q.store(out + static_cast<size_t>(t) * 8);
    }
  }
}

// Strip variant for the UNet's scale-2 layers: the gather above reads ~16 output-gradient vectors per input vector and
// is bound by L1/L2 traffic (168 us for the 250 MB of the 256^2 layer). Here a block owns P consecutive INPUT rows of
// one image and a thread owns TWO adjacent input columns x 8 channels; it streams the contributing output rows once,
// top to bottom. Per output row Y (block-uniform lerp_src(Y) = rows i0, i1 with weights w0, w1) it loads the union
// window of its two columns (<= kPairWin vectors, 6 at scale 2), folds it horizontally into hx0 / hx1 with the
// per-pair tap weights from shared memory, and adds w0 * hx to the accumulators of row i0 and w1 * hx to those of
// row i1, kept as a sliding (current, next) pair that is flushed when Y moves past a row. Loads per input vector:
// (2P + 2) * 6 / (2P) = 6.75 at P = 8 instead of 16; fixed summation order (deterministic); a few KB of shared
// memory and ~100 registers, so the blocks still co-reside with the conv CTAs of the weight-gradient stream.
constexpr int kPairWin = 8;
struct PairTaps { int lo, n; float wa[kPairWin], wb[kPairWin]; };
template <typename T>
__global__ void __launch_bounds__(256) upsample_nhwc_bwd_strip_kernel(const T* __restrict__ gy, T* __restrict__ gx,
                                                                      int h, int w, int H, int W, int C, float sh,
                                                                      float sw, int P, int strips, int accumulate) {
  extern __shared__ float s_raw[];
  PairTaps* s_tp = reinterpret_cast<PairTaps*>(s_raw);
  const int pairs = (w + 1) / 2;
  for (int j = threadIdx.x; j < pairs; j += blockDim.x) {
    const int xa = 2 * j, xb = 2 * j + 1;
    PairTaps t;
    t.lo = xa == 0 ? 0 : lerp_first_out(xa - 1, w, W, sw);
    const int hi = lerp_first_out(min(xb, w - 1) + 1, w, W, sw);
    t.n = hi - t.lo;
#pragma unroll
    for (int k = 0; k < kPairWin; ++k) {
      const int X = t.lo + k;
      t.wa[k] = (k < t.n) ? lerp_weight(X, xa, w, sw) : 0.f;
      t.wb[k] = (k < t.n && xb < w) ? lerp_weight(X, xb, w, sw) : 0.f;
    }
    s_tp[j] = t;
  }
  __syncthreads();
  const int n = blockIdx.x / strips, y0 = (blockIdx.x % strips) * P;
  const int y1 = min(y0 + P, h);
  const int Ybeg = y0 == 0 ? 0 : lerp_first_out(y0 - 1, h, H, sh), Yend = lerp_first_out(y1, h, H, sh);
  const int vecs = C / 8;
  const T* gimg = gy + static_cast<size_t>(n) * H * W * C;
  T* ximg = gx + static_cast<size_t>(n) * h * w * C;
  // grid.y walks the (column pair, channel vector) range in blockDim.x chunks: one pass over the strip per thread
  for (int t = blockIdx.y * blockDim.x + threadIdx.x; t < pairs * vecs; t += gridDim.y * blockDim.x) {
    const int j = t / vecs, v = t - j * vecs;
    const PairTaps& tp = s_tp[j];
    const int xa = 2 * j;
    const bool has_b = xa + 1 < w;
    float cur0[8], cur1[8], nxt0[8], nxt1[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) { cur0[c] = 0.f; cur1[c] = 0.f; nxt0[c] = 0.f; nxt1[c] = 0.f; }
    int cur_row = y0;
    auto flush = [&]() {   // row cur_row is complete: store it, slide the window down one row
      if (cur_row < y1) {
        T* o = ximg + (static_cast<size_t>(cur_row) * w + xa) * C + v * 8;
        Vec8<T> q;
        if (accumulate) {
          float old[8];
          q.load(o); q.get(old);
#pragma unroll
          for (int c = 0; c < 8; ++c) cur0[c] += old[c];
          if (has_b) {
            q.load(o + C); q.get(old);
#pragma unroll
            for (int c = 0; c < 8; ++c) cur1[c] += old[c];
          }
        }
        q.set(cur0); q.store(o);
        if (has_b) { q.set(cur1); q.store(o + C); }
      }
#pragma unroll
      for (int c = 0; c < 8; ++c) { cur0[c] = nxt0[c]; cur1[c] = nxt1[c]; nxt0[c] = 0.f; nxt1[c] = 0.f; }
      ++cur_row;
    };
    // software pipeline: the window of row Y + 1 is in flight while row Y is folded (the chain per thread is
    // ~2P dependent rounds of memory latency otherwise)
    Vec8<T> nx[kPairWin];
    auto load_row = [&](int Y) {
      const T* grow = gimg + (static_cast<size_t>(Y) * W + tp.lo) * C + v * 8;
#pragma unroll
      for (int k = 0; k < kPairWin; ++k)
        if (k < tp.n) nx[k].load(grow + static_cast<size_t>(k) * C);
    };
    if (Ybeg < Yend) load_row(Ybeg);
    for (int Y = Ybeg; Y < Yend; ++Y) {
      const Lerp ly = lerp_src(Y, h, sh);
      Vec8<T> pkr[kPairWin];
#pragma unroll
      for (int k = 0; k < kPairWin; ++k) pkr[k] = nx[k];
      if (Y + 1 < Yend) load_row(Y + 1);
      while (ly.i0 > cur_row) flush();
      float hx0[8], hx1[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) { hx0[c] = 0.f; hx1[c] = 0.f; }
#pragma unroll
      for (int k = 0; k < kPairWin; ++k) {
        if (k < tp.n) {
          float f[8];
          pkr[k].get(f);
          const float wa = tp.wa[k], wb = tp.wb[k];
#pragma unroll
          for (int c = 0; c < 8; ++c) { hx0[c] = fmaf(wa, f[c], hx0[c]); hx1[c] = fmaf(wb, f[c], hx1[c]); }
        }
      }
      // lower tap -> row i0, upper tap -> row i1 (i1 == i0 on the last input row); rows above the strip are not ours
      const float w_cur = (ly.i0 == cur_row ? ly.w0 : 0.f) + (ly.i1 == cur_row ? ly.w1 : 0.f);
      const float w_nxt = (ly.i1 == cur_row + 1) ? ly.w1 : 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        cur0[c] = fmaf(w_cur, hx0[c], cur0[c]); cur1[c] = fmaf(w_cur, hx1[c], cur1[c]);
        nxt0[c] = fmaf(w_nxt, hx0[c], nxt0[c]); nxt1[c] = fmaf(w_nxt, hx1[c], nxt1[c]);
      }
    }
    while (cur_row < y1) flush();
  }
}
// host mirror of the device tap window: widest union window of an input-column pair
static int upsample_pair_window(int w, int W, float sw) {
  auto first_out = [&](int i) {
    if (i <= 0) return 0;
    if (i >= w || sw <= 0.f) return W;
    int a = static_cast<int>(static_cast<float>(i) / sw);
    a = a < 0 ? 0 : (a > W ? W : a);
    while (a > 0 && static_cast<int>(sw * static_cast<float>(a - 1)) >= i) --a;
    while (a < W && static_cast<int>(sw * static_cast<float>(a)) < i) ++a;
    return a;
  };
  int widest = 0;
  for (int j = 0; j < (w + 1) / 2; ++j) {
    const int xa = 2 * j, xb = 2 * j + 1;
    const int lo = xa == 0 ? 0 : first_out(xa - 1);
    const int hi = first_out((xb < w - 1 ? xb : w - 1) + 1);
    widest = hi - lo > widest ? hi - lo : widest;
  }
  return widest;
}

int upsample_nhwc_fwd(int dtype, const void* x, void* y, int N, int h, int w, int H, int W, int C, cudaStream_t s) {
  PP_REQUIRE(C % 8 == 0, "upsample_nhwc_fwd: C=%d must be a multiple of 8", C);
  PP_REQUIRE_INT32(static_cast<long long>(N) * H * W * C, "upsample_nhwc_fwd");
  const int threads = W * (C / 8) >= 256 ? 256 : 128;
  PP_DISPATCH_T(dtype, upsample_nhwc_fwd_kernel<T><<<grid_for(static_cast<long long>(N) * H * 256, 256, 32), threads, 0, s>>>(
                           static_cast<const T*>(x), static_cast<T*>(y), N, h, w, H, W, C, ac_scale(h, H),
                           ac_scale(w, W)););
  PP_LAUNCH_CHECK();
  return PP_OK;
}
int upsample_nhwc_bwd(int dtype, const void* gy, void* gx, int N, int h, int w, int H, int W, int C, int accumulate,
                      cudaStream_t s) {
  PP_REQUIRE(C % 8 == 0, "upsample_nhwc_bwd: C=%d must be a multiple of 8", C);
  PP_REQUIRE_INT32(static_cast<long long>(N) * H * W * C, "upsample_nhwc_bwd");
  static const int strip_on = [] { const char* e = getenv("PP_UPSAMPLE_STRIP"); return (e && e[0] == '0') ? 0 : 1; }();
  const size_t smem_strip = static_cast<size_t>((w + 1) / 2) * sizeof(PairTaps);
  if (strip_on && H >= h && W > w && h > 1 && smem_strip <= 48 * 1024 &&
      upsample_pair_window(w, W, ac_scale(w, W)) <= kPairWin) {
    const int work = ((w + 1) / 2) * (C / 8);
    const int threads = work >= 256 ? 256 : (work >= 128 ? 128 : 64);
    const int ychunks = ceil_div(work, threads);
    int P = 8;
    while (P > 2 && static_cast<long long>(N) * ceil_div(h, P) * ychunks < 4LL * sm_count()) P >>= 1;
    const int strips = ceil_div(h, P);
    PP_DISPATCH_T(dtype, upsample_nhwc_bwd_strip_kernel<T><<<dim3(N * strips, ychunks), threads, smem_strip, s>>>(
                             static_cast<const T*>(gy), static_cast<T*>(gx), h, w, H, W, C, ac_scale(h, H),
                             ac_scale(w, W), P, strips, accumulate););
    PP_LAUNCH_CHECK();
    return PP_OK;
  }
  const int threads = w * (C / 8) >= 256 ? 256 : 128;
  const size_t smem = static_cast<size_t>(w) * (kUpWin + 2) * sizeof(float);
  PP_REQUIRE(smem <= 48 * 1024, "upsample_nhwc_bwd: input width %d too large", w);
  // (A separable variant that stages an fp32 row in 64 KB of shared memory halves the global loads, but its blocks can
  // no longer co-reside with the 200 KB conv CTAs this kernel overlaps with in the backward pass: 7.30 -> 7.51 ms/step.)
  PP_DISPATCH_T(dtype, upsample_nhwc_bwd_kernel<T><<<grid_for(static_cast<long long>(N) * h * 256, 256, 8), threads, smem, s>>>(
                           static_cast<const T*>(gy), static_cast<T*>(gx), N, h, w, H, W, C, ac_scale(h, H),
                           ac_scale(w, W), accumulate););
  PP_LAUNCH_CHECK();
  return PP_OK;
}

// NCHW fp32 planes (the aux-path logits, C = num_classes): x [NC][h][w] -> y [NC][H][W]. One block per output row.
__global__ void __launch_bounds__(256) upsample_planes_fwd_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                                  long long NC, int h, int w, int H, int W, float sh,
                                                                  float sw) {
  const long long rows = NC * H;
  for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
    const long long pl = row / H;
    const int Y = static_cast<int>(row - pl * H);
    const Lerp ly = lerp_src(Y, h, sh);
    const float* r0 = x + (pl * h + ly.i0) * w;
    const float* r1 = x + (pl * h + ly.i1) * w;
    for (int X = threadIdx.x; X < W; X += blockDim.x) {
      const Lerp lx = lerp_src(X, w, sw);
      y[row * W + X] = ly.w0 * (lx.w0 * __ldg(r0 + lx.i0) + lx.w1 * __ldg(r0 + lx.i1)) +
                       ly.w1 * (lx.w0 * __ldg(r1 + lx.i0) + lx.w1 * __ldg(r1 + lx.i1));
    }
  }
}
// Backward, separable inside a block: one block per INPUT row (plane, yi). Pass 1: tmp[X] = sum_Y wy(Y) gy[Y][X]
// (coalesced row reads, block-uniform vertical taps) into shared memory; pass 2: gx[yi][xi] = sum_X wx(X) tmp[X].
__global__ void __launch_bounds__(256) upsample_planes_bwd_kernel(const float* __restrict__ gy, float* __restrict__ gx,
                                                                  long long NC, int h, int w, int H, int W, float sh,
                                                                  float sw) {
  extern __shared__ float tmp[];   // [W]
  const long long rows = NC * h;
  for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
    const long long pl = row / h;
    const int yi = static_cast<int>(row - pl * h);
    int ylo, yhi;
    lerp_range(yi, H, sh, &ylo, &yhi);
    const float* g = gy + pl * H * W;
    for (int X = threadIdx.x; X < W; X += blockDim.x) {
      float a = 0.f;
      for (int Y = ylo; Y <= yhi; ++Y) {
        const float wy = lerp_weight(Y, yi, h, sh);   // block-uniform branch
        if (wy != 0.f) a = fmaf(wy, __ldg(g + static_cast<size_t>(Y) * W + X), a);
      }
      tmp[X] = a;
    }
    __syncthreads();
    for (int xi = threadIdx.x; xi < w; xi += blockDim.x) {
      int xlo, xhi;
      lerp_range(xi, W, sw, &xlo, &xhi);
      float a = 0.f;
      for (int X = xlo; X <= xhi; ++X) a = fmaf(lerp_weight(X, xi, w, sw), tmp[X], a);
      gx[row * w + xi] = a;
    }
    __syncthreads();
  }
}

int upsample_planes_fwd(const float* x, float* y, long long NC, int h, int w, int H, int W, cudaStream_t s) {
  upsample_planes_fwd_kernel<<<grid_for(NC * H * 256, 256, 32), W >= 256 ? 256 : 128, 0, s>>>(
      x, y, NC, h, w, H, W, ac_scale(h, H), ac_scale(w, W));
  PP_LAUNCH_CHECK();
  return PP_OK;
}
int upsample_planes_bwd(const float* gy, float* gx, long long NC, int h, int w, int H, int W, cudaStream_t s) {
  PP_REQUIRE(W <= 8192, "upsample_planes_bwd: W=%d too wide", W);
  upsample_planes_bwd_kernel<<<grid_for(NC * h * 256, 256, 32), W >= 256 ? 256 : 128, W * sizeof(float), s>>>(
      gy, gx, NC, h, w, H, W, ac_scale(h, H), ac_scale(w, W));
  PP_LAUNCH_CHECK();
  return PP_OK;
}

// ==============================================================================================
// Layout converts between the API's NCHW fp32 tensors and the internal NHWC T tensors.
// ==============================================================================================
template <typename T>
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float* __restrict__ src, T* __restrict__ dst, int C,
                                                           int HW) {
  __shared__ float tile[32][33];
  const long long n = blockIdx.z;
  const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r, p = p0 + tx;
    tile[r][tx] = (c < C && p < HW) ? src[(n * C + c) * HW + p] : 0.f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int p = p0 + r, c = c0 + tx;
    if (c < C && p < HW) dst[(n * HW + p) * C + c] = from_f32<T>(tile[tx][r]);
  }
}
template <typename T>
__global__ void __launch_bounds__(256) nhwc_to_nchw_kernel(const T* __restrict__ src, float* __restrict__ dst, int C,
                                                           int HW) {
  __shared__ float tile[32][33];
  const long long n = blockIdx.z;
  const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;
  for (int r = ty; r < 32; r += 8) {
    const int p = p0 + r, c = c0 + tx;
    tile[r][tx] = (c < C && p < HW) ? to_f32(src[(n * HW + p) * C + c]) : 0.f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r, p = p0 + tx;
    if (c < C && p < HW) dst[(n * C + c) * HW + p] = tile[tx][r];
  }
}
int nchw_to_nhwc(int dtype, const float* src, void* dst, int N, int C, int HW, cudaStream_t s) {
  dim3 grid(ceil_div(HW, 32), ceil_div(C, 32), N);
  PP_DISPATCH_T(dtype, nchw_to_nhwc_kernel<T><<<grid, 256, 0, s>>>(src, static_cast<T*>(dst), C, HW););
  PP_LAUNCH_CHECK();
  return PP_OK;
}
int nhwc_to_nchw(int dtype, const void* src, float* dst, int N, int C, int HW, cudaStream_t s) {
  dim3 grid(ceil_div(HW, 32), ceil_div(C, 32), N);
  PP_DISPATCH_T(dtype, nhwc_to_nchw_kernel<T><<<grid, 256, 0, s>>>(static_cast<const T*>(src), dst, C, HW););
  PP_LAUNCH_CHECK();
  return PP_OK;
}

// ==============================================================================================
// The strided-conv / transposed-conv UNet variant (unet.py:113-116 stride-2 first conv of an encoder block,
// unet.py:141 ConvTranspose2d(k = s, stride = s, bias=False) in the decoder) runs on the SAME stride-1 tcgen05
// kernels through two exact identities:
//   * conv3x3(stride 2, pad 1)(x) == conv3x3(stride 1, pad 1)(space_to_depth(x)) with the zero-embedded kernel
//       We[co, (sy,sx,c), ty, tx] = W[co, c, ky, kx],  ky -> (ty, sy): 0 -> (0,1), 1 -> (1,0), 2 -> (1,1)  (same for x)
//     because input row 2y + ky - 1 is block row y - 1 / sub-row 1 (ky = 0) or block row y / sub-row ky - 1;
//   * ConvTranspose2d(k = s = S)(x) == depth_to_space(conv1x1(x) to S*S*Cout channels), the 1x1 conv being the centre
//     tap of a 3x3 kernel:  We[(a*S+b)*Cout + co, ci, 1, 1] = Wt[ci, co, a, b].
// Zero taps contribute exactly 0, so results equal the direct computation up to summation order. The embedded fp32
// kernels are rebuilt from the master weights every forward pass; weight gradients are gathered back from the
// embedded gradient. space_to_depth / depth_to_space: big[n, 2Y+sy, 2X+sx, c] <-> small[n, Y, X, (sy*2+sx)*C + c].
// ==============================================================================================
template <typename T, bool TO_DEPTH>
__global__ void __launch_bounds__(256) space_depth_kernel(const T* __restrict__ src, T* __restrict__ dst, long long nvec,
                                                          int Hs, int Ws, int C, int accumulate) {
  const int cv = C / 8;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nvec;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % cv);
    long long r = i / cv;
    const int q = static_cast<int>(r % 4); r /= 4;
    const int X = static_cast<int>(r % Ws); r /= Ws;
    const int Y = static_cast<int>(r % Hs);
    const long long n = r / Hs;
    const long long big = (((n * 2 * Hs + 2 * Y + (q >> 1)) * 2 * Ws) + 2 * X + (q & 1)) * C + v * 8;
    const long long small = i * 8;
    Vec8<T> a;
    if (TO_DEPTH) {
      a.load(src + big);
      a.store(dst + small);
    } else {
      a.load(src + small);
      if (accumulate) {
        Vec8<T> b;
        b.load(dst + big);
        float fa[8], fb[8];
        a.get(fa); b.get(fb);
#pragma unroll
        for (int k = 0; k < 8; ++k) fa[k] += fb[k];
        a.set(fa);
      }
      a.store(dst + big);
    }
  }
}
// x [N, 2Hs, 2Ws, C] -> y [N, Hs, Ws, 4C]
int space_to_depth(int dtype, const void* x, void* y, int N, int Hs, int Ws, int C, cudaStream_t s) {
  PP_REQUIRE(C % 8 == 0 && C > 0, "space_to_depth: C=%d must be a multiple of 8", C);
  const long long nvec = static_cast<long long>(N) * Hs * Ws * 4 * (C / 8);
  PP_DISPATCH_T(dtype, (space_depth_kernel<T, true><<<grid_for(nvec, 256), 256, 0, s>>>(
                           static_cast<const T*>(x), static_cast<T*>(y), nvec, Hs, Ws, C, 0)););
  PP_LAUNCH_CHECK();
  return PP_OK;
}
// y [N, Hs, Ws, 4C] -> x [N, 2Hs, 2Ws, C] (accumulate: +=)
int depth_to_space(int dtype, const void* y, void* x, int N, int Hs, int Ws, int C, int accumulate, cudaStream_t s) {
  PP_REQUIRE(C % 8 == 0 && C > 0, "depth_to_space: C=%d must be a multiple of 8", C);
  const long long nvec = static_cast<long long>(N) * Hs * Ws * 4 * (C / 8);
  PP_DISPATCH_T(dtype, (space_depth_kernel<T, false><<<grid_for(nvec, 256), 256, 0, s>>>(
                           static_cast<const T*>(y), static_cast<T*>(x), nvec, Hs, Ws, C, accumulate)););
  PP_LAUNCH_CHECK();
  return PP_OK;
}

__device__ __forceinline__ int s2_k_of(int t, int sub) {   // embedded tap t, sub-row -> original ky (or -1: zero)
  return t == 0 ? (sub == 1 ? 0 : -1) : (t == 1 ? 1 + sub : -1);
}
__global__ void __launch_bounds__(256) embed_s2_weight_kernel(const float* __restrict__ w, float* __restrict__ we,
                                                              int Cout, int C) {
  const long long total = static_cast<long long>(Cout) * 4 * C * 9;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int t = static_cast<int>(i % 9);
    const int cc = static_cast<int>((i / 9) % (4 * C));
    const long long co = i / (9LL * 4 * C);
    const int q = cc / C, c = cc % C;
    const int ky = s2_k_of(t / 3, q >> 1), kx = s2_k_of(t % 3, q & 1);
    we[i] = (ky >= 0 && kx >= 0) ? w[((co * C + c) * 3 + ky) * 3 + kx] : 0.f;
  }
}
__global__ void __launch_bounds__(256) collapse_s2_wgrad_kernel(const float* __restrict__ dwe, float* __restrict__ dw,
                                                                int Cout, int C) {
  const long long total = static_cast<long long>(Cout) * C * 9;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(i % 9), ky = k / 3, kx = k % 3;
    const int c = static_cast<int>((i / 9) % C);
    const long long co = i / (9LL * C);
    const int ty = ky == 0 ? 0 : 1, sy = ky == 0 ? 1 : ky - 1;
    const int tx = kx == 0 ? 0 : 1, sx = kx == 0 ? 1 : kx - 1;
    dw[i] += dwe[((co * 4 * C + (sy * 2 + sx) * C + c) * 9) + ty * 3 + tx];
  }
}
int embed_s2_weight(const float* w, float* we, int Cout, int C, cudaStream_t s) {
  embed_s2_weight_kernel<<<grid_for(static_cast<long long>(Cout) * 4 * C * 9, 256), 256, 0, s>>>(w, we, Cout, C);
  PP_LAUNCH_CHECK();
  return PP_OK;
}
int collapse_s2_wgrad(const float* dwe, float* dw, int Cout, int C, cudaStream_t s) {
  collapse_s2_wgrad_kernel<<<grid_for(static_cast<long long>(Cout) * C * 9, 256), 256, 0, s>>>(dwe, dw, Cout, C);
  PP_LAUNCH_CHECK();
  return PP_OK;
}

// wt: ConvTranspose2d weight [Cin][Cout][S][S]; we: [S*S*Cout][Cin][3][3]
__global__ void __launch_bounds__(256) embed_ct_weight_kernel(const float* __restrict__ wt, float* __restrict__ we,
                                                              int Cin, int Cout, int S) {
  const long long total = static_cast<long long>(S) * S * Cout * Cin * 9;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int t = static_cast<int>(i % 9);
    const int ci = static_cast<int>((i / 9) % Cin);
    const long long r = i / (9LL * Cin);
    const int q = static_cast<int>(r / Cout), co = static_cast<int>(r % Cout);
    we[i] = t == 4 ? wt[((static_cast<long long>(ci) * Cout + co) * S + q / S) * S + q % S] : 0.f;
  }
}
__global__ void __launch_bounds__(256) collapse_ct_wgrad_kernel(const float* __restrict__ dwe, float* __restrict__ dwt,
                                                                int Cin, int Cout, int S) {
  const long long total = static_cast<long long>(Cin) * Cout * S * S;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int q = static_cast<int>(i % (S * S));
    const int co = static_cast<int>((i / (S * S)) % Cout);
    const long long ci = i / (static_cast<long long>(S) * S * Cout);
    dwt[i] += dwe[((static_cast<long long>(q) * Cout + co) * Cin + ci) * 9 + 4];
  }
}
int embed_ct_weight(const float* wt, float* we, int Cin, int Cout, int S, cudaStream_t s) {
  PP_REQUIRE(S == 1 || S == 2, "embed_ct_weight: kernel/stride %d unsupported (1 or 2)", S);
  embed_ct_weight_kernel<<<grid_for(static_cast<long long>(S) * S * Cout * Cin * 9, 256), 256, 0, s>>>(wt, we, Cin, Cout, S);
  PP_LAUNCH_CHECK();
  return PP_OK;
}
int collapse_ct_wgrad(const float* dwe, float* dwt, int Cin, int Cout, int S, cudaStream_t s) {
  collapse_ct_wgrad_kernel<<<grid_for(static_cast<long long>(Cin) * Cout * S * S, 256), 256, 0, s>>>(dwe, dwt, Cin, Cout, S);
  PP_LAUNCH_CHECK();
  return PP_OK;
}

// ==============================================================================================
// nn.Dropout2d of the aux path (aux_path_memory.py:23,31): whole channels of a sample are zeroed and the rest
// scaled by 1/(1-p). The caller supplies the per-(sample, channel) factors (0 or 1/(1-p)); the same kernel applies
// them to the activations in the forward pass and to their gradients in the backward pass (y may alias x).
// scale row n starts at scale + n * ld (ld >= C: the two concat sources share one [N][C0 + C1] table).
// ==============================================================================================
template <typename T>
__global__ void __launch_bounds__(256) channel_scale_kernel(const T* x, const float* __restrict__ scale, T* y,
                                                            long long nvec, int HW, int C, int ld) {
  const int cv = C / 8;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nvec;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % cv);
    const long long n = i / (static_cast<long long>(cv) * HW);
    const float* sc = scale + n * ld + v * 8;
    Vec8<T> a;
    a.load(x + i * 8);
    float f[8];
    a.get(f);
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] *= __ldg(sc + k);
    a.set(f);
    a.store(y + i * 8);
  }
}
int channel_scale(int dtype, const void* x, const float* scale, void* y, int N, int HW, int C, int ld, cudaStream_t s) {
  PP_REQUIRE(C % 8 == 0 && C > 0 && ld >= C, "channel_scale: C=%d (multiple of 8) / ld=%d", C, ld);
  const long long nvec = static_cast<long long>(N) * HW * (C / 8);
  PP_DISPATCH_T(dtype, (channel_scale_kernel<T><<<grid_for(nvec, 256), 256, 0, s>>>(
                           static_cast<const T*>(x), scale, static_cast<T*>(y), nvec, HW, C, ld)););
  PP_LAUNCH_CHECK();
  return PP_OK;
}

// ==============================================================================================
// Strong colour augmentation of the two-stream dataset on the device: Brightness -> Contrast ->
// GammaAugmentation(retain_stats) of datasets/augmentations.py:98-166, applied by CHAOSTwoStream.__getitem__
// (chaos_dataset.py:68-75) to the base-transformed slice. Each transform needs whole-image statistics (mean / std /
// min / max), so one 1024-thread block owns one image (<= 256 KB, L2 resident) and walks it up to four times:
//   pass 1: stats of x1 = x + b                         (Brightness; Contrast's mean_/min_/max_)
//   pass 2: stats of x2 = clip((x1 - mean1) a + mean1, min1, max1)       (Gamma's mean_/std_/min_/max_)
//   pass 3: stats of x3 = ((x2 - min2) / (max2 - min2 + eps)) ^ gamma    (retain_stats: its mean / std)
//   pass 4: out = (x3 - mean3) / (std3 + eps) * std2 + mean2
// params per image (8 floats): [apply_brightness, b, apply_contrast, a, apply_gamma, gamma, 0, 0]; the caller draws
// them (pacingpseudo_b200.data.sample_strong_params mirrors the reference's probabilities and ranges).
// ==============================================================================================
struct ImgStats { double sum, sq; float mn, mx; };
__device__ __forceinline__ ImgStats block_stats(double sum, double sq, float mn, float mx) {
  __shared__ double s_sum[32], s_sq[32];
  __shared__ float s_mn[32], s_mx[32];
  __syncthreads();   // protects the shared arrays between consecutive calls
  sum = warp_sum_d(sum);
  sq = warp_sum_d(sq);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (lane == 0) { s_sum[warp] = sum; s_sq[warp] = sq; s_mn[warp] = mn; s_mx[warp] = mx; }
  __syncthreads();
  ImgStats r{0.0, 0.0, INFINITY, -INFINITY};
  for (int w = 0; w < nw; ++w) {
    r.sum += s_sum[w]; r.sq += s_sq[w];
    r.mn = fminf(r.mn, s_mn[w]); r.mx = fmaxf(r.mx, s_mx[w]);
  }
  return r;
}
__global__ void __launch_bounds__(1024) strong_augment_kernel(const float* __restrict__ image,
                                                              const float* __restrict__ params,
                                                              float* __restrict__ out, int HW) {
  const float* x = image + static_cast<size_t>(blockIdx.x) * HW;
  float* y = out + static_cast<size_t>(blockIdx.x) * HW;
  const float* p = params + blockIdx.x * 8;
  const bool do_c = p[2] != 0.f, do_g = p[4] != 0.f;
  const float b = p[0] != 0.f ? p[1] : 0.f, a = p[3], gamma = p[5];
  const float eps = 1e-8f;
  const double n = static_cast<double>(HW);
  float mean1 = 0.f, min1 = 0.f, max1 = 0.f;
  if (do_c) {
    double s = 0.0, q = 0.0;
    float mn = INFINITY, mx = -INFINITY;
    for (int i = threadIdx.x; i < HW; i += blockDim.x) {
      const float v = x[i] + b;
      s += v; mn = fminf(mn, v); mx = fmaxf(mx, v);
    }
    const ImgStats st = block_stats(s, q, mn, mx);
    mean1 = static_cast<float>(st.sum / n); min1 = st.mn; max1 = st.mx;
  }
  auto x2_of = [&](float v) {
    v += b;
    return do_c ? fminf(fmaxf((v - mean1) * a + mean1, min1), max1) : v;
  };
  if (!do_g) {
    for (int i = threadIdx.x; i < HW; i += blockDim.x) y[i] = x2_of(x[i]);
    return;
  }
  double s = 0.0, q = 0.0;
  float mn = INFINITY, mx = -INFINITY;
  for (int i = threadIdx.x; i < HW; i += blockDim.x) {
    const float v = x2_of(x[i]);
    s += v; q += static_cast<double>(v) * v; mn = fminf(mn, v); mx = fmaxf(mx, v);
  }
  const ImgStats s2 = block_stats(s, q, mn, mx);
  const float mean2 = static_cast<float>(s2.sum / n);
  const float std2 = static_cast<float>(sqrt(fmax(s2.sq / n - (s2.sum / n) * (s2.sum / n), 0.0)));
  const float min2 = s2.mn, inv_rng = 1.f / (s2.mx - s2.mn + eps);
  auto x3_of = [&](float v) { return powf((x2_of(v) - min2) * inv_rng, gamma); };
  s = 0.0; q = 0.0;
  for (int i = threadIdx.x; i < HW; i += blockDim.x) {
    const float v = x3_of(x[i]);
    s += v; q += static_cast<double>(v) * v;
  }
  const ImgStats s3 = block_stats(s, q, 0.f, 0.f);
  const float mean3 = static_cast<float>(s3.sum / n);
  const float std3 = static_cast<float>(sqrt(fmax(s3.sq / n - (s3.sum / n) * (s3.sum / n), 0.0)));
  const float k = std2 / (std3 + eps);
  for (int i = threadIdx.x; i < HW; i += blockDim.x) y[i] = (x3_of(x[i]) - mean3) * k + mean2;
}
int strong_color_augment(const float* image, const float* params, float* out, int N, int HW, cudaStream_t s) {
  PP_REQUIRE(N >= 1 && HW >= 1, "strong_color_augment: bad shape N=%d HW=%d", N, HW);
  strong_augment_kernel<<<N, 1024, 0, s>>>(image, params, out, HW);
  PP_LAUNCH_CHECK();
  return PP_OK;
}

// ==============================================================================================
// Adam with L2 weight decay folded into the gradient (torch.optim.Adam semantics,
// train_chaos.py:219), over one flat fp32 parameter buffer. step is the 1-based step count.
// ==============================================================================================
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v, long long n, float lr,
                                                   float beta1, float beta2, float eps, float wd, float bc1,
                                                   float bc2_sqrt, float grad_scale) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float pv = p[i];
    const float gv = fmaf(wd, pv, g[i] * grad_scale);
    const float mv = fmaf(beta1, m[i], (1.f - beta1) * gv);
    const float vv = fmaf(beta2, v[i], (1.f - beta2) * gv * gv);
    m[i] = mv;
    v[i] = vv;
    const float denom = sqrtf(vv) / bc2_sqrt + eps;
    p[i] = pv - (lr / bc1) * (mv / denom);
  }
}
int adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
              float wd, int step, float grad_scale, cudaStream_t s) {
  PP_REQUIRE(step >= 1, "adam_step: step must be >= 1");
  const double bc1 = 1.0 - pow(static_cast<double>(beta1), step);
  const double bc2 = 1.0 - pow(static_cast<double>(beta2), step);
  adam_kernel<<<grid_for(n, 256), 256, 0, s>>>(p, g, m, v, n, lr, beta1, beta2, eps, wd, static_cast<float>(bc1),
                                               static_cast<float>(sqrt(bc2)), grad_scale);
  PP_LAUNCH_CHECK();
  return PP_OK;
}

}  // namespace pp
