// ops.cu — the HBM-bound operators around the convolutions, all on NHWC activations of type T
// (bf16 by default, fp32 in the fp32 precision mode):
//   weight (un)packing, first conv (Cin = 1), 1x1 heads, BatchNorm statistics / finalize / apply /
//   backward (train-mode batch statistics per statistics group, or eval-mode running statistics),
//   LeakyReLU(0.01), MaxPool2d(2,2), bilinear align_corners=True upsampling, NCHW<->NHWC converts,
//   Adam. Reference call sites: /root/reference/models/unet.py:60,109,144,188-193,
//   /root/reference/models/aux_path_memory.py:22-33,52 and /root/reference/train_chaos.py:219.
#include "pp_common.cuh"

namespace pp {

static inline int grid_for(long long work, int block, int max_blocks_per_sm = 16) {
  long long g = ceil_div_ll(work, block);
  long long cap = static_cast<long long>(sm_count()) * max_blocks_per_sm;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

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
// ==============================================================================================
template <typename T>
__global__ void __launch_bounds__(256) first_conv_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                             const float* __restrict__ bias, T* __restrict__ y, int N,
                                                             int H, int W, int Cout) {
  extern __shared__ float sw[];  // [Cout][9] + [Cout]
  for (int i = threadIdx.x; i < Cout * 9; i += blockDim.x) sw[i] = w[i];
  for (int i = threadIdx.x; i < Cout; i += blockDim.x) sw[Cout * 9 + i] = bias ? bias[i] : 0.f;
  __syncthreads();
  const int vecs = Cout / 8;
  const int total = N * H * W * vecs;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int v = i % vecs;
    const int p = i / vecs;
    const int px = p % W, py = (p / W) % H;
    const int img = p / (W * H);
    float xin[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int yy = py + t / 3 - 1, xx = px + t % 3 - 1;
      xin[t] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(x + (static_cast<size_t>(img) * H + yy) * W + xx) : 0.f;
    }
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int co = v * 8 + j;
      float a = sw[Cout * 9 + co];
#pragma unroll
      for (int t = 0; t < 9; ++t) a = fmaf(xin[t], sw[co * 9 + t], a);
      o[j] = a;
    }
    Vec8<T> pk;
    pk.set(o);
    pk.store(y + static_cast<size_t>(p) * Cout + v * 8);
  }
}

int first_conv_fwd(int dtype, const float* x, const float* w, const float* bias, void* y, int N, int H, int W, int Cout,
                   cudaStream_t s) {
  PP_REQUIRE(Cout % 8 == 0 && Cout <= 256, "first_conv_fwd: Cout=%d unsupported", Cout);
  const long long total = static_cast<long long>(N) * H * W * (Cout / 8);
  PP_REQUIRE_INT32(total * 8, "first_conv_fwd");
  PP_DISPATCH_T(dtype, first_conv_fwd_kernel<T><<<grid_for(total, 256), 256, (Cout * 10) * sizeof(float), s>>>(
                           x, w, bias, static_cast<T*>(y), N, H, W, Cout););
  PP_LAUNCH_CHECK();
  return PP_OK;
}

// dW[co][tap] += sum_p dy[p][co] * x[p + off(tap)]   (dw is the OIHW fp32 grad [Cout][1][3][3])
template <typename T>
__global__ void __launch_bounds__(256) first_conv_wgrad_kernel(const T* __restrict__ dy, const float* __restrict__ x,
                                                               float* __restrict__ dw, int N, int H, int W, int Cout) {
  extern __shared__ float sacc[];  // [Cout*9]
  for (int i = threadIdx.x; i < Cout * 9; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warp_id = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int P = N * H * W;
  const int chunk = (P + nwarps - 1) / nwarps;
  const int pb = min(warp_id * chunk, P), pe = min(pb + chunk, P);
  for (int cb = 0; cb < Cout; cb += 32) {  // channel block handled by lane
    const int co = cb + lane;
    float acc[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[t] = 0.f;
    constexpr int U = 4;  // pixels in flight per warp (independent loads)
    int cx = pb % W, cy = (pb / W) % H, cimg = pb / (W * H);   // running coordinates of pixel p0 (no per-pixel division)
    for (int p0 = pb; p0 < pe; p0 += U) {
      float xv[U], g[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int p = p0 + u;
        xv[u] = 0.f;
        g[u] = 0.f;
        if (p < pe) {
          const int px = cx, py = cy, img = cimg;
          if (++cx == W) { cx = 0; if (++cy == H) { cy = 0; ++cimg; } }
          if (lane < 9) {
            const int yy = py + lane / 3 - 1, xx = px + lane % 3 - 1;
            if (yy >= 0 && yy < H && xx >= 0 && xx < W) xv[u] = __ldg(x + (static_cast<size_t>(img) * H + yy) * W + xx);
          }
          if (co < Cout) g[u] = to_f32(dy[static_cast<size_t>(p) * Cout + co]);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
#pragma unroll
        for (int t = 0; t < 9; ++t) acc[t] = fmaf(g[u], __shfl_sync(0xffffffffu, xv[u], t), acc[t]);
      }
    }
    if (co < Cout) {
#pragma unroll
      for (int t = 0; t < 9; ++t) atomicAdd(&sacc[co * 9 + t], acc[t]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Cout * 9; i += blockDim.x) atomicAdd(dw + i, sacc[i]);
}

int first_conv_wgrad(int dtype, const void* dy, const float* x, float* dw, int N, int H, int W, int Cout,
                     cudaStream_t s) {
  PP_REQUIRE(Cout <= 256, "first_conv_wgrad: Cout=%d unsupported", Cout);
  PP_REQUIRE_INT32(static_cast<long long>(N) * H * W, "first_conv_wgrad");
  const int blocks = sm_count() * 4;
  PP_DISPATCH_T(dtype, first_conv_wgrad_kernel<T><<<blocks, 256, Cout * 9 * sizeof(float), s>>>(
                           static_cast<const T*>(dy), x, dw, N, H, W, Cout););
  PP_LAUNCH_CHECK();
  return PP_OK;
}

// ==============================================================================================
// 1x1 heads (unet.py:60 final_conv with bias; aux_path_memory.py:32 fc_cls without bias).
// Input NHWC T [P][Cin], output logits NCHW fp32 [N][C][HW]. C <= 8.
// ==============================================================================================
constexpr int kMaxClasses = 8;

template <typename T, int CIN>
__global__ void __launch_bounds__(256) head_fwd_kernel(const T* __restrict__ a, const float* __restrict__ w,
                                                       const float* __restrict__ bias, float* __restrict__ logits,
                                                       int P, int HW, int C) {
  __shared__ float sw[kMaxClasses * CIN + kMaxClasses];
  for (int i = threadIdx.x; i < C * CIN; i += blockDim.x) sw[i] = w[i];
  for (int i = threadIdx.x; i < C; i += blockDim.x) sw[kMaxClasses * CIN + i] = bias ? bias[i] : 0.f;
  __syncthreads();
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < P; p += gridDim.x * blockDim.x) {
    float acc[kMaxClasses];
#pragma unroll
    for (int c = 0; c < kMaxClasses; ++c) acc[c] = (c < C) ? sw[kMaxClasses * CIN + c] : 0.f;
#pragma unroll
    for (int v = 0; v < CIN / 8; ++v) {
      Vec8<T> pk;
      pk.load(a + static_cast<size_t>(p) * CIN + v * 8);
      float f[8];
      pk.get(f);
#pragma unroll
      for (int c = 0; c < kMaxClasses; ++c)
        if (c < C) {
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[c] = fmaf(f[j], sw[c * CIN + v * 8 + j], acc[c]);
        }
    }
    const int n = p / HW, hw = p % HW;
#pragma unroll
    for (int c = 0; c < kMaxClasses; ++c)
      if (c < C) logits[(static_cast<size_t>(n) * C + c) * HW + hw] = acc[c];
  }
}

int head_fwd(int dtype, const void* a, const float* w, const float* bias, float* logits, long long P, int HW, int Cin,
             int C, cudaStream_t s) {
  PP_REQUIRE(C >= 1 && C <= kMaxClasses, "head_fwd: num_classes=%d unsupported (max %d)", C, kMaxClasses);
  PP_REQUIRE(Cin == 32 || Cin == 64 || Cin == 128 || Cin == 256, "head_fwd: Cin=%d unsupported (32/64/128/256)", Cin);
  PP_REQUIRE_INT32(P * Cin, "head_fwd");
#define PP_HEAD_FWD(CIN_) \
  head_fwd_kernel<T, CIN_><<<grid_for(P, 256), 256, 0, s>>>(static_cast<const T*>(a), w, bias, logits, int(P), HW, C)
  PP_DISPATCH_T(dtype, if (Cin == 32) PP_HEAD_FWD(32); else if (Cin == 64) PP_HEAD_FWD(64);
                else if (Cin == 128) PP_HEAD_FWD(128); else PP_HEAD_FWD(256););
#undef PP_HEAD_FWD
  PP_LAUNCH_CHECK();
  return PP_OK;
}

// da[p][ci] = sum_c dl[c][p] * w[c][ci]
template <typename T, int CIN>
__global__ void __launch_bounds__(256) head_bwd_data_kernel(const float* __restrict__ dlogits,
                                                            const float* __restrict__ w, T* __restrict__ da,
                                                            int P, int HW, int C) {
  __shared__ float sw[kMaxClasses * CIN];
  for (int i = threadIdx.x; i < C * CIN; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < P; p += gridDim.x * blockDim.x) {
    const int n = p / HW, hw = p % HW;
    float dl[kMaxClasses];
#pragma unroll
    for (int c = 0; c < kMaxClasses; ++c) dl[c] = (c < C) ? dlogits[(static_cast<size_t>(n) * C + c) * HW + hw] : 0.f;
#pragma unroll
    for (int v = 0; v < CIN / 8; ++v) {
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < kMaxClasses; ++c)
          if (c < C) acc = fmaf(dl[c], sw[c * CIN + v * 8 + j], acc);
        f[j] = acc;
      }
      Vec8<T> pk;
      pk.set(f);
      pk.store(da + static_cast<size_t>(p) * CIN + v * 8);
    }
  }
}

// dW[c][ci] += sum_p dl[c][p] * a[p][ci];  db[c] += sum_p dl[c][p]
template <typename T, int CIN>
__global__ void __launch_bounds__(256) head_bwd_weight_kernel(const float* __restrict__ dlogits,
                                                              const T* __restrict__ a, float* __restrict__ dw,
                                                              float* __restrict__ db, int P, int HW, int C) {
  __shared__ float sacc[kMaxClasses * CIN + kMaxClasses];
  for (int i = threadIdx.x; i < kMaxClasses * CIN + kMaxClasses; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  constexpr int R = CIN / 32;
  const int lane = threadIdx.x & 31;
  const int warp_id = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int chunk = (P + nwarps - 1) / nwarps;
  const int pb = min(warp_id * chunk, P), pe = min(pb + chunk, P);
  float acc[kMaxClasses][R];
  float accb = 0.f;
#pragma unroll
  for (int c = 0; c < kMaxClasses; ++c)
#pragma unroll
    for (int r = 0; r < R; ++r) acc[c][r] = 0.f;
  constexpr int U = 4;  // pixels in flight per warp
  int cn = pb / HW, chw = pb % HW;   // running (image, pixel) of p0
  for (int p0 = pb; p0 < pe; p0 += U) {
    float dlv[U], av[U][R];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int p = p0 + u;
      dlv[u] = 0.f;
#pragma unroll
      for (int r = 0; r < R; ++r) av[u][r] = 0.f;
      if (p < pe) {
        const int n = cn, hw = chw;
        if (++chw == HW) { chw = 0; ++cn; }
        if (lane < C) dlv[u] = dlogits[(static_cast<size_t>(n) * C + lane) * HW + hw];
#pragma unroll
        for (int r = 0; r < R; ++r) av[u][r] = to_f32(a[static_cast<size_t>(p) * CIN + r * 32 + lane]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      accb += dlv[u];
#pragma unroll
      for (int c = 0; c < kMaxClasses; ++c) {
        const float d = __shfl_sync(0xffffffffu, dlv[u], c);
#pragma unroll
        for (int r = 0; r < R; ++r) acc[c][r] = fmaf(d, av[u][r], acc[c][r]);
      }
    }
  }
#pragma unroll
  for (int c = 0; c < kMaxClasses; ++c)
    if (c < C) {
#pragma unroll
      for (int r = 0; r < R; ++r) atomicAdd(&sacc[c * CIN + r * 32 + lane], acc[c][r]);
    }
  if (lane < C) atomicAdd(&sacc[kMaxClasses * CIN + lane], accb);
  __syncthreads();
  for (int i = threadIdx.x; i < C * CIN; i += blockDim.x) atomicAdd(dw + i, sacc[i]);
  if (db != nullptr)
    for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(db + i, sacc[kMaxClasses * CIN + i]);
}

int head_bwd(int dtype, const float* dlogits, const void* a, const float* w, void* da, float* dw, float* db, long long P,
             int HW, int Cin, int C, cudaStream_t s) {
  PP_REQUIRE(C >= 1 && C <= kMaxClasses, "head_bwd: num_classes=%d unsupported", C);
  PP_REQUIRE(Cin == 32 || Cin == 64 || Cin == 128 || Cin == 256, "head_bwd: Cin=%d unsupported (32/64/128/256)", Cin);
  PP_REQUIRE_INT32(P * Cin, "head_bwd");
  const int wblocks = sm_count() * 4;
#define PP_HEAD_BWD(CIN_)                                                                                          \
  do {                                                                                                             \
    if (da) head_bwd_data_kernel<T, CIN_><<<grid_for(P, 256), 256, 0, s>>>(dlogits, w, static_cast<T*>(da), int(P), HW, C); \
    head_bwd_weight_kernel<T, CIN_><<<wblocks, 256, 0, s>>>(dlogits, static_cast<const T*>(a), dw, db, int(P), HW, C);   \
  } while (0)
  PP_DISPATCH_T(dtype, if (Cin == 32) PP_HEAD_BWD(32); else if (Cin == 64) PP_HEAD_BWD(64);
                else if (Cin == 128) PP_HEAD_BWD(128); else PP_HEAD_BWD(256););
#undef PP_HEAD_BWD
  PP_LAUNCH_CHECK_N(da ? 2 : 1);
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
template <typename T>
__global__ void __launch_bounds__(256) bn_apply_kernel(const T* __restrict__ y, const float* __restrict__ coef,
                                                       T* __restrict__ a, int Pg, int C, int chunk,
                                                       int chunks_per_group, float slope) {
  constexpr int U = 4;
  const int vecs = C / 8;
  const int pl = 256 / vecs;
  const int v = threadIdx.x % vecs, l = threadIdx.x / vecs;
  const int g = blockIdx.x / chunks_per_group;
  const int p0 = (blockIdx.x % chunks_per_group) * chunk;
  const int p1 = min(p0 + chunk, Pg);
  if (l >= pl) return;
  float sc[8], sh[8];
  const float* cf = coef + static_cast<size_t>(g) * 4 * C + v * 8;
#pragma unroll
  for (int j = 0; j < 8; ++j) { sc[j] = cf[j]; sh[j] = cf[C + j]; }
  const size_t base = static_cast<size_t>(g) * Pg * C + v * 8;
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

static void bn_chunks(int G, long long Pg, int blocks_per_sm, int* chunk, int* cpg) {
  int c = (sm_count() * blocks_per_sm) / G;
  if (c < 1) c = 1;
  long long ch = ceil_div_ll(Pg, c);
  if (ch < 64) ch = 64;
  *chunk = static_cast<int>(ch);
  *cpg = static_cast<int>(ceil_div_ll(Pg, ch));
}

int bn_apply(int dtype, const void* y, const float* coef, void* a, int G, long long Pg, int C, float slope,
             cudaStream_t s) {
  PP_REQUIRE(C % 8 == 0 && C / 8 <= 256 && 256 % (C / 8) == 0, "bn_apply: C=%d unsupported", C);
  PP_REQUIRE_INT32(Pg * C, "bn_apply");
  int chunk, cpg;
  bn_chunks(G, Pg, 8, &chunk, &cpg);
  PP_DISPATCH_T(dtype, bn_apply_kernel<T><<<G * cpg, 256, 0, s>>>(static_cast<const T*>(y), coef, static_cast<T*>(a),
                                                                  int(Pg), C, chunk, cpg, slope););
  PP_LAUNCH_CHECK();
  return PP_OK;
}

// Backward reduce: bsums[g][c][0..1] += sum dz, sum dz * xhat   with  z = y*scale+shift,
// dz = da * (z > 0 ? 1 : slope),  xhat = (y - mean) * rstd.
template <typename T>
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const T* __restrict__ da, const T* __restrict__ y,
                                                            const float* __restrict__ coef, double* __restrict__ bsums,
                                                            long long Pg, int C, long long chunk, int chunks_per_group,
                                                            float slope) {
  __shared__ float red[256][17];
  const int vecs = C / 8;
  const int pl = 256 / vecs;
  const int v = threadIdx.x % vecs, l = threadIdx.x / vecs;
  const int g = blockIdx.x / chunks_per_group;
  const long long p0 = static_cast<long long>(blockIdx.x % chunks_per_group) * chunk;
  const long long p1 = (p0 + chunk < Pg) ? p0 + chunk : Pg;
  float s[8], ss[8], sc[8], sh[8], mu[8], rs[8];
  const float* cf = coef + static_cast<long long>(g) * 4 * C + v * 8;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    s[j] = 0.f; ss[j] = 0.f;
    sc[j] = cf[j]; sh[j] = cf[C + j]; mu[j] = cf[2 * C + j]; rs[j] = cf[3 * C + j];
  }
  if (l < pl) {
    constexpr int U = 4;
    const long long base = (static_cast<long long>(g) * Pg) * C + v * 8;
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
  for (int t = threadIdx.x; t < 2 * C; t += 256) {
    const int c = t % C, st = t / C;
    double acc = 0.0;
    for (int k = 0; k < pl; ++k) acc += static_cast<double>(red[k * vecs + c / 8][st * 8 + c % 8]);
    atomicAdd(bsums + (static_cast<long long>(g) * C + c) * 2 + st, acc);
  }
}

// Backward finalize: parameter grads (+=) and the per-group coefficients of the apply pass.
// bcoef layout [G][2][C]: k1 = sum dz / Pg, k2 = sum dz*xhat / Pg  (zeros in eval mode).
__global__ void bn_bwd_finalize_kernel(const double* __restrict__ bsums, const float* __restrict__ coef,
                                       float* __restrict__ dgamma, float* __restrict__ dbeta,
                                       float* __restrict__ dbias, float* __restrict__ bcoef, int G, long long Pg, int C,
                                       int training) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double dg = 0.0, dbt = 0.0, dbs = 0.0;
  for (int g = 0; g < G; ++g) {
    const double sdz = bsums[(static_cast<long long>(g) * C + c) * 2];
    const double sdx = bsums[(static_cast<long long>(g) * C + c) * 2 + 1];
    dg += sdx;
    dbt += sdz;
    float* bc = bcoef + static_cast<long long>(g) * 2 * C;
    if (training) {
      bc[c] = static_cast<float>(sdz / static_cast<double>(Pg));
      bc[C + c] = static_cast<float>(sdx / static_cast<double>(Pg));
    } else {
      bc[c] = 0.f;
      bc[C + c] = 0.f;
      dbs += sdz * static_cast<double>(coef[static_cast<long long>(g) * 4 * C + c]);  // scale * sum dz
    }
  }
  dgamma[c] += static_cast<float>(dg);
  dbeta[c] += static_cast<float>(dbt);
  if (dbias != nullptr) dbias[c] += static_cast<float>(dbs);  // exactly 0 under batch statistics
}

// dy = scale * (dz - k1 - xhat * k2); same (group, chunk) decomposition, coefficients in registers
template <typename T>
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const T* __restrict__ da, const T* __restrict__ y,
                                                           const float* __restrict__ coef,
                                                           const float* __restrict__ bcoef, T* __restrict__ dy, int Pg,
                                                           int C, int chunk, int chunks_per_group, float slope) {
  constexpr int U = 2;
  const int vecs = C / 8;
  const int pl = 256 / vecs;
  const int v = threadIdx.x % vecs, l = threadIdx.x / vecs;
  const int g = blockIdx.x / chunks_per_group;
  const int p0 = (blockIdx.x % chunks_per_group) * chunk;
  const int p1 = min(p0 + chunk, Pg);
  if (l >= pl) return;
  float sc[8], sh[8], mu[8], rs[8], k1[8], k2[8];
  const float* cf = coef + static_cast<size_t>(g) * 4 * C + v * 8;
  const float* bc = bcoef + static_cast<size_t>(g) * 2 * C + v * 8;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = cf[j]; sh[j] = cf[C + j]; mu[j] = cf[2 * C + j]; rs[j] = cf[3 * C + j];
    k1[j] = bc[j]; k2[j] = bc[C + j];
  }
  const size_t base = static_cast<size_t>(g) * Pg * C + v * 8;
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
           cudaStream_t s) {
  PP_REQUIRE(C % 8 == 0 && C / 8 <= 256 && 256 % (C / 8) == 0, "bn_bwd: C=%d unsupported", C);
  PP_REQUIRE_INT32(Pg * C, "bn_bwd");
  PP_CHECK_CUDA(cudaMemsetAsync(bsums, 0, sizeof(double) * 2 * G * C, s));
  int cpg = (sm_count() * 4) / G;
  if (cpg < 1) cpg = 1;
  long long chunk = ceil_div_ll(Pg, cpg);
  if (chunk < 256) chunk = 256;
  cpg = static_cast<int>(ceil_div_ll(Pg, chunk));
  int achunk, acpg;
  bn_chunks(G, Pg, 8, &achunk, &acpg);
  PP_DISPATCH_T(dtype,
                bn_bwd_reduce_kernel<T><<<G * cpg, 256, 0, s>>>(static_cast<const T*>(da), static_cast<const T*>(y),
                                                                coef, bsums, Pg, C, chunk, cpg, slope);
                bn_bwd_finalize_kernel<<<ceil_div(C, 128), 128, 0, s>>>(bsums, coef, dgamma, dbeta, dbias, bcoef, G, Pg,
                                                                        C, training);
                bn_bwd_apply_kernel<T><<<G * acpg, 256, 0, s>>>(static_cast<const T*>(da), static_cast<const T*>(y),
                                                                coef, bcoef, static_cast<T*>(dy), int(Pg), C, achunk,
                                                                acpg, slope););
  PP_LAUNCH_CHECK_N(3);
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
struct Lerp { int i0, i1; float w0, w1; };
__device__ __forceinline__ Lerp lerp_src(int dst, int in_size, float scale) {
  const float r = scale * static_cast<float>(dst);
  Lerp l;
  l.i0 = static_cast<int>(r);
  l.i1 = l.i0 + ((l.i0 < in_size - 1) ? 1 : 0);
  l.w1 = r - static_cast<float>(l.i0);
  l.w0 = 1.f - l.w1;
  return l;
}
static inline float ac_scale(int in_size, int out_size) {
  return out_size > 1 ? static_cast<float>(in_size - 1) / static_cast<float>(out_size - 1) : 0.f;
}
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

template <typename T>
__global__ void __launch_bounds__(256) upsample_nhwc_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int N,
                                                                int h, int w, int H, int W, int C, float sh, float sw) {
  const int vecs = C / 8;
  const int total = N * H * W * vecs;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int v = i % vecs;
    const int p = i / vecs;
    const int X = p % W, Y = (p / W) % H;
    const size_t n = p / (W * H);
    const Lerp ly = lerp_src(Y, h, sh), lx = lerp_src(X, w, sw);
    const T* b = x + n * h * w * C + v * 8;
    Vec8<T> p00, p01, p10, p11;
    p00.load(b + (static_cast<long long>(ly.i0) * w + lx.i0) * C);
    p01.load(b + (static_cast<long long>(ly.i0) * w + lx.i1) * C);
    p10.load(b + (static_cast<long long>(ly.i1) * w + lx.i0) * C);
    p11.load(b + (static_cast<long long>(ly.i1) * w + lx.i1) * C);
    float f00[8], f01[8], f10[8], f11[8], o[8];
    p00.get(f00); p01.get(f01); p10.get(f10); p11.get(f11);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      o[j] = ly.w0 * (lx.w0 * f00[j] + lx.w1 * f01[j]) + ly.w1 * (lx.w0 * f10[j] + lx.w1 * f11[j]);
    Vec8<T> out;
    out.set(o);
    out.store(y + static_cast<size_t>(i) * 8);
  }
}

// gx[n,i,j,:] = sum over outputs (Y,X) reading (i,j) of wy*wx*gy[n,Y,X,:]  (gather, deterministic)
template <typename T>
__global__ void __launch_bounds__(256) upsample_nhwc_bwd_kernel(const T* __restrict__ gy, T* __restrict__ gx, int N,
                                                                int h, int w, int H, int W, int C, float sh, float sw,
                                                                int accumulate) {
  const int vecs = C / 8;
  const int total = N * h * w * vecs;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int v = i % vecs;
    const int p = i / vecs;
    const int xi = p % w, yi = (p / w) % h;
    const size_t n = p / (w * h);
    int ylo, yhi, xlo, xhi;
    lerp_range(yi, H, sh, &ylo, &yhi);
    lerp_range(xi, W, sw, &xlo, &xhi);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    constexpr int KMAX = 8;   // candidate window per axis for scale factors >= 2 (wider windows take the slow path)
    if (yhi - ylo < KMAX && xhi - xlo < KMAX) {
      float wyv[KMAX], wxv[KMAX];
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        wyv[k] = (ylo + k <= yhi) ? lerp_weight(ylo + k, yi, h, sh) : 0.f;
        wxv[k] = (xlo + k <= xhi) ? lerp_weight(xlo + k, xi, w, sw) : 0.f;
      }
#pragma unroll
      for (int a = 0; a < KMAX; ++a) {
        if (wyv[a] == 0.f) continue;
#pragma unroll
        for (int b = 0; b < KMAX; ++b) {
          if (wxv[b] == 0.f) continue;
          Vec8<T> pk;
          pk.load(gy + ((n * H + ylo + a) * W + xlo + b) * C + v * 8);
          float f[8];
          pk.get(f);
          const float ww = wyv[a] * wxv[b];
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = fmaf(ww, f[j], acc[j]);
        }
      }
    } else {
      for (int Y = ylo; Y <= yhi; ++Y) {
        const float wy = lerp_weight(Y, yi, h, sh);
        if (wy == 0.f) continue;
        for (int X = xlo; X <= xhi; ++X) {
          const float wx = lerp_weight(X, xi, w, sw);
          if (wx == 0.f) continue;
          Vec8<T> pk;
          pk.load(gy + ((n * H + Y) * W + X) * C + v * 8);
          float f[8];
          pk.get(f);
          const float ww = wy * wx;
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = fmaf(ww, f[j], acc[j]);
        }
      }
    }
    Vec8<T> out;
    if (accumulate) {
      float old[8];
      out.load(gx + static_cast<size_t>(i) * 8);
      out.get(old);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += old[j];
    }
    out.set(acc);
    out.store(gx + static_cast<size_t>(i) * 8);
  }
}

int upsample_nhwc_fwd(int dtype, const void* x, void* y, int N, int h, int w, int H, int W, int C, cudaStream_t s) {
  PP_REQUIRE(C % 8 == 0, "upsample_nhwc_fwd: C=%d must be a multiple of 8", C);
  PP_REQUIRE_INT32(static_cast<long long>(N) * H * W * C, "upsample_nhwc_fwd");
  const long long total = static_cast<long long>(N) * H * W * (C / 8);
  PP_DISPATCH_T(dtype, upsample_nhwc_fwd_kernel<T><<<grid_for(total, 256), 256, 0, s>>>(
                           static_cast<const T*>(x), static_cast<T*>(y), N, h, w, H, W, C, ac_scale(h, H),
                           ac_scale(w, W)););
  PP_LAUNCH_CHECK();
  return PP_OK;
}
int upsample_nhwc_bwd(int dtype, const void* gy, void* gx, int N, int h, int w, int H, int W, int C, int accumulate,
                      cudaStream_t s) {
  PP_REQUIRE(C % 8 == 0, "upsample_nhwc_bwd: C=%d must be a multiple of 8", C);
  PP_REQUIRE_INT32(static_cast<long long>(N) * H * W * C, "upsample_nhwc_bwd");
  const long long total = static_cast<long long>(N) * h * w * (C / 8);
  PP_DISPATCH_T(dtype, upsample_nhwc_bwd_kernel<T><<<grid_for(total, 256), 256, 0, s>>>(
                           static_cast<const T*>(gy), static_cast<T*>(gx), N, h, w, H, W, C, ac_scale(h, H),
                           ac_scale(w, W), accumulate););
  PP_LAUNCH_CHECK();
  return PP_OK;
}

// NCHW fp32 planes (the aux-path logits, C = num_classes): x [NC][h][w] -> y [NC][H][W]
__global__ void __launch_bounds__(256) upsample_planes_fwd_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                                  long long NC, int h, int w, int H, int W, float sh,
                                                                  float sw) {
  const long long total = NC * H * W;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int X = int(i % W), Y = int((i / W) % H);
    const long long pl = i / (static_cast<long long>(W) * H);
    const Lerp ly = lerp_src(Y, h, sh), lx = lerp_src(X, w, sw);
    const float* b = x + pl * h * w;
    y[i] = ly.w0 * (lx.w0 * b[ly.i0 * w + lx.i0] + lx.w1 * b[ly.i0 * w + lx.i1]) +
           ly.w1 * (lx.w0 * b[ly.i1 * w + lx.i0] + lx.w1 * b[ly.i1 * w + lx.i1]);
  }
}
__global__ void __launch_bounds__(128) upsample_planes_bwd_kernel(const float* __restrict__ gy, float* __restrict__ gx,
                                                                  long long NC, int h, int w, int H, int W, float sh,
                                                                  float sw) {
  // one warp per input element: lanes stride over the candidate output window
  const long long total = NC * h * w;
  const int lane = threadIdx.x & 31;
  const long long warp_id = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long i = warp_id; i < total; i += nwarps) {
    const int xi = int(i % w), yi = int((i / w) % h);
    const long long pl = i / (static_cast<long long>(w) * h);
    int ylo, yhi, xlo, xhi;
    lerp_range(yi, H, sh, &ylo, &yhi);
    lerp_range(xi, W, sw, &xlo, &xhi);
    const int nx = xhi - xlo + 1, cnt = (yhi - ylo + 1) * nx;
    float acc = 0.f;
    for (int k = lane; k < cnt; k += 32) {
      const int Y = ylo + k / nx, X = xlo + k % nx;
      const float ww = lerp_weight(Y, yi, h, sh) * lerp_weight(X, xi, w, sw);
      if (ww != 0.f) acc = fmaf(ww, gy[(pl * H + Y) * W + X], acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) gx[i] = acc;
  }
}

int upsample_planes_fwd(const float* x, float* y, long long NC, int h, int w, int H, int W, cudaStream_t s) {
  upsample_planes_fwd_kernel<<<grid_for(NC * H * W, 256), 256, 0, s>>>(x, y, NC, h, w, H, W, ac_scale(h, H),
                                                                       ac_scale(w, W));
  PP_LAUNCH_CHECK();
  return PP_OK;
}
int upsample_planes_bwd(const float* gy, float* gx, long long NC, int h, int w, int H, int W, cudaStream_t s) {
  upsample_planes_bwd_kernel<<<grid_for(NC * h * w * 32, 128), 128, 0, s>>>(gy, gx, NC, h, w, H, W, ac_scale(h, H),
                                                                            ac_scale(w, W));
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
