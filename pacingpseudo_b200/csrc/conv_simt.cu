// conv_simt.cu — CUDA-core implicit-GEMM 3x3 convolution (forward / dgrad / wgrad), templated on
// the activation type. Two jobs:
//   * the fp32 precision mode of the library (BASELINE north_star: "1e-4 in an fp32 mode"), and
//   * the on-device cross-check of the tcgen05 kernels in tests (same bf16 inputs, fp32 accumulate).
// Semantics are identical to conv_tc.cu (two concat sources, two scatter destinations, dilation).
#include "pp_common.cuh"

namespace pp {

template <typename T>
__global__ void __launch_bounds__(256)
conv3x3_simt_kernel(const T* __restrict__ x0, int C0, const T* __restrict__ x1, int C1,
                    const T* __restrict__ wpack, const float* __restrict__ bias, T* out0, int outc0, int acc0,
                    T* out1, int outc1, int acc1, int N, int H, int W, int dil) {
  constexpr int TM = 64, TN = 64, TK = 16;
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  __shared__ int pn[TM], py[TM], pxx[TM];

  const int tid = threadIdx.x;
  const int ctot = C0 + C1;
  const int cout = outc0 + outc1;
  const long long P = static_cast<long long>(N) * H * W;
  const long long p0 = static_cast<long long>(blockIdx.x) * TM;
  const int col0 = blockIdx.y * TN;

  if (tid < TM) {
    long long p = p0 + tid;
    if (p < P) {
      pxx[tid] = int(p % W);
      py[tid] = int((p / W) % H);
      pn[tid] = int(p / (static_cast<long long>(W) * H));
    } else {
      pn[tid] = -1; py[tid] = 0; pxx[tid] = 0;
    }
  }
  __syncthreads();

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int ty = tid / 16, tx = tid % 16;
  const int lrow = tid / 4, lk = (tid % 4) * 4;  // loader mapping: row (pixel / cout), 4 consecutive k

  for (int tap = 0; tap < 9; ++tap) {
    const int oy = (tap / 3 - 1) * dil, ox = (tap % 3 - 1) * dil;
    for (int k0 = 0; k0 < ctot; k0 += TK) {
      // A tile: 64 pixels x 16 channels of the (virtually concatenated) input at the tap offset
      {
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        const int n = pn[lrow];
        const int yy = py[lrow] + oy, xx = pxx[lrow] + ox;
        if (n >= 0 && yy >= 0 && yy < H && xx >= 0 && xx < W) {
          const int c = k0 + lk;
          const long long pix = (static_cast<long long>(n) * H + yy) * W + xx;
          const T* src = (c < C0) ? (x0 + pix * C0 + c) : (x1 + pix * C1 + (c - C0));
#pragma unroll
          for (int j = 0; j < 4; ++j) v[j] = to_f32(src[j]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) As[lk + j][lrow] = v[j];
      }
      // B tile: 64 output channels x 16 k
      {
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        const int co = col0 + lrow;
        if (co < cout) {
          const T* src = wpack + (static_cast<long long>(tap) * cout + co) * ctot + k0 + lk;
#pragma unroll
          for (int j = 0; j < 4; ++j) v[j] = to_f32(src[j]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) Bs[lk + j][lrow] = v[j];
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < TK; ++k) {
        float a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long p = p0 + ty * 4 + i;
    if (p >= P) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = col0 + tx * 4 + j;
      if (col >= cout) continue;
      float v = acc[i][j];
      if (bias != nullptr) v += bias[col];
      T* o;
      int a;
      if (col < outc0) { o = out0 + p * outc0 + col; a = acc0; }
      else             { o = out1 + p * outc1 + (col - outc0); a = acc1; }
      if (a) v += to_f32(*o);
      *o = from_f32<T>(v);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
conv3x3_wgrad_simt_kernel(const T* __restrict__ dy, int Cout, const T* __restrict__ x0, int C0,
                          const T* __restrict__ x1, int C1, float* __restrict__ dw, int N, int H, int W, int dil,
                          long long pix_per_split) {
  constexpr int TM = 64, TN = 64, TK = 16;
  __shared__ float As[TK][TM + 4];  // [pixel][co]
  __shared__ float Bs[TK][TN + 4];  // [pixel][ci]
  const int tid = threadIdx.x;
  const int ctot = C0 + C1;
  const int ci_tiles = (ctot + TN - 1) / TN;
  const int co0 = (blockIdx.x / ci_tiles) * TM;
  const int ci0 = (blockIdx.x % ci_tiles) * TN;
  const int tap = blockIdx.y;
  const int oy = (tap / 3 - 1) * dil, ox = (tap % 3 - 1) * dil;
  const long long P = static_cast<long long>(N) * H * W;
  const long long pbeg = static_cast<long long>(blockIdx.z) * pix_per_split;
  const long long pend = (pbeg + pix_per_split < P) ? pbeg + pix_per_split : P;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int ty = tid / 16, tx = tid % 16;
  const int lp = tid / 16, lc = (tid % 16) * 4;  // loader: pixel in chunk, 4 consecutive channels

  for (long long pc = pbeg; pc < pend; pc += TK) {
    const long long p = pc + lp;
    float va[4] = {0.f, 0.f, 0.f, 0.f}, vb[4] = {0.f, 0.f, 0.f, 0.f};
    if (p < pend) {
      const int px = int(p % W), pyy = int((p / W) % H), n = int(p / (static_cast<long long>(W) * H));
      if (co0 + lc < Cout) {
        const T* s = dy + p * Cout + co0 + lc;
#pragma unroll
        for (int j = 0; j < 4; ++j) va[j] = to_f32(s[j]);
      }
      const int yy = pyy + oy, xx = px + ox;
      const int c = ci0 + lc;
      if (c < ctot && yy >= 0 && yy < H && xx >= 0 && xx < W) {
        const long long q = (static_cast<long long>(n) * H + yy) * W + xx;
        const T* s = (c < C0) ? (x0 + q * C0 + c) : (x1 + q * C1 + (c - C0));
#pragma unroll
        for (int j = 0; j < 4; ++j) vb[j] = to_f32(s[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) { As[lp][lc + j] = va[j]; Bs[lp][lc + j] = vb[j]; }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = co0 + ty * 4 + i;
    if (co >= Cout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ci = ci0 + tx * 4 + j;
      if (ci < ctot) atomicAdd(dw + (static_cast<long long>(tap) * Cout + co) * ctot + ci, acc[i][j]);
    }
  }
}

template <typename T>
static int conv3x3_simt_t(const void* x0, int C0, const void* x1, int C1, const void* wpack, const float* bias,
                          void* out0, int outc0, int acc0, void* out1, int outc1, int acc1, int N, int H, int W,
                          int dil, cudaStream_t stream) {
  PP_REQUIRE(C0 % 16 == 0 && C1 % 16 == 0 && C0 > 0, "conv3x3_simt: input channels must be multiples of 16");
  PP_REQUIRE((x1 == nullptr) == (C1 == 0) && (out1 == nullptr) == (outc1 == 0), "conv3x3_simt: pointer/channel mismatch");
  const long long P = static_cast<long long>(N) * H * W;
  dim3 grid(static_cast<unsigned>(ceil_div_ll(P, 64)), ceil_div(outc0 + outc1, 64));
  conv3x3_simt_kernel<T><<<grid, 256, 0, stream>>>(static_cast<const T*>(x0), C0, static_cast<const T*>(x1), C1,
                                                   static_cast<const T*>(wpack), bias, static_cast<T*>(out0), outc0,
                                                   acc0, static_cast<T*>(out1), outc1, acc1, N, H, W, dil);
  PP_LAUNCH_CHECK();
  return PP_OK;
}

int conv3x3_simt(int dtype, const void* x0, int C0, const void* x1, int C1, const void* wpack, const float* bias,
                 void* out0, int outc0, int acc0, void* out1, int outc1, int acc1, int N, int H, int W, int dil,
                 cudaStream_t stream) {
  if (dtype == PP_F32)
    return conv3x3_simt_t<float>(x0, C0, x1, C1, wpack, bias, out0, outc0, acc0, out1, outc1, acc1, N, H, W, dil, stream);
  return conv3x3_simt_t<__nv_bfloat16>(x0, C0, x1, C1, wpack, bias, out0, outc0, acc0, out1, outc1, acc1, N, H, W, dil,
                                       stream);
}

template <typename T>
static int conv3x3_wgrad_simt_t(const void* dy, int Cout, const void* x0, int C0, const void* x1, int C1, float* dw,
                                int N, int H, int W, int dil, cudaStream_t stream) {
  PP_REQUIRE(Cout % 4 == 0 && C0 % 4 == 0 && C1 % 4 == 0 && C0 > 0, "conv3x3_wgrad_simt: channels must be multiples of 4");
  const long long P = static_cast<long long>(N) * H * W;
  const int ctot = C0 + C1;
  const int base = ceil_div(Cout, 64) * ceil_div(ctot, 64) * 9;
  long long splits = ceil_div(4 * sm_count(), base);
  if (splits < 1) splits = 1;
  long long pps = ceil_div_ll(P, splits);
  pps = ceil_div_ll(pps, 16) * 16;
  splits = ceil_div_ll(P, pps);
  dim3 grid(ceil_div(Cout, 64) * ceil_div(ctot, 64), 9, static_cast<unsigned>(splits));
  conv3x3_wgrad_simt_kernel<T><<<grid, 256, 0, stream>>>(static_cast<const T*>(dy), Cout, static_cast<const T*>(x0), C0,
                                                         static_cast<const T*>(x1), C1, dw, N, H, W, dil, pps);
  PP_LAUNCH_CHECK();
  return PP_OK;
}

int conv3x3_wgrad_simt(int dtype, const void* dy, int Cout, const void* x0, int C0, const void* x1, int C1, float* dw,
                       int N, int H, int W, int dil, cudaStream_t stream) {
  if (dtype == PP_F32) return conv3x3_wgrad_simt_t<float>(dy, Cout, x0, C0, x1, C1, dw, N, H, W, dil, stream);
  return conv3x3_wgrad_simt_t<__nv_bfloat16>(dy, Cout, x0, C0, x1, C1, dw, N, H, W, dil, stream);
}

}  // namespace pp
