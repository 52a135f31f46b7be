// Standalone on-GPU check of the tcgen05 conv kernels against the CUDA-core kernels (same bf16
// inputs, fp32 accumulation) and of the CUDA-core kernels against a CPU loop at a tiny shape.
// Usage: test_conv_tc <case index | -1 = count>. One case per process so a faulting kernel cannot
// poison the others.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../pacingpseudo_b200/csrc/pp_common.cuh"

#include "../../pacingpseudo_b200/csrc/pp_ops.h"

namespace pp {
int init_device(int device);
const char* last_error();
}  // namespace pp

struct Case { int N, H, W, C0, C1, oc0, oc1, dil, acc; const char* name; };
static const Case kCases[] = {
    {2, 32, 32, 64, 0, 64, 0, 1, 0, "basic bk64 bn64"},
    {2, 32, 32, 32, 0, 32, 0, 1, 0, "bk32 bn32"},
    {3, 32, 32, 128, 0, 256, 0, 2, 0, "dil2 bn256"},
    {2, 32, 32, 512, 512, 512, 0, 1, 0, "concat 512+512"},
    {2, 16, 16, 64, 32, 32, 0, 1, 0, "concat 64+32 (bk32) 16x16"},
    {4, 8, 8, 64, 0, 128, 0, 4, 0, "8x8 box spans images, dil4"},
    {1, 256, 256, 32, 0, 32, 0, 1, 0, "full-res rows"},
    {1, 28, 28, 64, 0, 64, 0, 1, 0, "28x28 ragged"},
    {1, 224, 224, 32, 0, 64, 0, 1, 0, "224 ragged rows"},
    {2, 32, 32, 128, 0, 64, 32, 1, 1, "dgrad style two dests, accumulate"},
    {2, 64, 64, 256, 128, 128, 0, 1, 0, "concat 256+128 64x64"},
    {12, 32, 32, 512, 0, 512, 0, 4, 0, "bottleneck dil4 full batch"},
    {3, 7, 5, 32, 0, 32, 0, 1, 0, "tiny odd 7x5"},
    {2, 32, 32, 128, 64, 64, 0, 1, 0, "wgrad: wide source 0 + narrow source 1"},
    {1, 128, 128, 64, 0, 64, 0, 1, 0, "narrow wgrad 64->64 128x128"},
    {2, 64, 64, 32, 0, 64, 0, 2, 0, "narrow wgrad 32->64 dil2"},
    {2, 128, 128, 64, 0, 64, 0, 1, 0, "halo conv 64->64 128x128"},
    {1, 256, 256, 64, 32, 32, 0, 1, 0, "halo conv concat 64+32->32 256x256"},
    {1, 256, 256, 32, 0, 64, 32, 1, 1, "halo dgrad 32->(64+32) accumulate"},
    {2, 224, 224, 32, 0, 32, 0, 1, 0, "halo conv 224 ragged strip"},
    {1, 112, 112, 32, 0, 64, 0, 1, 0, "halo conv 112 (single partial strip)"},
    {3, 128, 128, 64, 0, 32, 0, 1, 0, "halo conv 64->32"},
};

static float bf16_round(float f) { return __bfloat162float(__float2bfloat16(f)); }
static float frand() { return float(rand()) / RAND_MAX * 2.f - 1.f; }

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 2; } } while (0)
#define PPCK(x) do { int r = (x); if (r != 0) { printf("pp error %d: %s\n", r, pp::last_error()); return 3; } } while (0)

static std::vector<__nv_bfloat16> rand_bf16(size_t n, float scale) {
  std::vector<__nv_bfloat16> v(n);
  for (size_t i = 0; i < n; ++i) v[i] = __float2bfloat16(frand() * scale);
  return v;
}

static int cpu_anchor() {
  // SIMT forward + wgrad vs CPU loops, tiny fp32 problem.
  const int N = 1, H = 5, W = 6, C0 = 16, C1 = 16, CO = 8, dil = 2;
  const int ct = C0 + C1;
  std::vector<float> x0(N * H * W * C0), x1(N * H * W * C1), w(9 * CO * ct), b(CO), y(N * H * W * CO), dw(9 * CO * ct);
  for (auto& v : x0) v = frand();
  for (auto& v : x1) v = frand();
  for (auto& v : w) v = frand();
  for (auto& v : b) v = frand();
  float *dx0, *dx1, *dwp, *db, *dyv, *ddw;
  CK(cudaMalloc(&dx0, x0.size() * 4)); CK(cudaMalloc(&dx1, x1.size() * 4)); CK(cudaMalloc(&dwp, w.size() * 4));
  CK(cudaMalloc(&db, b.size() * 4)); CK(cudaMalloc(&dyv, y.size() * 4)); CK(cudaMalloc(&ddw, dw.size() * 4));
  CK(cudaMemcpy(dx0, x0.data(), x0.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dx1, x1.data(), x1.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dwp, w.data(), w.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, b.data(), b.size() * 4, cudaMemcpyHostToDevice));
  PPCK(pp::conv3x3_simt(pp::PP_F32, dx0, C0, dx1, C1, dwp, db, dyv, CO, 0, nullptr, 0, 0, N, H, W, dil, 0));
  CK(cudaMemcpy(y.data(), dyv, y.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0;
  std::vector<float> yref(y.size());
  for (int yy = 0; yy < H; ++yy) for (int xx = 0; xx < W; ++xx) for (int co = 0; co < CO; ++co) {
    double s = b[co];
    for (int tap = 0; tap < 9; ++tap) {
      int sy = yy + (tap / 3 - 1) * dil, sx = xx + (tap % 3 - 1) * dil;
      if (sy < 0 || sy >= H || sx < 0 || sx >= W) continue;
      for (int c = 0; c < ct; ++c) {
        float xv = c < C0 ? x0[(sy * W + sx) * C0 + c] : x1[(sy * W + sx) * C1 + c - C0];
        s += double(xv) * w[(tap * CO + co) * ct + c];
      }
    }
    yref[(yy * W + xx) * CO + co] = float(s);
    maxerr = fmax(maxerr, fabs(s - y[(yy * W + xx) * CO + co]));
  }
  printf("anchor: simt fwd vs cpu max abs err %.3e\n", maxerr);
  if (maxerr > 1e-4) return 1;
  // wgrad: dy := y (device), dw vs cpu
  CK(cudaMemset(ddw, 0, dw.size() * 4));
  PPCK(pp::conv3x3_wgrad_simt(pp::PP_F32, dyv, CO, dx0, C0, dx1, C1, ddw, N, H, W, dil, 0));
  CK(cudaMemcpy(dw.data(), ddw, dw.size() * 4, cudaMemcpyDeviceToHost));
  maxerr = 0;
  for (int tap = 0; tap < 9; ++tap) for (int co = 0; co < CO; ++co) for (int c = 0; c < ct; ++c) {
    double s = 0;
    for (int yy = 0; yy < H; ++yy) for (int xx = 0; xx < W; ++xx) {
      int sy = yy + (tap / 3 - 1) * dil, sx = xx + (tap % 3 - 1) * dil;
      if (sy < 0 || sy >= H || sx < 0 || sx >= W) continue;
      float xv = c < C0 ? x0[(sy * W + sx) * C0 + c] : x1[(sy * W + sx) * C1 + c - C0];
      s += double(y[(yy * W + xx) * CO + co]) * xv;
    }
    maxerr = fmax(maxerr, fabs(s - dw[(tap * CO + co) * ct + c]));
  }
  printf("anchor: simt wgrad vs cpu max abs err %.3e\n", maxerr);
  return maxerr > 1e-3 ? 1 : 0;
}

int main(int argc, char** argv) {
  const int ncases = int(sizeof(kCases) / sizeof(kCases[0]));
  int idx = argc > 1 ? atoi(argv[1]) : -1;
  if (idx < 0) { printf("%d\n", ncases + 1); return 0; }
  PPCK(pp::init_device(0));
  if (idx == ncases) {
    int r = cpu_anchor();
    printf("%s case %d (cpu anchor)\n", r ? "FAIL" : "PASS", idx);
    return r;
  }
  const Case c = kCases[idx];
  srand(1234 + idx);
  const size_t P = size_t(c.N) * c.H * c.W;
  const int ct = c.C0 + c.C1, co = c.oc0 + c.oc1;
  auto hx0 = rand_bf16(P * c.C0, 1.f);
  auto hx1 = rand_bf16(P * (c.C1 ? c.C1 : 1), 1.f);
  auto hw = rand_bf16(size_t(9) * co * ct, 1.f / sqrtf(9.f * ct));
  std::vector<float> hb(co);
  for (auto& v : hb) v = frand();
  auto hinit0 = rand_bf16(P * c.oc0, 1.f);
  auto hinit1 = rand_bf16(P * (c.oc1 ? c.oc1 : 1), 1.f);

  __nv_bfloat16 *x0, *x1, *w, *o0, *o1, *r0, *r1;
  float* b;
  CK(cudaMalloc(&x0, hx0.size() * 2)); CK(cudaMalloc(&x1, hx1.size() * 2)); CK(cudaMalloc(&w, hw.size() * 2));
  CK(cudaMalloc(&o0, hinit0.size() * 2)); CK(cudaMalloc(&o1, hinit1.size() * 2));
  CK(cudaMalloc(&r0, hinit0.size() * 2)); CK(cudaMalloc(&r1, hinit1.size() * 2));
  CK(cudaMalloc(&b, co * 4));
  CK(cudaMemcpy(x0, hx0.data(), hx0.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(x1, hx1.data(), hx1.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(w, hw.data(), hw.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(b, hb.data(), co * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(o0, hinit0.data(), hinit0.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(r0, hinit0.data(), hinit0.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(o1, hinit1.data(), hinit1.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(r1, hinit1.data(), hinit1.size() * 2, cudaMemcpyHostToDevice));

  const void* x1p = c.C1 ? x1 : nullptr;
  void* o1p = c.oc1 ? o1 : nullptr;
  void* r1p = c.oc1 ? r1 : nullptr;
  int fail = 0;

  // ---- forward / dgrad kernel ----
  PPCK(pp::conv3x3_simt(pp::PP_BF16, x0, c.C0, x1p, c.C1, w, b, r0, c.oc0, c.acc, r1p, c.oc1, 0, c.N, c.H, c.W, c.dil, 0));
  CK(cudaDeviceSynchronize());
  PPCK(pp::conv3x3_tc(x0, c.C0, x1p, c.C1, w, b, o0, c.oc0, c.acc, o1p, c.oc1, 0, c.N, c.H, c.W, c.dil, 0));
  CK(cudaDeviceSynchronize());
  for (int d = 0; d < (c.oc1 ? 2 : 1); ++d) {
    const size_t n = P * (d ? c.oc1 : c.oc0);
    std::vector<__nv_bfloat16> a(n), r(n);
    CK(cudaMemcpy(a.data(), d ? o1 : o0, n * 2, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(r.data(), d ? r1 : r0, n * 2, cudaMemcpyDeviceToHost));
    double maxdiff = 0, maxref = 0, sumsq = 0, refsq = 0;
    size_t bad = 0;
    for (size_t i = 0; i < n; ++i) {
      double av = __bfloat162float(a[i]), rv = __bfloat162float(r[i]);
      double df = fabs(av - rv);
      maxdiff = fmax(maxdiff, df); maxref = fmax(maxref, fabs(rv));
      sumsq += df * df; refsq += rv * rv;
      if (df > 0.02 * fmax(1.0, fabs(rv))) ++bad;
    }
    const double rel = sqrt(sumsq / fmax(refsq, 1e-30));
    printf("  fwd dest%d: max|diff| %.4e  max|ref| %.4e  rel-l2 %.4e  bad %zu/%zu\n", d, maxdiff, maxref, rel, bad, n);
    if (!(rel < 5e-3) || bad) fail = 1;
  }

  // ---- wgrad: dy := r0 contents re-used as gradient (oc0 channels) ----
  {
    const size_t nw = size_t(9) * c.oc0 * ct;
    float *dw_tc, *dw_ref;
    CK(cudaMalloc(&dw_tc, nw * 4)); CK(cudaMalloc(&dw_ref, nw * 4));
    CK(cudaMemset(dw_tc, 0, nw * 4)); CK(cudaMemset(dw_ref, 0, nw * 4));
    PPCK(pp::conv3x3_wgrad_simt(pp::PP_BF16, r0, c.oc0, x0, c.C0, x1p, c.C1, dw_ref, c.N, c.H, c.W, c.dil, 0));
    CK(cudaDeviceSynchronize());
    PPCK(pp::conv3x3_wgrad_tc(r0, c.oc0, x0, c.C0, x1p, c.C1, dw_tc, nullptr, c.N, c.H, c.W, c.dil, 0));
    CK(cudaDeviceSynchronize());
    std::vector<float> a(nw), r(nw);
    CK(cudaMemcpy(a.data(), dw_tc, nw * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(r.data(), dw_ref, nw * 4, cudaMemcpyDeviceToHost));
    double maxdiff = 0, maxref = 0, sumsq = 0, refsq = 0;
    for (size_t i = 0; i < nw; ++i) {
      double df = fabs(double(a[i]) - r[i]);
      maxdiff = fmax(maxdiff, df); maxref = fmax(maxref, fabs(double(r[i])));
      sumsq += df * df; refsq += double(r[i]) * r[i];
    }
    const double rel = sqrt(sumsq / fmax(refsq, 1e-30));
    printf("  wgrad: max|diff| %.4e  max|ref| %.4e  rel-l2 %.4e\n", maxdiff, maxref, rel);
    if (!(rel < 1e-3)) fail = 1;
  }
  printf("%s case %d (%s) N=%d H=%d W=%d C0=%d C1=%d oc0=%d oc1=%d dil=%d acc=%d\n", fail ? "FAIL" : "PASS", idx, c.name,
         c.N, c.H, c.W, c.C0, c.C1, c.oc0, c.oc1, c.dil, c.acc);
  return fail;
}
