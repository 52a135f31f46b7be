// common.cu — library state, error buffer, TMA tensor-map encoding.
#include <stdarg.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <vector>

#include "pp_common.cuh"

namespace pp {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

static std::mutex g_mu;
static int g_sm_count = 0;
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

int sm_count() { return g_sm_count > 0 ? g_sm_count : 148; }

static std::atomic<long long> g_launches{0};
void count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
long long launch_count() { return g_launches.load(std::memory_order_relaxed); }

struct ProfRec { cudaEvent_t a, b; double flops; int family; };
static std::vector<ProfRec> g_prof;
static size_t g_prof_used = 0;
static std::atomic<int> g_prof_on{0};
static std::mutex g_prof_mu;

void prof_enable(int on) { g_prof_on.store(on); }
bool prof_enabled() { return g_prof_on.load() != 0; }
int prof_begin(int family, double flops, cudaStream_t s) {
  if (!g_prof_on.load()) return -1;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (g_prof_used == g_prof.size()) {
    ProfRec r{};
    if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return -1;
    g_prof.push_back(r);
  }
  ProfRec& r = g_prof[g_prof_used];
  r.flops = flops;
  r.family = family;
  cudaEventRecord(r.a, s);
  return static_cast<int>(g_prof_used++);
}
void prof_end(int slot, cudaStream_t s) {
  if (slot < 0) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  cudaEventRecord(g_prof[slot].b, s);
}
// Time during which at least one launch of `family` (or of any family, family < 0) was running: the union of the
// [start, end] intervals on the device timeline. With the weight-gradient kernels on a side stream, launches overlap,
// and a plain sum of durations would count the shared time twice.
int prof_collect(int family, double* ms, double* flops, long long* launches) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  *ms = 0; *flops = 0; *launches = 0;
  if (g_prof_used == 0) return PP_OK;
  std::vector<std::pair<float, float>> iv;
  for (size_t i = 0; i < g_prof_used; ++i) {
    if (family >= 0 ? g_prof[i].family != family : g_prof[i].family > PROF_WGRAD) continue;   // < 0: both conv families
    PP_CHECK_CUDA(cudaEventSynchronize(g_prof[i].b));
    float ta = 0.f, tb = 0.f;
    if (i > 0) PP_CHECK_CUDA(cudaEventElapsedTime(&ta, g_prof[0].a, g_prof[i].a));
    PP_CHECK_CUDA(cudaEventElapsedTime(&tb, g_prof[0].a, g_prof[i].b));
    iv.emplace_back(ta, tb);
    *flops += g_prof[i].flops; *launches += 1;
  }
  std::sort(iv.begin(), iv.end());
  float cur_a = 0.f, cur_b = -1.f;
  for (const auto& p : iv) {
    if (cur_b < cur_a || p.first > cur_b) {
      if (cur_b >= cur_a) *ms += cur_b - cur_a;
      cur_a = p.first; cur_b = p.second;
    } else if (p.second > cur_b) {
      cur_b = p.second;
    }
  }
  if (cur_b >= cur_a) *ms += cur_b - cur_a;
  return PP_OK;
}
void prof_reset() {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof_used = 0;
}

int init_device(int device) {
  std::lock_guard<std::mutex> lk(g_mu);
  PP_CHECK_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  PP_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("pacingpseudo_b200 requires an sm_100 (Blackwell B200) device, found sm_%d%d (%s)", prop.major,
              prop.minor, prop.name);
    return PP_ERR_NOT_INIT;
  }
  g_sm_count = prop.multiProcessorCount;
  if (g_encode == nullptr) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    PP_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (qres != cudaDriverEntryPointSuccess || fn == nullptr) {
      set_error("cuTensorMapEncodeTiled entry point not available (query result %d)", int(qres));
      return PP_ERR_NOT_INIT;
    }
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  return PP_OK;
}

static int encode(CUtensorMap* out, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                  const cuuint32_t* box, bool swizzle128) {
  if (g_encode == nullptr) {
    set_error("pp_init() has not been called on this process");
    return PP_ERR_NOT_INIT;
  }
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu %llu %llu, box %u %u %u, base %p)",
              int(r), rank, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2],
              box[0], box[1], box[2], base);
    return PP_ERR_CUDA;
  }
  return PP_OK;
}

int encode_tmap_nhwc(CUtensorMap* out, const void* base, int N, int H, int W, int C, int box_c, int bw, int bh,
                     int bn, bool swizzle128) {
  const cuuint64_t dims[4] = {cuuint64_t(C), cuuint64_t(W), cuuint64_t(H), cuuint64_t(N)};
  const cuuint64_t strides[3] = {cuuint64_t(C) * 2, cuuint64_t(W) * C * 2, cuuint64_t(H) * W * C * 2};
  const cuuint32_t box[4] = {cuuint32_t(box_c), cuuint32_t(bw), cuuint32_t(bh), cuuint32_t(bn)};
  return encode(out, base, 4, dims, strides, box, swizzle128);
}

int encode_tmap_weights(CUtensorMap* out, const void* base, int taps, int Cout, int K, int box_k, int box_n,
                        bool swizzle128) {
  const cuuint64_t dims[3] = {cuuint64_t(K), cuuint64_t(Cout), cuuint64_t(taps)};
  const cuuint64_t strides[2] = {cuuint64_t(K) * 2, cuuint64_t(Cout) * K * 2};
  const cuuint32_t box[3] = {cuuint32_t(box_k), cuuint32_t(box_n), 1};
  return encode(out, base, 3, dims, strides, box, swizzle128);
}

}  // namespace pp
