// Hardware experiment (B200): can a UMMA K-major swizzled shared-memory descriptor start at a row offset that is not a
// multiple of the swizzle atom (8 rows)? A [rows x BK] tile is loaded ONCE by TMA; the MMA then reads the 128-row window
// starting at row `shift`. This is what a shared-memory-resident im2col needs (the +-1 pixel taps of a 3x3 conv are
// row shifts of one halo tile). Two descriptor variants are tried: base_offset = 0 and base_offset = (addr >> 7) & 7.
// Usage: exp_desc_shift  -> prints max |error| per (swizzle, shift, variant).
#include <math.h>
#include <stdlib.h>

#include <vector>

#include "../../pacingpseudo_b200/csrc/pp_common.cuh"

namespace pp {
int init_device(int device);
const char* last_error();
}  // namespace pp

template <int BK>
__global__ void __launch_bounds__(128, 1)
shift_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* out, int shift,
             int variant, int rows) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int ROW = BK * 2;
  uint8_t* sa = smem;                       // rows x ROW bytes
  uint8_t* sb = smem + 160 * 128;           // 32 x ROW bytes (1024-aligned)
  __shared__ uint64_t bar_load, bar_mma;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    pp::mbar_init(&bar_load, 1);
    pp::mbar_init(&bar_mma, 1);
    pp::fence_mbar_init();
  }
  if (warp == 0) pp::tmem_alloc<32>(&tmem_slot);
  pp::tc_fence_before();
  __syncthreads();
  pp::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    pp::mbar_arrive_expect_tx(&bar_load, rows * ROW + 32 * ROW);
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(pp::smem_u32(sa)), "l"(reinterpret_cast<uint64_t>(&tmA)), "r"(pp::smem_u32(&bar_load)), "r"(0), "r"(0) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(pp::smem_u32(sb)), "l"(reinterpret_cast<uint64_t>(&tmB)), "r"(pp::smem_u32(&bar_load)), "r"(0), "r"(0) : "memory");
    pp::mbar_wait(&bar_load, 0);
    pp::tc_fence_after();
    constexpr uint32_t SWZ = BK == 64 ? pp::SWZ_128B : pp::SWZ_64B;
    const uint32_t idesc = pp::make_idesc_bf16(128, 32, 0, 0);
    const uint32_t a0 = pp::smem_u32(sa) + shift * ROW;
    for (int k = 0; k < BK / 16; ++k) {
      uint64_t da = pp::make_smem_desc(a0 + k * 32, 16, 8 * ROW, SWZ);
      if (variant == 1) da |= uint64_t((a0 >> 7) & 7) << 49;
      const uint64_t db = pp::make_smem_desc(pp::smem_u32(sb) + k * 32, 16, 8 * ROW, SWZ);
      pp::umma_bf16(tmem, da, db, idesc, k != 0);
    }
    pp::umma_commit(&bar_mma);
  }
  __syncwarp();
  pp::mbar_wait(&bar_mma, 0);
  pp::tc_fence_after();
  uint32_t v[32];
  pp::tmem_ld_32x32(tmem + (uint32_t(warp * 32) << 16), v);
  pp::tmem_wait_ld();
  for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 32 + j] = __uint_as_float(v[j]);
  pp::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    pp::tc_fence_after();
    pp::tmem_dealloc<32>(tmem);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int BK>
static int run(EncodeTiledFn enc) {
  const int rows = 144;
  std::vector<__nv_bfloat16> hA(rows * BK), hB(32 * BK);
  for (auto& x : hA) x = __float2bfloat16(float(rand()) / RAND_MAX - 0.5f);
  for (auto& x : hB) x = __float2bfloat16(float(rand()) / RAND_MAX - 0.5f);
  __nv_bfloat16 *dA, *dB;
  float* dOut;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dOut, 128 * 32 * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap tA, tB;
  cuuint32_t es[2] = {1, 1};
  const CUtensorMapSwizzle sw = BK == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  {
    cuuint64_t dims[2] = {cuuint64_t(BK), cuuint64_t(rows)}, str[1] = {cuuint64_t(BK) * 2};
    cuuint32_t box[2] = {cuuint32_t(BK), cuuint32_t(rows)};
    if (enc(&tA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return 1;
  }
  {
    cuuint64_t dims[2] = {cuuint64_t(BK), 32}, str[1] = {cuuint64_t(BK) * 2};
    cuuint32_t box[2] = {cuuint32_t(BK), 32};
    if (enc(&tB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return 1;
  }
  cudaFuncSetAttribute(shift_kernel<BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  std::vector<float> out(128 * 32);
  for (int variant = 0; variant < 2; ++variant) {
    for (int shift = 0; shift <= 12; ++shift) {
      cudaMemset(dOut, 0, 128 * 32 * 4);
      shift_kernel<BK><<<1, 128, 48 * 1024, 0>>>(tA, tB, dOut, shift, variant, rows);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("BK=%d shift=%d variant=%d: CUDA error %s\n", BK, shift, variant, cudaGetErrorString(e)); return 2; }
      cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost);
      double maxerr = 0;
      for (int r = 0; r < 128; ++r)
        for (int n = 0; n < 32; ++n) {
          double s = 0;
          for (int k = 0; k < BK; ++k)
            s += double(__bfloat162float(hA[(r + shift) * BK + k])) * __bfloat162float(hB[n * BK + k]);
          maxerr = fmax(maxerr, fabs(s - out[r * 32 + n]));
        }
      printf("swizzle %3dB shift %2d rows, base_offset %s: max |err| %.3e %s\n", BK * 2, shift,
             variant ? "(addr>>7)&7" : "0          ", maxerr, maxerr < 1e-3 ? "OK" : "WRONG");
    }
  }
  return 0;
}

int main() {
  if (pp::init_device(0)) { printf("init failed: %s\n", pp::last_error()); return 1; }
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fn);
  int rc = run<64>(enc);
  rc |= run<32>(enc);
  return rc;
}
