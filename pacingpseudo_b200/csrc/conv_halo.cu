// conv_halo.cu — 3x3 convolution for the NARROW, HIGH-RESOLUTION layers (32..96 channels at 256^2 / 128^2):
// shared-memory-resident im2col on tcgen05.
//
// The generic kernel (conv_tc.cu) re-streams the 128-pixel A tile from L2 once per tap (9x) and the weight tile once
// per M tile; for the narrow layers that L2->SM traffic, not the tensor core, is the bound (ncu: tensor pipe 10 %,
// ~2.5 us per tile). Here one CTA owns a strip of 128 output columns x L output rows of one image:
//   * the packed weights of ALL 9 taps stay resident in shared memory (9 * Cout * Cin * 2 B <= ~72 KB);
//   * each INPUT row segment (128 + 2 halo pixels, all channels) is loaded exactly once by TMA into a 4..8-slot ring;
//   * the nine taps of an output row are nine views of three ring slots: the +-1 pixel shifts are shared-memory
//     descriptor start offsets of +-1 row (128 B / 64 B). A B200 experiment (tests/cuda/exp_desc_shift.cu,
//     profiles/r01_exp_umma_descriptor_row_shift.txt) shows the UMMA swizzle is a function of the absolute smem
//     address, so a K-major swizzled descriptor may start at ANY row with base_offset = 0;
//   * two TMEM accumulators: the epilogue of row y (eight warps, two per TMEM lane quarter) overlaps the MMAs of
//     row y+1.
// HBM traffic is the algorithmic minimum (input read once, output written once).
//
// Same semantics as conv3x3_tc (two concat sources, two dgrad destinations with accumulate flags, bias, fused
// BatchNorm statistics). Reference call site: nn.Conv2d in /root/reference/models/unet.py:188.
#include "pp_common.cuh"
#include "pp_ops.h"

namespace pp {

static constexpr int kHaloThreads = 320;   // TMA producer + MMA issuer + 8 epilogue warps (two per TMEM lane quarter)
static constexpr int kMaxRing = 8;   // input-row ring slots: 3 live rows + up to 5 rows of TMA prefetch in flight
static constexpr int kTW = 128;      // output columns per strip == MMA M
static constexpr int kBoxW = kTW + 2;

struct HaloParams {
  int N, H, W;
  int strips, segs, L;         // column strips per row, row segments per image, rows per segment
  int kc0, kc1, c0;            // K chunks per source, channels of source 0
  int ctot;
  __nv_bfloat16* out0;
  __nv_bfloat16* out1;
  int outc0, outc1, acc0, acc1;
  const float* bias;
  double* stats;
  int imgs_per_group, groups;
  int chunk_bytes;             // bytes of one ring chunk (kBoxW rows, rounded up to 1 KB)
  int ring;                    // ring slots in use (4..kMaxRing)
  const float* ep_scale;       // optional eval-mode BatchNorm + LeakyReLU epilogue: out = lrelu(acc * scale + shift)
  const float* ep_shift;
  float ep_slope;
};

template <int BLOCK_N, int BK>
__global__ void __launch_bounds__(kHaloThreads, BLOCK_N == 32 ? 2 : 1)
conv3x3_halo_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                       const __grid_constant__ CUtensorMap tmB, const HaloParams p) {
  constexpr int ROW = BK * 2;
  constexpr uint32_t SWZ = (BK == 64) ? SWZ_128B : SWZ_64B;
  constexpr uint32_t SBO = 8 * ROW;
  constexpr int WTILE = BLOCK_N * ROW;                      // one (tap, chunk) weight tile
  constexpr int TMEM_BUF = BLOCK_N <= 32 ? 32 : (BLOCK_N <= 64 ? 64 : 128);
  constexpr int TMEM_COLS = 2 * TMEM_BUF;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int chunks = p.kc0 + p.kc1;
  uint8_t* s_w = smem;                                      // [9][chunks][BLOCK_N][BK]
  uint8_t* s_ring = s_w + 9 * chunks * WTILE;               // [ring][chunks][chunk_bytes]
  const int slot_bytes = chunks * p.chunk_bytes;
  const int kRing = p.ring;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_ring + kRing * slot_bytes);
  uint64_t* w_full = bars;
  uint64_t* row_full = bars + 1;
  uint64_t* row_empty = row_full + kMaxRing;
  uint64_t* acc_full = row_empty + kMaxRing;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* s_stats = reinterpret_cast<float*>(tmem_slot + 2);  // [4 warps][2][BLOCK_N]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // job = (image, column strip, row segment)
  const int job = blockIdx.x;
  const int seg = job % p.segs;
  const int strip = (job / p.segs) % p.strips;
  const int img = job / (p.segs * p.strips);
  const int x0 = strip * kTW;
  const int r0 = seg * p.L;
  const int rows = min(p.L, p.H - r0);                       // output rows of this job

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    if (p.kc1 > 0) tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmB);
    mbar_init(w_full, 1);
    for (int s = 0; s < kRing; ++s) { mbar_init(&row_full[s], 1); mbar_init(&row_empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 8); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = warp_uniform(*tmem_slot);

  if (warp == 0) {
    if (elect_one()) {
      // ===== TMA producer: resident weights, then one box per (input row, chunk) =====
      mbar_arrive_expect_tx(w_full, 9 * chunks * WTILE);
      for (int tap = 0; tap < 9; ++tap)
        for (int kc = 0; kc < chunks; ++kc) {
          const int kofs = kc < p.kc0 ? kc * BK : p.c0 + (kc - p.kc0) * BK;
          tma_load_3d(s_w + (tap * chunks + kc) * WTILE, &tmB, w_full, kofs, 0, tap);
        }
      for (int i = 0; i < rows + 2; ++i) {                   // input row r0 - 1 + i (out-of-range rows load as zeros)
        const int slot = i % kRing;
        mbar_wait(&row_empty[slot], ((i / kRing) & 1) ^ 1);
        mbar_arrive_expect_tx(&row_full[slot], chunks * kBoxW * ROW);
        uint8_t* dst = s_ring + slot * slot_bytes;
        for (int kc = 0; kc < chunks; ++kc) {
          if (kc < p.kc0) tma_load_4d(dst + kc * p.chunk_bytes, &tmA0, &row_full[slot], kc * BK, x0 - 1, r0 - 1 + i, img);
          else tma_load_4d(dst + kc * p.chunk_bytes, &tmA1, &row_full[slot], (kc - p.kc0) * BK, x0 - 1, r0 - 1 + i, img);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // ===== MMA issuer =====
      // These MMAs are short (N = 32..96: 16..48 tensor cycles each), so the issue loop itself must be lean:
      // descriptors advance by 32-bit adds on the uniform datapath (umma_bf16_lohi), the ring slot index is tracked
      // incrementally, and the TMEM address is provably warp-uniform (warp_uniform() above).
      constexpr uint32_t idesc = make_idesc_bf16(128, BLOCK_N, 0, 0);
      constexpr uint32_t dhi = smem_desc_hi(SBO, SWZ);
      constexpr uint32_t WT16 = WTILE >> 4;
      mbar_wait(w_full, 0);
      mbar_wait(&row_full[0], 0);
      mbar_wait(&row_full[1], 0);
      const uint32_t w_lo = smem_desc_lo(smem_u32(s_w), 16);
      const uint32_t ring_lo = smem_desc_lo(smem_u32(s_ring), 16);
      const uint32_t slot16 = static_cast<uint32_t>(slot_bytes) >> 4, chunk16 = static_cast<uint32_t>(p.chunk_bytes) >> 4;
      const uint32_t tap16 = static_cast<uint32_t>(chunks) * WT16;   // weight tiles of consecutive taps
      int s0 = 0;                                                     // ring slot of input row t (== output row t - 1)
      int snew = 2 % kRing;                                           // ring slot of the newest row needed (t + 2)
      uint32_t pnew = 0;
      for (int t = 0; t < rows; ++t) {
        mbar_wait(&row_full[snew], pnew);
        const int b = t & 1;
        mbar_wait(&acc_empty[b], ((t >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + b * TMEM_BUF;
        int sl = s0;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          uint32_t a_lo = ring_lo + static_cast<uint32_t>(sl) * slot16;
          uint32_t b_lo = w_lo + static_cast<uint32_t>(ky * 3) * tap16;
          for (int kc = 0; kc < chunks; ++kc) {
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
              for (int k = 0; k < BK / 16; ++k)     // +-1 pixel == +-1 smem row; K step of 16 channels == 32 bytes
                umma_bf16_lohi(d_tmem, a_lo + kx * (ROW >> 4) + k * 2, dhi, b_lo + kx * tap16 + k * 2, dhi, idesc,
                               (kx | k) != 0 ? 1u : ((ky | kc) != 0 ? 1u : 0u));
            }
            a_lo += chunk16;
            b_lo += WT16;
          }
          if (++sl == kRing) sl = 0;
        }
        umma_commit(&acc_full[b]);                 // accumulator of output row t complete
        umma_commit(&row_empty[s0]);               // input row t is not needed by later output rows
        if (++s0 == kRing) s0 = 0;
        if (++snew == kRing) { snew = 0; pnew ^= 1; }
      }
    }
  } else {
    // ===== epilogue: TMEM -> registers -> (+bias, +old) -> bf16 NHWC (+ BatchNorm partial sums) =====
    // The epilogue of a row is a serial instruction stream per warp (ncu: issue slots 40 % busy, tensor pipe 15 %) and
    // paces these narrow tiles, so EIGHT warps share it: two per TMEM lane quarter, each owning half of the columns in
    // units of 16. The bias lives in L1, the BatchNorm column sums are accumulated per thread across all rows of the
    // job and the lane transpose-reduce runs once per job instead of once per row.
    constexpr int UNITS = BLOCK_N / 32;        // 16-column units per warp (half of the tile's BLOCK_N / 16 units)
    constexpr bool kStatsOk = BLOCK_N <= 64;   // statistics are a forward-pass feature (Cout 32 / 64)
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int cbeg = half * (BLOCK_N / 2);     // first column of this warp
    const int px = x0 + q * 32 + lane;
    const bool col_ok = px < p.W;
    const bool bias_vec = (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0;
    float s1[kStatsOk ? UNITS * 16 : 1], s2[kStatsOk ? UNITS * 16 : 1];
    if constexpr (kStatsOk) {
#pragma unroll
      for (int j = 0; j < UNITS * 16; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
    }
    for (int t = 0; t < rows; ++t) {
      const int b = t & 1;
      mbar_wait(&acc_full[b], (t >> 1) & 1);
      tc_fence_after();
      const long long pix = (static_cast<long long>(img) * p.H + (r0 + t)) * p.W + px;
#pragma unroll
      for (int u = 0; u < UNITS; ++u) {
        const int c = cbeg + u * 16;
        __nv_bfloat16* dst;
        int dstc, acc, ch;
        if (c < p.outc0) { dst = p.out0; dstc = p.outc0; acc = p.acc0; ch = c; }
        else             { dst = p.out1; dstc = p.outc1; acc = p.acc1; ch = c - p.outc0; }
        uint32_t v[16];
        tmem_ld_32x16(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(b * TMEM_BUF + c), v);
        tmem_wait_ld();
        float f[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
        if (p.bias != nullptr) {
          if (bias_vec) {   // broadcast vector loads
            const float4* b4 = reinterpret_cast<const float4*>(p.bias + c);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 bv = __ldg(b4 + j);
              f[4 * j] += bv.x; f[4 * j + 1] += bv.y; f[4 * j + 2] += bv.z; f[4 * j + 3] += bv.w;
            }
          } else {          // any 4-byte aligned pointer is legal at the C ABI
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] += __ldg(p.bias + c + j);
          }
        }
        if (p.ep_scale != nullptr) {   // eval-mode BatchNorm + LeakyReLU on the fp32 accumulator
          const float4* s4 = reinterpret_cast<const float4*>(p.ep_scale + c);
          const float4* h4 = reinterpret_cast<const float4*>(p.ep_shift + c);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 sv = __ldg(s4 + j), hv = __ldg(h4 + j);
            f[4 * j] = lrelu(fmaf(f[4 * j], sv.x, hv.x), p.ep_slope);
            f[4 * j + 1] = lrelu(fmaf(f[4 * j + 1], sv.y, hv.y), p.ep_slope);
            f[4 * j + 2] = lrelu(fmaf(f[4 * j + 2], sv.z, hv.z), p.ep_slope);
            f[4 * j + 3] = lrelu(fmaf(f[4 * j + 3], sv.w, hv.w), p.ep_slope);
          }
        }
        if (col_ok) {
          __nv_bfloat16* o = dst + pix * dstc + ch;
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            Vec8<__nv_bfloat16> pk;
            float tt[8];
            if (acc) {
              pk.load(o + g * 8);
              pk.get(tt);
#pragma unroll
              for (int j = 0; j < 8; ++j) f[g * 8 + j] += tt[j];
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) tt[j] = f[g * 8 + j];
            pk.set(tt);
            pk.store(o + g * 8);
            if constexpr (kStatsOk) {
              pk.get(tt);   // statistics of the ROUNDED values (what BatchNorm reads back)
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                s1[u * 16 + g * 8 + j] += tt[j];
                s2[u * 16 + g * 8 + j] = fmaf(tt[j], tt[j], s2[u * 16 + g * 8 + j]);
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[b]);
    }
    if constexpr (kStatsOk) {
      if (p.stats != nullptr) {
#pragma unroll
        for (int u = 0; u < UNITS; ++u) {
          // column sums over this warp's 32 pixel columns: butterfly transpose-reduce of 16 values over lane bits
          // 8,4,2,1 (lane j & 15 ends with channel j & 15), then one plain exchange across lane bit 16
          float a[16], bq[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) { a[j] = s1[u * 16 + j]; bq[j] = s2[u * 16 + j]; }
#pragma unroll
          for (int w = 8; w >= 1; w >>= 1) {
            const bool hi = (lane & w) != 0;
#pragma unroll
            for (int j = 0; j < w; ++j) {
              const float a1 = hi ? a[j] : a[j + w], a2 = hi ? bq[j] : bq[j + w];
              const float k1 = hi ? a[j + w] : a[j], k2 = hi ? bq[j + w] : bq[j];
              a[j] = k1 + __shfl_xor_sync(0xffffffffu, a1, w);
              bq[j] = k2 + __shfl_xor_sync(0xffffffffu, a2, w);
            }
          }
          a[0] += __shfl_xor_sync(0xffffffffu, a[0], 16);
          bq[0] += __shfl_xor_sync(0xffffffffu, bq[0], 16);
          if (lane < 16) {   // one slot per (quarter, column): summed in a fixed order below
            s_stats[(q * 2 + 0) * BLOCK_N + cbeg + u * 16 + lane] = a[0];
            s_stats[(q * 2 + 1) * BLOCK_N + cbeg + u * 16 + lane] = bq[0];
          }
        }
      }
    }
  }
  __syncwarp();
  __syncthreads();
  if (p.stats != nullptr) {
    const int grp = img / p.imgs_per_group;
    const int cout = p.outc0 + p.outc1;
    for (int i = threadIdx.x; i < 2 * BLOCK_N; i += kHaloThreads) {
      const int st = i / BLOCK_N, j = i % BLOCK_N;
      const double tot = (static_cast<double>(s_stats[(0 * 2 + st) * BLOCK_N + j]) + s_stats[(1 * 2 + st) * BLOCK_N + j]) +
                         (static_cast<double>(s_stats[(2 * 2 + st) * BLOCK_N + j]) + s_stats[(3 * 2 + st) * BLOCK_N + j]);
      atomicAdd(p.stats + ((static_cast<long long>(job % kStatReplicas) * p.groups + grp) * cout + j) * 2 + st, tot);
    }
  }
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// ----------------------------------------------------------------------------------------------
// host
// ----------------------------------------------------------------------------------------------
static int halo_chunk_bytes(int bk) { return ((kBoxW * bk * 2) + 1023) / 1024 * 1024; }

static long long halo_smem_bytes(int cout, int ctot, int bk, int ring) {
  const int chunks = ctot / bk;
  return 9LL * chunks * cout * bk * 2 + static_cast<long long>(ring) * chunks * halo_chunk_bytes(bk) + 1024 + 512 +
         8 * cout * 4;
}
// deepest ring that fits: 8 slots if two CTAs still fit an SM (~100 KB each), else whatever fits in 200 KB
static int halo_ring(int cout, int ctot, int bk) {
  if (halo_smem_bytes(cout, ctot, bk, kMaxRing) <= 100 * 1024) return kMaxRing;
  int ring = kMaxRing;
  while (ring > 4 && halo_smem_bytes(cout, ctot, bk, ring) > 200 * 1024) --ring;
  return ring;
}

bool conv3x3_halo_applicable(int C0, int C1, int cout, int outc0, int outc1, int H, int W, int dil) {
  if (dil != 1 || W < 112 || H < 2) return false;
  if (cout != 32 && cout != 64 && cout != 96) return false;
  if (outc0 % 32 != 0 || outc1 % 32 != 0) return false;
  const int bk = (C0 % 64 == 0 && C1 % 64 == 0) ? 64 : 32;
  if (C0 % bk != 0 || C1 % bk != 0) return false;
  return halo_smem_bytes(cout, C0 + C1, bk, 4) <= 200 * 1024;
}

template <int BLOCK_N, int BK>
static int launch_halo(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const HaloParams& p, int jobs,
                       int smem, double flops, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    PP_CHECK_CUDA(cudaFuncSetAttribute(conv3x3_halo_tc_kernel<BLOCK_N, BK>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, 202 * 1024));
    attr_set = true;
  }
  const int slot = prof_begin(PROF_CONV, flops, stream);
  conv3x3_halo_tc_kernel<BLOCK_N, BK><<<jobs, kHaloThreads, smem, stream>>>(a0, a1, b, p);
  prof_end(slot, stream);
  PP_LAUNCH_CHECK();
  return PP_OK;
}

int conv3x3_halo_tc(const void* x0, int C0, const void* x1, int C1, const void* wpack, const float* bias, void* out0,
                    int outc0, int acc0, void* out1, int outc1, int acc1, int N, int H, int W, cudaStream_t stream,
                    double* stats, int groups, const ConvAffine* affine) {
  const int cout = outc0 + outc1, ctot = C0 + C1;
  const int bk = (C0 % 64 == 0 && C1 % 64 == 0) ? 64 : 32;
  HaloParams p{};
  p.N = N; p.H = H; p.W = W;
  p.strips = ceil_div(W, kTW);
  // rows per job: one wave of jobs over the resident CTA slots (2 per SM when the footprint allows), so that no
  // second, mostly idle wave is needed and the 2-row halo overhead stays small
  const int ring = halo_ring(cout, ctot, bk);
  const int slots = sm_count() * (halo_smem_bytes(cout, ctot, bk, ring) <= 100 * 1024 ? 2 : 1);
  int per = slots / (N * p.strips);
  if (per < 1) per = 1;
  int L = ceil_div(H, per);
  if (L < 8) L = 8;
  if (L > H) L = H;
  p.L = L;
  p.segs = ceil_div(H, L);
  p.kc0 = C0 / bk; p.kc1 = C1 / bk; p.c0 = C0; p.ctot = ctot;
  p.out0 = static_cast<__nv_bfloat16*>(out0); p.out1 = static_cast<__nv_bfloat16*>(out1);
  p.outc0 = outc0; p.outc1 = outc1; p.acc0 = acc0; p.acc1 = acc1; p.bias = bias;
  p.stats = stats;
  if (affine != nullptr) { p.ep_scale = affine->scale; p.ep_shift = affine->shift; p.ep_slope = affine->slope; }
  p.groups = groups > 0 ? groups : 1;
  p.imgs_per_group = N / p.groups;
  p.chunk_bytes = halo_chunk_bytes(bk);
  p.ring = halo_ring(cout, ctot, bk);
  const int jobs = N * p.strips * p.segs;

  CUtensorMap a0, a1, b;
  int rc = encode_tmap_nhwc(&a0, x0, N, H, W, C0, bk, kBoxW, 1, 1, bk == 64);
  if (rc) return rc;
  if (C1 > 0) rc = encode_tmap_nhwc(&a1, x1, N, H, W, C1, bk, kBoxW, 1, 1, bk == 64);
  else a1 = a0;
  if (rc) return rc;
  rc = encode_tmap_weights(&b, wpack, 9, cout, ctot, bk, cout, bk == 64);
  if (rc) return rc;
  const int smem = static_cast<int>(halo_smem_bytes(cout, ctot, bk, p.ring));
  const double flops = 2.0 * N * H * W * 9.0 * ctot * cout;
#define PP_HALO_CASE(BN_, BK_) \
  if (cout == BN_ && bk == BK_) return launch_halo<BN_, BK_>(a0, a1, b, p, jobs, smem, flops, stream);
  PP_HALO_CASE(32, 32) PP_HALO_CASE(64, 32) PP_HALO_CASE(96, 32) PP_HALO_CASE(32, 64) PP_HALO_CASE(64, 64)
  PP_HALO_CASE(96, 64)
#undef PP_HALO_CASE
  set_error("conv3x3_halo_tc: no kernel for cout=%d bk=%d", cout, bk);
  return PP_ERR_INVALID;
}

}  // namespace pp
