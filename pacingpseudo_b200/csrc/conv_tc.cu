// conv_tc.cu — 3x3 (dilated, stride-1, zero-padded) convolution as an implicit GEMM on the
// Blackwell 5th-gen tensor cores: TMA-im2col -> shared memory -> tcgen05.mma -> TMEM -> epilogue.
//
// Replaces the cuDNN calls behind nn.Conv2d in the reference ConvLayer
// (/root/reference/models/unet.py:188) and the aux-path bottleneck conv
// (/root/reference/models/aux_path_memory.py:24): forward, data gradient and weight gradient.
//
//   forward / dgrad :  D[pixel, co] = sum_{tap, ci} X[pixel + off(tap), ci] * Wp[tap, co, ci]
//       M = MT x 128 output pixels (MT consecutive {bw x bh x bn} boxes of the NHWC tensor, one TMEM accumulator
//       each), N = BLOCK_N channels, K = 9 taps x (C0 + C1) input channels walked in BK-channel slices. The A slice
//       of one tap is ONE TMA box of the activation tensor shifted by (dx*dil, dy*dil): out-of-bounds rows/cols are
//       zero-filled by the TMA unit, which is exactly the conv's zero padding. The channel concat of the decoder
//       (torch.cat((up, skip), 1), unet.py:151) is never materialised: the K loop walks two tensor maps. dgrad is
//       the same kernel on spatially flipped, transposed weights and can scatter its N tiles to two destination
//       tensors (the two concat sources). 320 threads: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer
//       (one elected lane each), warps 2..9 = epilogue (TMEM lane quarter = warp_id % 4, two warps per quarter).
//   wgrad (wide) :  dW[tap, co, ci] = sum_pixel dY[pixel, co] * X[pixel + off(tap), ci]
//       M = MT x 128 output channels, N = BLOCK_N input channels, K = pixels; both operands are the same NHWC TMA
//       boxes, consumed as MN-major UMMA operands. Split-K over pixel ranges: either fp32 red.global accumulation,
//       or (training step) per-split scratch slabs + a fixed-order reduction (deterministic). 192 threads.
//   wgrad (32/64-channel sources) :  taps packed into the MMA M dimension; the row variant fetches X as three
//       66-pixel row boxes per K block and realises the horizontal taps as descriptor row offsets; the default
//       (conv3x3_wgrad_rowsn_tc_kernel) packs the vertical taps into M and the horizontal taps into N (dY operand chunks
//       one pixel row apart) and walks strips of R image rows.
//
#include <array>
#include <map>
#include <mutex>

#include "pp_common.cuh"
#include "pp_ops.h"

namespace pp {

static constexpr int kTcThreads = 192;      // weight-gradient kernels: producer + MMA + 4 epilogue warps
static constexpr int kConvThreads = 320;    // forward / dgrad: producer + MMA + 8 epilogue warps (two per TMEM lane
                                            // quarter, alternating 32-column chunks): the epilogue of a tile is a
                                            // serial instruction stream per warp and is NOT overlapped with MMAs
static constexpr int kMaxStages = 8;

struct ConvTcParams {
  int N, H, W;
  int dil;
  int bw, bh, bn;              // pixel box (bw*bh*bn == 128)
  int tiles_w, tiles_h;        // tiles per row / column (tiles_n = gridDim.x / (tiles_w*tiles_h))
  int kc0, kc1;                // K slices per tap in source 0 / source 1
  int ctot;                    // C0 + C1
  int c0;                      // channels of source 0 (K offset of source 1 inside a tap)
  int stages;
  __nv_bfloat16* out0;         // destination of GEMM columns [0, outc0)
  __nv_bfloat16* out1;         // destination of GEMM columns [outc0, outc0+outc1)
  int outc0, outc1;
  int acc0, acc1;              // 1: out += result (read-modify-write)
  const float* bias;           // [outc0 + outc1] or null
  double* stats;               // optional [kStatReplicas][groups][cout][2] (sum, sum of squares) of the rounded output
  int imgs_per_group, groups;  // images per BatchNorm statistics group / number of groups (stats != null)
  const float* ep_scale;       // optional eval-mode BatchNorm + LeakyReLU epilogue: out = lrelu(acc * scale + shift)
  const float* ep_shift;
  float ep_slope;
};

// MT = consecutive 128-pixel M tiles per CTA (1, 2 or 4), each with its own TMEM accumulator; they share every weight
// (B operand) stage. The kernel is bound by L2 -> shared-memory traffic (TMA), not by the tensor pipe: per 64-channel
// K block a CTA fetches MT*16 KB of activations + BLOCK_N*128 B of weights for MT*128*BLOCK_N*64 MACs, so a larger
// MT x BLOCK_N footprint raises the FLOPs per fetched byte (128x256: 96 B/cycle/SM at full tensor rate, 256x256: 64).
template <int BLOCK_N, int BK, int MT>
__global__ void __launch_bounds__(kConvThreads, (BLOCK_N <= 96 && MT == 1) ? 2 : 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                  const __grid_constant__ CUtensorMap tmB, const ConvTcParams p) {
  constexpr int A_TILE = 128 * BK * 2;
  constexpr int A_BYTES = MT * A_TILE;
  constexpr int B_BYTES = BLOCK_N * BK * 2;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t SWZ = (BK == 64) ? SWZ_128B : SWZ_64B;
  constexpr uint32_t SBO = 8 * BK * 2;  // 8 rows of one swizzle atom
  constexpr int TMEM_NEED = MT * BLOCK_N;
  constexpr int TMEM_COLS = TMEM_NEED <= 32 ? 32 : (TMEM_NEED <= 64 ? 64 : (TMEM_NEED <= 128 ? 128 : (TMEM_NEED <= 256 ? 256 : 512)));
  static_assert(TMEM_NEED <= 512, "accumulators do not fit TMEM");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.stages * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tmem_full_bar = empty_bar + kMaxStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  float* s_stats = reinterpret_cast<float*>(tmem_slot + 2);   // [4 warps][2][BLOCK_N] BatchNorm partial sums

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int n_tile = blockIdx.y;
  const int tiles_per_img = p.tiles_w * p.tiles_h;
  const int num_k = 9 * (p.kc0 + p.kc1);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    if (p.kc1 > 0) tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = warp_uniform(*tmem_slot);

  if (warp == 0) {
    if (elect_one()) {
      // ===== TMA producer =====
      int tx0[MT], ty0[MT], tn0[MT];
#pragma unroll
      for (int m = 0; m < MT; ++m) {   // tiles past the end load out-of-range images: zero fill, never stored
        const int mt = blockIdx.x * MT + m;
        tx0[m] = (mt % p.tiles_w) * p.bw;
        ty0[m] = ((mt / p.tiles_w) % p.tiles_h) * p.bh;
        tn0[m] = (mt / tiles_per_img) * p.bn;
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int tap = 0; tap < 9; ++tap) {
        const int oy = (tap / 3 - 1) * p.dil, ox = (tap % 3 - 1) * p.dil;
        for (int kc = 0; kc < p.kc0 + p.kc1; ++kc) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
          int kofs;
          if (kc < p.kc0) {
#pragma unroll
            for (int m = 0; m < MT; ++m)
              tma_load_4d(sa + m * A_TILE, &tmA0, &full_bar[stage], kc * BK, tx0[m] + ox, ty0[m] + oy, tn0[m]);
            kofs = kc * BK;
          } else {
#pragma unroll
            for (int m = 0; m < MT; ++m)
              tma_load_4d(sa + m * A_TILE, &tmA1, &full_bar[stage], (kc - p.kc0) * BK, tx0[m] + ox, ty0[m] + oy, tn0[m]);
            kofs = p.c0 + (kc - p.kc0) * BK;
          }
          tma_load_3d(sb, &tmB, &full_bar[stage], kofs, n_tile * BLOCK_N, tap);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // ===== MMA issuer (elected thread; descriptors advance by 32-bit adds, see umma_bf16_lohi) =====
      constexpr uint32_t idesc = make_idesc_bf16(128, BLOCK_N, 0, 0);
      constexpr uint32_t dhi = smem_desc_hi(SBO, SWZ);
      const uint32_t base_lo = smem_desc_lo(smem_u32(smem), 16);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t a_lo = base_lo;
      for (int it = 0; it < num_k; ++it) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
#pragma unroll
          for (int m = 0; m < MT; ++m)
            umma_bf16_lohi(tmem_base + m * BLOCK_N, a_lo + m * (A_TILE >> 4) + k * 2, dhi,
                           a_lo + (A_BYTES >> 4) + k * 2, dhi, idesc, (it | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
        a_lo += STAGE_BYTES >> 4;
        if (++stage == p.stages) { stage = 0; phase ^= 1; a_lo = base_lo; }
      }
      umma_commit(tmem_full_bar);  // accumulators complete
    }
  } else {
    // ===== epilogue: TMEM -> registers -> (+bias, +old) -> bf16 NHWC =====
    const int q = warp & 3;          // TMEM lane quarter accessible to this warp
    const int half = (warp - 2) >> 2;   // the two warps of a quarter take alternating 32-column chunks
    const int r = q * 32 + lane;     // tile row == pixel index inside the box
    const int lx = r % p.bw, ly = (r / p.bw) % p.bh, ln = r / (p.bw * p.bh);
    const int col0 = n_tile * BLOCK_N;  // first GEMM column of this CTA
    const int et = threadIdx.x - 64;    // 0..255 among the epilogue threads
    const int cout = p.outc0 + p.outc1;

    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
#pragma unroll 1
    for (int m = 0; m < MT; ++m) {
      const int mt = blockIdx.x * MT + m;
      const int x0 = (mt % p.tiles_w) * p.bw, y0 = ((mt / p.tiles_w) % p.tiles_h) * p.bh;
      const int n0 = (mt / tiles_per_img) * p.bn;
      const int px = x0 + lx, py = y0 + ly, pn = n0 + ln;
      const bool valid = (px < p.W) && (py < p.H) && (pn < p.N);
      const long long pix = (static_cast<long long>(pn) * p.H + py) * p.W + px;
#pragma unroll 1
      for (int c = half * 32; c < BLOCK_N; c += 64) {
        // destination of this 32-column chunk (a CTA tile may straddle the two concat sources in dgrad)
        const int col = col0 + c;
        __nv_bfloat16* dst;
        int dstc, acc, ch;
        if (col < p.outc0) { dst = p.out0; dstc = p.outc0; acc = p.acc0; ch = col; }
        else               { dst = p.out1; dstc = p.outc1; acc = p.acc1; ch = col - p.outc0; }
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(m * BLOCK_N + c), v);
        tmem_wait_ld();
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
        if (p.bias != nullptr) {
          if ((reinterpret_cast<uintptr_t>(p.bias) & 15) == 0) {   // 8 broadcast vector loads per chunk
            const float4* b4 = reinterpret_cast<const float4*>(p.bias + col);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 bv = __ldg(b4 + j);
              f[4 * j] += bv.x; f[4 * j + 1] += bv.y; f[4 * j + 2] += bv.z; f[4 * j + 3] += bv.w;
            }
          } else {                                                 // any 4-byte aligned pointer is legal at the C ABI
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] += __ldg(p.bias + col + j);
          }
        }
        if (p.ep_scale != nullptr) {   // eval-mode BatchNorm + LeakyReLU on the fp32 accumulator (16-byte aligned arrays)
          const float4* s4 = reinterpret_cast<const float4*>(p.ep_scale + col);
          const float4* h4 = reinterpret_cast<const float4*>(p.ep_shift + col);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 sv = __ldg(s4 + j), hv = __ldg(h4 + j);
            f[4 * j] = lrelu(fmaf(f[4 * j], sv.x, hv.x), p.ep_slope);
            f[4 * j + 1] = lrelu(fmaf(f[4 * j + 1], sv.y, hv.y), p.ep_slope);
            f[4 * j + 2] = lrelu(fmaf(f[4 * j + 2], sv.z, hv.z), p.ep_slope);
            f[4 * j + 3] = lrelu(fmaf(f[4 * j + 3], sv.w, hv.w), p.ep_slope);
          }
        }
        if (valid) {
          __nv_bfloat16* o = dst + pix * dstc + ch;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            Vec8<__nv_bfloat16> pk;
            float t[8];
            if (acc) {
              pk.load(o + g * 8);
              pk.get(t);
#pragma unroll
              for (int j = 0; j < 8; ++j) f[g * 8 + j] += t[j];
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) t[j] = f[g * 8 + j];
            pk.set(t);
            pk.store(o + g * 8);
            pk.get(t);   // statistics are taken of the bf16-ROUNDED values (what BatchNorm will read back)
#pragma unroll
            for (int j = 0; j < 8; ++j) f[g * 8 + j] = t[j];
          }
        }
        if (p.stats != nullptr) {
          // column sums over this warp's 32 rows: butterfly transpose-reduce, lane j ends with column j
          float s1[32], s2[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) { s1[j] = valid ? f[j] : 0.f; s2[j] = s1[j] * s1[j]; }
#pragma unroll
          for (int w = 16; w >= 1; w >>= 1) {
            const bool hi = (lane & w) != 0;
#pragma unroll
            for (int j = 0; j < w; ++j) {
              const float a1 = hi ? s1[j] : s1[j + w], a2 = hi ? s2[j] : s2[j + w];
              const float k1 = hi ? s1[j + w] : s1[j], k2 = hi ? s2[j + w] : s2[j];
              s1[j] = k1 + __shfl_xor_sync(0xffffffffu, a1, w);
              s2[j] = k2 + __shfl_xor_sync(0xffffffffu, a2, w);
            }
          }
          s_stats[(q * 2 + 0) * BLOCK_N + c + lane] = s1[0];   // one slot per warp: summed in a fixed order below,
          s_stats[(q * 2 + 1) * BLOCK_N + c + lane] = s2[0];   // so a forward pass is bit-reproducible
        }
      }
      if (p.stats != nullptr) {
        // all rows of a tile belong to one statistics group (host guarantees bn | imgs_per_group or bn == 1)
        asm volatile("bar.sync 1, 256;" ::: "memory");   // the eight epilogue warps only
        if (n0 < p.N) {
          const int grp = n0 / p.imgs_per_group;
          for (int i = et; i < 2 * BLOCK_N; i += 256) {
            const int st = i / BLOCK_N, j = i % BLOCK_N, cc = col0 + j;
            // kStatReplicas interleaved copies of the accumulator spread the same-address atomics of thousands of CTAs
            const double tot = (static_cast<double>(s_stats[(0 * 2 + st) * BLOCK_N + j]) + s_stats[(1 * 2 + st) * BLOCK_N + j]) +
                               (static_cast<double>(s_stats[(2 * 2 + st) * BLOCK_N + j]) + s_stats[(3 * 2 + st) * BLOCK_N + j]);
            atomicAdd(p.stats + ((static_cast<long long>(mt % kStatReplicas) * p.groups + grp) * cout + cc) * 2 + st, tot);
          }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
    }
    tc_fence_before();
  }
  __syncwarp();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

bool conv3x3_halo_applicable(int C0, int C1, int cout, int outc0, int outc1, int H, int W, int dil);
int conv3x3_halo_tc(const void* x0, int C0, const void* x1, int C1, const void* wpack, const float* bias, void* out0,
                    int outc0, int acc0, void* out1, int outc1, int acc1, int N, int H, int W, cudaStream_t stream,
                    double* stats, int groups, const ConvAffine* affine);

// ----------------------------------------------------------------------------------------------
// wgrad
// ----------------------------------------------------------------------------------------------
struct WgradTcParams {
  int N, H, W, dil;
  int bw, bh, bn;            // pixel box of one K block (bw*bh*bn == 64)
  int tiles_w, tiles_h, tiles_total;
  int Cout, C0, C1;
  int ci_tiles0, ci_tiles1;  // N tiles per source
  int ctot, cbase0, cbase1;  // row length of dw and the column offset of each source inside it
  int oihw;                  // 1: dw is the OIHW gradient [Cout][ctot][9] itself (accumulated in place)
  int kb_per_split;
  int stages;
  float* dw;                 // [9][Cout][C0+C1] fp32, pre-zeroed, accumulated with red.global
  float* ws_split;           // != null: split s STORES its partial tile to ws_split[s][9][Cout][ctot] (no atomics,
                             // coalesced 16-byte stores after a shared-memory transpose; summed by wgrad_reduce)
};

// MT = number of 128-row output-channel blocks per CTA (1 or 2). MT = 2 gives a 256 x BLOCK_N tile in two TMEM
// accumulators that share every X (B operand) stage: the L2 -> SM traffic per FLOP, which bounds this kernel, drops
// by a third (48 KB -> 32 KB per 128x256x64 MMA block).
template <int BLOCK_N, int MT>
__global__ void __launch_bounds__(kTcThreads, 1)
conv3x3_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX0,
                        const __grid_constant__ CUtensorMap tmX1, const WgradTcParams p) {
  constexpr int PIXK = 64;                    // pixels (GEMM K) per pipeline stage
  constexpr int BOX_BYTES = PIXK * 128;       // one {64 ch x 64 px} box, 128-byte rows
  constexpr int A_BLOCK = 2 * BOX_BYTES;      // 128 output channels = 2 boxes
  constexpr int A_BYTES = MT * A_BLOCK;
  constexpr int NB = BLOCK_N / 64;            // boxes on the N side
  constexpr int B_BYTES = NB * BOX_BYTES;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int TMEM_COLS = MT * BLOCK_N < 32 ? 32 : MT * BLOCK_N;
  static_assert(MT * BLOCK_N <= 512, "accumulators do not fit TMEM");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.stages * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tmem_full_bar = empty_bar + kMaxStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int ci_tiles = p.ci_tiles0 + p.ci_tiles1;
  const int co_tile = blockIdx.x / ci_tiles;
  const int ci_tile = blockIdx.x % ci_tiles;
  const int tap = blockIdx.y;
  const int split = blockIdx.z;
  const bool src1 = ci_tile >= p.ci_tiles0;
  const int ci0 = (src1 ? ci_tile - p.ci_tiles0 : ci_tile) * BLOCK_N;  // channel offset inside the source
  const int co0 = co_tile * 128 * MT;
  const int kb_begin = split * p.kb_per_split;
  const int kb_end = min(kb_begin + p.kb_per_split, p.tiles_total);
  const int num_k = kb_end - kb_begin;
  const int oy = (tap / 3 - 1) * p.dil, ox = (tap % 3 - 1) * p.dil;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmDY);
    tma_prefetch_desc(src1 ? &tmX1 : &tmX0);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = warp_uniform(*tmem_slot);

  if (num_k > 0) {
    if (warp == 0) {
      if (elect_one()) {
        const CUtensorMap* tmX = src1 ? &tmX1 : &tmX0;
        int stage = 0;
        uint32_t phase = 0;
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          const int tw = kb % p.tiles_w;
          const int th = (kb / p.tiles_w) % p.tiles_h;
          const int tn = kb / (p.tiles_w * p.tiles_h);
          const int x0 = tw * p.bw, y0 = th * p.bh, n0 = tn * p.bn;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
#pragma unroll
          for (int b = 0; b < 2 * MT; ++b)
            tma_load_4d(sa + b * BOX_BYTES, &tmDY, &full_bar[stage], co0 + b * 64, x0, y0, n0);
#pragma unroll
          for (int b = 0; b < NB; ++b)
            tma_load_4d(sb + b * BOX_BYTES, tmX, &full_bar[stage], ci0 + b * 64, x0 + ox, y0 + oy, n0);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    } else if (warp == 1) {
      if (elect_one()) {
        constexpr uint32_t idesc = make_idesc_bf16(128, BLOCK_N, 1, 1);  // both operands MN-major
        // 16 pixels (K) = 16 rows of 128 B; MN chunks of 64 channels are BOX_BYTES apart (LBO);
        // groups of 8 K rows are 1024 B apart (SBO).
        constexpr uint32_t dhi = smem_desc_hi(1024, SWZ_128B);
        const uint32_t base_lo = smem_desc_lo(smem_u32(smem), BOX_BYTES);
        int stage = 0;
        uint32_t phase = 0;
        uint32_t a_lo = base_lo;
        for (int it = 0; it < num_k; ++it) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < PIXK / 16; ++k) {
#pragma unroll
            for (int m = 0; m < MT; ++m)
              umma_bf16_lohi(tmem_base + m * BLOCK_N, a_lo + m * (A_BLOCK >> 4) + k * (2048 >> 4), dhi,
                             a_lo + (A_BYTES >> 4) + k * (2048 >> 4), dhi, idesc, (it | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          a_lo += STAGE_BYTES >> 4;
          if (++stage == p.stages) { stage = 0; phase ^= 1; a_lo = base_lo; }
        }
        umma_commit(tmem_full_bar);
      }
    } else {
      const int q = warp & 3;
      const int csrc = src1 ? p.C1 : p.C0;
      const int ctot = p.ctot;
      const int cbase = (src1 ? p.cbase1 : p.cbase0) + ci0;
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after();
      if (p.ws_split != nullptr) {
        // ---- split-scratch mode: TMEM -> registers -> shared-memory transpose -> coalesced 16-byte stores.
        // The pipeline stages are idle now (every MMA has retired), so the transpose tile lives there.
        constexpr int PITCH = BLOCK_N + 4;   // floats; +4 keeps the float4 row writes of 32 lanes conflict-free
        float* tile = reinterpret_cast<float*>(smem);
        float* dst = p.ws_split + static_cast<long long>(split) * 9 * p.Cout * ctot;
#pragma unroll 1
        for (int m = 0; m < MT; ++m) {
          const int r = q * 32 + lane;
#pragma unroll 1
          for (int c = 0; c < BLOCK_N; c += 32) {
            uint32_t v[32];
            tmem_ld_32x32(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(m * BLOCK_N + c), v);
            tmem_wait_ld();
            float4* trow = reinterpret_cast<float4*>(tile + r * PITCH + c);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              trow[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                    __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");   // the four epilogue warps only
          for (int rr = q; rr < 128; rr += 4) {
            const int co = co0 + m * 128 + rr;
            if (co >= p.Cout) break;
            float* orow = dst + (static_cast<long long>(tap) * p.Cout + co) * ctot + cbase;
            for (int cc = lane * 4; cc < BLOCK_N; cc += 128)
              if (ci0 + cc < csrc)
                *reinterpret_cast<float4*>(orow + cc) = *reinterpret_cast<const float4*>(tile + rr * PITCH + cc);
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
        }
      } else {
        const int co = co0 + q * 32 + lane;
#pragma unroll 1
        for (int c = 0; c < BLOCK_N; c += 32) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(c), v);
          tmem_wait_ld();
          if (co < p.Cout) {
            if (p.oihw) {
              float* o = p.dw + (static_cast<long long>(co) * ctot + cbase + c) * 9 + tap;
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (ci0 + c + j < csrc) atomicAdd(o + j * 9, __uint_as_float(v[j]));
            } else {
              float* o = p.dw + (static_cast<long long>(tap) * p.Cout + co) * ctot + cbase + c;
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (ci0 + c + j < csrc) atomicAdd(o + j, __uint_as_float(v[j]));
            }
          }
        }
      }
      tc_fence_before();
    }
  } else if (p.ws_split != nullptr && warp >= 2) {
    // an empty split (cannot happen with the host's split sizing) must still define its scratch tile
    const int q = warp & 3;
    const int csrc = src1 ? p.C1 : p.C0;
    const int cbase = (src1 ? p.cbase1 : p.cbase0) + ci0;
    float* dst = p.ws_split + static_cast<long long>(split) * 9 * p.Cout * p.ctot;
    for (int rr = q; rr < 128 * MT; rr += 4) {
      const int co = co0 + rr;
      if (co >= p.Cout) break;
      float* orow = dst + (static_cast<long long>(tap) * p.Cout + co) * p.ctot + cbase;
      for (int cc = lane * 4; cc < BLOCK_N; cc += 128)
        if (ci0 + cc < csrc) *reinterpret_cast<float4*>(orow + cc) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  __syncwarp();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// ----------------------------------------------------------------------------------------------
// wgrad for narrow layers (one source with 32 or 64 channels, Cout 32 or 64): the full-resolution
// layers have millions of pixels (GEMM K) but a tiny 32x32..64x64 output per tap, so the generic kernel
// above would re-stream dY for each of the 9 taps and pad M to 128. Here ONE CTA owns a pixel range and all
// 9 taps: per K block it loads dY once and the 9 shifted X boxes, and packs taps into the MMA M dimension
//     D_g[(tap - g0) * CI + ci, co] += sum_pixel X[pixel + off(tap), ci] * dY[pixel, co]
// with 128 / CI taps per group g (CI = 32: tap groups {0-3, 4-7, 5-8}; CI = 64: {0-1, 2-3, 4-5, 6-7, 7-8};
// the last group overlaps its predecessor and only its new tap is written back). The X boxes of consecutive
// taps sit LBO bytes apart in shared memory, so one MN-major descriptor spans a whole group.
// ----------------------------------------------------------------------------------------------
struct WgradNarrowParams {
  int N, H, W, dil;
  int bw, bh, bn;            // pixel box of one K block (bw*bh*bn == pixk)
  int tiles_w, tiles_h, tiles_total;
  int Cout, Csrc, ctot, cbase;
  int kb_per_cta, stages, pixk;
  float* dw;
  int ci0;                   // first channel of the source tensor this launch covers (kx-in-N row kernel; else 0)
};

template <int CI, int NCOUT>
__global__ void __launch_bounds__(kTcThreads, 1)
conv3x3_wgrad_narrow_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY,
                               const WgradNarrowParams p) {
  constexpr int TPG = 128 / CI;                       // taps per MMA group
  constexpr int NG = (9 + TPG - 1) / TPG;             // 3 (CI = 32) or 5 (CI = 64)
  constexpr int ROW_A = CI * 2, ROW_B = (NCOUT < 64 ? NCOUT : 64) * 2;   // bytes per pixel row of a box
  constexpr uint32_t SWZ_A = CI == 64 ? SWZ_128B : SWZ_64B;
  constexpr uint32_t SWZ_B = NCOUT >= 64 ? SWZ_128B : SWZ_64B;
  constexpr int NB_B = NCOUT > 64 ? NCOUT / 64 : 1;   // dY boxes
  constexpr int TMEM_NEED = NG * NCOUT;
  constexpr int TMEM_COLS = TMEM_NEED <= 128 ? 128 : (TMEM_NEED <= 256 ? 256 : 512);
  static_assert(TMEM_NEED <= 512, "accumulators do not fit TMEM");

  const int xbox = p.pixk * ROW_A;                    // bytes of one X box
  const int ybox = p.pixk * ROW_B;                    // bytes of one dY box
  const int stage_bytes = 9 * xbox + NB_B * ybox;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tmem_full_bar = empty_bar + kMaxStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int kb_begin = blockIdx.x * p.kb_per_cta;
  const int kb_end = min(kb_begin + p.kb_per_cta, p.tiles_total);
  const int num_k = kb_end - kb_begin;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmDY);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = warp_uniform(*tmem_slot);

  if (num_k > 0) {
    if (warp == 0) {
      if (elect_one()) {
        int stage = 0;
        uint32_t phase = 0;
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          const int tw = kb % p.tiles_w;
          const int th = (kb / p.tiles_w) % p.tiles_h;
          const int tn = kb / (p.tiles_w * p.tiles_h);
          const int x0 = tw * p.bw, y0 = th * p.bh, n0 = tn * p.bn;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sx = smem + stage * stage_bytes;
          uint8_t* sy = sx + 9 * xbox;
          mbar_arrive_expect_tx(&full_bar[stage], stage_bytes);
#pragma unroll
          for (int tap = 0; tap < 9; ++tap)
            tma_load_4d(sx + tap * xbox, &tmX, &full_bar[stage], 0, x0 + (tap % 3 - 1) * p.dil,
                        y0 + (tap / 3 - 1) * p.dil, n0);
#pragma unroll
          for (int b = 0; b < NB_B; ++b) tma_load_4d(sy + b * ybox, &tmDY, &full_bar[stage], b * 64, x0, y0, n0);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    } else if (warp == 1) {
      if (elect_one()) {
        constexpr uint32_t idesc = make_idesc_bf16(128, NCOUT, 1, 1);  // both operands MN-major
        // K step = 16 pixel rows; MN chunks (one tap each) are xbox bytes apart (LBO); 8-row groups 8*ROW apart (SBO)
        constexpr uint32_t ahi = smem_desc_hi(8 * ROW_A, SWZ_A), bhi = smem_desc_hi(8 * ROW_B, SWZ_B);
        const int ksteps = p.pixk / 16;
        const uint32_t base_a = smem_desc_lo(smem_u32(smem), xbox);
        const uint32_t base_b = smem_desc_lo(smem_u32(smem) + 9 * xbox, ybox);
        const uint32_t stage16 = static_cast<uint32_t>(stage_bytes) >> 4, xbox16 = static_cast<uint32_t>(xbox) >> 4;
        int stage = 0;
        uint32_t phase = 0;
        uint32_t soff = 0;
        for (int it = 0; it < num_k; ++it) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
#pragma unroll
          for (int g = 0; g < NG; ++g) {
            const int g0 = (g == NG - 1) ? 9 - TPG : g * TPG;   // first tap of the group
            const uint32_t a_lo = base_a + soff + g0 * xbox16, b_lo = base_b + soff;
            for (int k = 0; k < ksteps; ++k)
              umma_bf16_lohi(tmem_base + g * NCOUT, a_lo + k * (16 * ROW_A >> 4), ahi, b_lo + k * (16 * ROW_B >> 4), bhi,
                             idesc, (it | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          soff += stage16;
          if (++stage == p.stages) { stage = 0; phase ^= 1; soff = 0; }
        }
        umma_commit(tmem_full_bar);
      }
    } else {
      const int q = warp & 3;
      const int r = q * 32 + lane;          // accumulator row = (tap - g0) * CI + ci
      const int ci = r % CI;
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after();
#pragma unroll 1
      for (int g = 0; g < NG; ++g) {
        const int g0 = (g == NG - 1) ? 9 - TPG : g * TPG;
        const int tap = g0 + r / CI;
        const int first_new = (g == NG - 1) ? (NG - 1) * TPG : g0;   // taps below were written by the previous group
        const bool row_ok = (tap >= first_new) && (ci < p.Csrc);
#pragma unroll 1
        for (int c = 0; c < NCOUT; c += 32) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(g * NCOUT + c), v);
          tmem_wait_ld();
          if (row_ok) {
            float* o = p.dw + (static_cast<long long>(tap) * p.Cout + c) * p.ctot + p.cbase + ci;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c + j < p.Cout) atomicAdd(o + static_cast<long long>(j) * p.ctot, __uint_as_float(v[j]));
          }
        }
      }
      tc_fence_before();
    }
  }
  __syncwarp();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// ----------------------------------------------------------------------------------------------
// Row variant of the narrow wgrad (dilation 1, W >= 64): the kernel above fetches NINE shifted X boxes per K block
// and is bound by that L2 -> shared-memory re-fetch (ncu: ~1 GB through TMA for 200 MB of operands). Here a K block
// is 64 consecutive pixels of one image row, and X arrives as THREE boxes of 66 pixels (rows y-1, y, y+1 with a
// one-pixel halo each side); the horizontal taps are shared-memory row offsets of the MN-major descriptor (+0, +1, +2
// pixel rows), the vertical taps are its MN chunks (LBO = box stride), so taps (ky = 0..2, kx) pack into one M = 128
// MMA per kx (CI = 32: 3 x 32 rows used) or two (CI = 64: ky {0,1} and {2,-}). The unused chunk reads whatever follows
// in shared memory; its accumulator rows are never stored.
// ----------------------------------------------------------------------------------------------
template <int CI, int NCOUT>
__global__ void __launch_bounds__(kTcThreads, 1)
conv3x3_wgrad_rows_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY,
                             const WgradNarrowParams p) {
  static_assert(NCOUT <= 64, "one dY box");
  constexpr int PIXK = 64, XROWS = PIXK + 2;
  constexpr int ROW_A = CI * 2, ROW_B = NCOUT * 2;
  constexpr uint32_t SWZ_A = CI == 64 ? SWZ_128B : SWZ_64B;
  constexpr uint32_t SWZ_B = NCOUT == 64 ? SWZ_128B : SWZ_64B;
  constexpr int XBOX = (XROWS * ROW_A + 1023) / 1024 * 1024;   // box stride (1 KB aligned)
  constexpr int YBOX = PIXK * ROW_B;
  constexpr int STAGE_BYTES = 3 * XBOX + YBOX;
  constexpr int GPK = CI == 32 ? 1 : 2;                         // MMA groups per kx
  constexpr int NG = 3 * GPK;
  constexpr int TMEM_NEED = NG * NCOUT;
  constexpr int TMEM_COLS = TMEM_NEED <= 128 ? 128 : (TMEM_NEED <= 256 ? 256 : 512);
  static_assert(TMEM_NEED <= 512, "accumulators do not fit TMEM");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // one spare X box after the last stage: the unused MN chunk of the last stage's MMAs reads up to XBOX bytes past it
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.stages * STAGE_BYTES + XBOX);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tmem_full_bar = empty_bar + kMaxStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int kb_begin = blockIdx.x * p.kb_per_cta;
  const int kb_end = min(kb_begin + p.kb_per_cta, p.tiles_total);
  const int num_k = kb_end - kb_begin;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmDY);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = warp_uniform(*tmem_slot);

  if (num_k > 0) {
    if (warp == 0) {
      if (elect_one()) {
        int stage = 0;
        uint32_t phase = 0;
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          const int tw = kb % p.tiles_w;
          const int y = (kb / p.tiles_w) % p.H;
          const int n = kb / (p.tiles_w * p.H);
          const int x0 = tw * PIXK;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sx = smem + stage * STAGE_BYTES;
          uint8_t* sy = sx + 3 * XBOX;
          mbar_arrive_expect_tx(&full_bar[stage], 3 * XROWS * ROW_A + YBOX);
#pragma unroll
          for (int ky = 0; ky < 3; ++ky)   // out-of-range rows / columns arrive as zeros: the conv's padding
            tma_load_4d(sx + ky * XBOX, &tmX, &full_bar[stage], 0, x0 - 1, y + ky - 1, n);
          tma_load_4d(sy, &tmDY, &full_bar[stage], 0, x0, y, n);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    } else if (warp == 1) {
      if (elect_one()) {
        constexpr uint32_t idesc = make_idesc_bf16(128, NCOUT, 1, 1);  // both operands MN-major
        constexpr uint32_t ahi = smem_desc_hi(8 * ROW_A, SWZ_A), bhi = smem_desc_hi(8 * ROW_B, SWZ_B);
        const uint32_t base_a = smem_desc_lo(smem_u32(smem), XBOX);
        const uint32_t base_b = smem_desc_lo(smem_u32(smem) + 3 * XBOX, YBOX);
        int stage = 0;
        uint32_t phase = 0;
        uint32_t soff = 0;
        for (int it = 0; it < num_k; ++it) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
#pragma unroll
          for (int g = 0; g < NG; ++g) {
            const int kx = g / GPK, kyb = (g % GPK) * 2;              // first vertical tap of the group
            const uint32_t a_lo = base_a + soff + kyb * (XBOX >> 4) + kx * (ROW_A >> 4), b_lo = base_b + soff;
#pragma unroll
            for (int k = 0; k < PIXK / 16; ++k)
              umma_bf16_lohi(tmem_base + g * NCOUT, a_lo + k * ROW_A, ahi, b_lo + k * ROW_B, bhi, idesc,
                             (it | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          soff += STAGE_BYTES >> 4;
          if (++stage == p.stages) { stage = 0; phase ^= 1; soff = 0; }
        }
        umma_commit(tmem_full_bar);
      }
    } else {
      const int q = warp & 3;
      const int r = q * 32 + lane;          // accumulator row = chunk * CI + ci
      const int ci = r % CI, chunk = r / CI;
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after();
#pragma unroll 1
      for (int g = 0; g < NG; ++g) {
        const int kx = g / GPK, ky = (g % GPK) * 2 + chunk;
        const bool row_ok = (ky < 3) && (chunk < (CI == 32 ? 3 : 2)) && (ci < p.Csrc);
        const int tap = ky * 3 + kx;
#pragma unroll 1
        for (int c = 0; c < NCOUT; c += 32) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(g * NCOUT + c), v);
          tmem_wait_ld();
          if (row_ok) {
            float* o = p.dw + (static_cast<long long>(tap) * p.Cout + c) * p.ctot + p.cbase + ci;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c + j < p.Cout) atomicAdd(o + static_cast<long long>(j) * p.ctot, __uint_as_float(v[j]));
          }
        }
      }
      tc_fence_before();
    }
  }
  __syncwarp();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// ----------------------------------------------------------------------------------------------
// Row variant with the horizontal taps packed into the MMA N dimension. The kernel above issues one M = 128, N = Cout
// (32 / 64) MMA per (kx, 16 pixels): 4 KB of X and 1-2 KB of dY read from shared memory for 16-32 tensor cycles, i.e.
// 130-320 B/cycle against the SM's 128 B/cycle — the operand reads, not HBM, pace it (measured 2.5-2.9 TB/s of
// algorithmic traffic). Here a K block is a strip of R image rows x 64 pixels: X arrives as R + 2 UNhaloed 64-pixel
// boxes (rows y0-1 ... y0+R; output row r uses boxes r, r+1, r+2 as its M chunks, so X crosses L2 -> shared memory
// (R + 2) / R times instead of 3 times) and every dY row as ONE 66-pixel box starting one pixel to the left. The three
// horizontal taps are three N chunks of the dY operand whose leading byte offset is ONE PIXEL ROW (64 / 128 B): chunk c
// reads dY shifted by c pixels, which is tap kx = 2 - c,
//     D[(ky, ci), (c, co)] += sum_k X[y + ky - 1, x0 + k, ci] * dY[y, x0 - 1 + k + c, co].
// One M = 128, N = 3 * Cout MMA per 16 pixels (two for CI = 64) instead of three (six): 2.1x fewer operand bytes per
// FLOP. Out-of-image X / dY pixels arrive as zeros (TMA), which is the convolution's padding on both sides.
// ----------------------------------------------------------------------------------------------
// image rows per K block of the kx-in-N row kernel: a strip of R rows needs R + 2 X rows, so X crosses L2 -> shared
// memory (R + 2) / R times instead of 3 times; R is what still leaves >= 3 pipeline stages in ~190 KB
__host__ __device__ constexpr int wgrad_rowsn_rows(int ci, int ncout) { return (ci == 32 && ncout == 32) ? 4 : 2; }

template <int CI, int NCOUT>
__global__ void __launch_bounds__(kTcThreads, 1)
conv3x3_wgrad_rowsn_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY,
                              const WgradNarrowParams p) {
  static_assert(NCOUT <= 64, "one dY box");
  constexpr int PIXK = 64, YROWS = PIXK + 2;
  constexpr int ROW_A = CI * 2, ROW_B = NCOUT * 2;
  constexpr uint32_t SWZ_A = CI == 64 ? SWZ_128B : SWZ_64B;
  constexpr uint32_t SWZ_B = NCOUT == 64 ? SWZ_128B : SWZ_64B;
  constexpr int R = wgrad_rowsn_rows(CI, NCOUT);                // image rows per K block
  constexpr int XBOX = PIXK * ROW_A;                            // 4 / 8 KB: 1 KB aligned
  constexpr int YBOX = (YROWS * ROW_B + 1023) / 1024 * 1024;
  constexpr int STAGE_BYTES = (R + 2) * XBOX + R * YBOX;        // [X rows y0-1 .. y0+R] [dY rows y0 .. y0+R-1]
  constexpr int NG = CI == 32 ? 1 : 2;                          // MMA groups: ky {0,1,2,-} or {0,1} + {2,-}
  constexpr int NN = 3 * NCOUT;                                 // MMA N: (kx chunk, co)
  constexpr int TMEM_NEED = NG * NN;
  constexpr int TMEM_COLS = TMEM_NEED <= 128 ? 128 : (TMEM_NEED <= 256 ? 256 : 512);
  static_assert(TMEM_NEED <= 512, "accumulators do not fit TMEM");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // 8 KB spare after the last stage: the unused M chunk of the strip's last rows reads one X box past the X boxes (into
  // the dY boxes, and for the last stage possibly past them); its accumulator rows are never stored
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.stages * STAGE_BYTES + 8192);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tmem_full_bar = empty_bar + kMaxStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int kb_begin = blockIdx.x * p.kb_per_cta;
  const int kb_end = min(kb_begin + p.kb_per_cta, p.tiles_total);
  const int num_k = kb_end - kb_begin;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmDY);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = warp_uniform(*tmem_slot);

  if (num_k > 0) {
    if (warp == 0) {
      if (elect_one()) {
        int stage = 0;
        uint32_t phase = 0;
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          const int tw = kb % p.tiles_w;
          const int y0 = ((kb / p.tiles_w) % p.tiles_h) * R;
          const int n = kb / (p.tiles_w * p.tiles_h);
          const int x0 = tw * PIXK;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sx = smem + stage * STAGE_BYTES;
          uint8_t* sy = sx + (R + 2) * XBOX;
          mbar_arrive_expect_tx(&full_bar[stage], (R + 2) * XBOX + R * YROWS * ROW_B);
#pragma unroll
          for (int j = 0; j < R + 2; ++j)   // out-of-range rows / columns arrive as zeros: the conv's padding
            tma_load_4d(sx + j * XBOX, &tmX, &full_bar[stage], p.ci0, x0, y0 + j - 1, n);
#pragma unroll
          for (int j = 0; j < R; ++j) tma_load_4d(sy + j * YBOX, &tmDY, &full_bar[stage], 0, x0 - 1, y0 + j, n);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    } else if (warp == 1) {
      if (elect_one()) {
        constexpr uint32_t idesc = make_idesc_bf16(128, NN, 1, 1);  // both operands MN-major
        constexpr uint32_t ahi = smem_desc_hi(8 * ROW_A, SWZ_A), bhi = smem_desc_hi(8 * ROW_B, SWZ_B);
        const uint32_t base_a = smem_desc_lo(smem_u32(smem), XBOX);                // M chunks: one X box (ky) apart
        const uint32_t base_b = smem_desc_lo(smem_u32(smem) + (R + 2) * XBOX, ROW_B);   // N chunks: ONE pixel row apart
        int stage = 0;
        uint32_t phase = 0;
        uint32_t soff = 0;
        for (int it = 0; it < num_k; ++it) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
#pragma unroll
          for (int r = 0; r < R; ++r) {      // output row y0 + r: X rows r, r+1, r+2 of the strip, dY row r
#pragma unroll
            for (int g = 0; g < NG; ++g) {
              const uint32_t a_lo = base_a + soff + (r + g * 2) * (XBOX >> 4), b_lo = base_b + soff + r * (YBOX >> 4);
#pragma unroll
              for (int k = 0; k < PIXK / 16; ++k)
                umma_bf16_lohi(tmem_base + g * NN, a_lo + k * ROW_A, ahi, b_lo + k * ROW_B, bhi, idesc,
                               (it | r | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit(&empty_bar[stage]);
          soff += STAGE_BYTES >> 4;
          if (++stage == p.stages) { stage = 0; phase ^= 1; soff = 0; }
        }
        umma_commit(tmem_full_bar);
      }
    } else {
      const int q = warp & 3;
      const int r = q * 32 + lane;          // accumulator row = chunk * CI + ci
      const int ci = r % CI, chunk = r / CI;
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after();
#pragma unroll 1
      for (int g = 0; g < NG; ++g) {
        const int ky = g * 2 + chunk;        // CI = 32: chunk 0..3 (3 unused); CI = 64: group 0 = {0,1}, group 1 = {2,-}
        const bool row_ok = (ky < 3) && (ci < p.Csrc);
#pragma unroll 1
        for (int c = 0; c < NN; c += 32) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(g * NN + c), v);
          tmem_wait_ld();
          if (row_ok) {
            const int kx = 2 - c / NCOUT, co0 = c % NCOUT;   // N chunk c / NCOUT reads dY shifted by that many pixels
            float* o = p.dw + (static_cast<long long>(ky * 3 + kx) * p.Cout + co0) * p.ctot + p.cbase + ci;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (co0 + j < p.Cout) atomicAdd(o + static_cast<long long>(j) * p.ctot, __uint_as_float(v[j]));
          }
        }
      }
      tc_fence_before();
    }
  }
  __syncwarp();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// ----------------------------------------------------------------------------------------------
// Host launchers
// ----------------------------------------------------------------------------------------------
static int pow2_ceil(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}
static void pixel_box(int W, int H, int pixels, int* bw, int* bh, int* bn) {
  *bw = pow2_ceil(W) < pixels ? pow2_ceil(W) : pixels;
  const int rest = pixels / *bw;
  *bh = pow2_ceil(H) < rest ? pow2_ceil(H) : rest;
  *bn = rest / *bh;
}
static int gcd_int(int a, int b) { return b == 0 ? a : gcd_int(b, a % b); }

static constexpr int conv_tc_stages(int block_n, int bk, int mt) {
  const int stage_bytes = mt * 128 * bk * 2 + block_n * bk * 2;
  // narrow single tiles are latency-bound per CTA: keep the footprint small enough for 2 CTAs per SM
  int stages = (((block_n <= 96 && mt == 1) ? 100 : 200) * 1024) / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  return stages;
}

template <int BLOCK_N, int BK, int MT>
static int launch_conv_tc(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, ConvTcParams p,
                          int m_tiles, int n_tiles, cudaStream_t stream) {
  constexpr int STAGE_BYTES = MT * 128 * BK * 2 + BLOCK_N * BK * 2;
  constexpr int TAIL = 1024 + 256 + 8 * BLOCK_N * 4;
  int stages = conv_tc_stages(BLOCK_N, BK, MT);
  if (stages < 2) stages = 2;
  p.stages = stages;
  const int smem = stages * STAGE_BYTES + TAIL;
  static bool attr_set = false;  // benign race: idempotent
  if (!attr_set) {
    PP_CHECK_CUDA(cudaFuncSetAttribute(conv3x3_tc_kernel<BLOCK_N, BK, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       200 * 1024 + TAIL));
    attr_set = true;
  }
  const double flops = 2.0 * p.N * p.H * p.W * 9.0 * p.ctot * (p.outc0 + p.outc1);
  const int slot = prof_begin(PROF_CONV, flops, stream);
  conv3x3_tc_kernel<BLOCK_N, BK, MT><<<dim3(ceil_div(m_tiles, MT), n_tiles), kConvThreads, smem, stream>>>(a0, a1, b, p);
  prof_end(slot, stream);
  PP_LAUNCH_CHECK();
  return PP_OK;
}

// Tile shape (output channels per CTA, M tiles per CTA) from a small cost model: a CTA's time is the larger of its
// tensor-pipe cycles and its L2 -> shared-memory fetch cycles (the chip sustains ~5.9 KB/cycle, i.e. ~40 B/cycle per
// SM when all SMs stream), plus a fixed prologue and a per-accumulator-chunk epilogue; the kernel's time is that times
// the number of waves of one CTA per SM. PP_CONV_MT=1 pins single-tile CTAs (A/B experiments).
static void conv_tc_pick_tile(int cout, int ktot, int bk, int m_tiles, int* block_n_out, int* mt_out) {
  static const int kTiles[6] = {256, 192, 128, 96, 64, 32};
  static int max_mt = -1;
  if (max_mt < 0) {
    const char* e = getenv("PP_CONV_MT");
    max_mt = (e != nullptr && e[0] >= '1' && e[0] <= '4') ? e[0] - '0' : 4;
  }
  const double sms = sm_count();
  double best = 1e300;
  *block_n_out = 32; *mt_out = 1;
  for (int t = 0; t < 6; ++t) {
    const int bn = kTiles[t];
    if (cout % bn != 0) continue;
    for (int mt = 1; mt <= max_mt; mt *= 2) {
      if (mt * bn > 512) break;
      if (mt > 1 && conv_tc_stages(bn, bk, mt) < 3) break;
      const double ctas = static_cast<double>(ceil_div(m_tiles, mt)) * (cout / bn);
      const int per_sm = (bn <= 96 && mt == 1) ? 2 : 1;               // co-resident CTAs
      const double waves = ceil(ctas / (sms * per_sm));
      const double mma = static_cast<double>(mt) * 128.0 * bn * 9.0 * ktot / 4096.0 * per_sm;
      const double l2 = 9.0 * ktot * 2.0 * (128.0 * mt + bn) / 40.0 * per_sm;
      const double cta = (mma > l2 ? mma : l2) + 5000.0 + mt * ((bn + 63) / 64) * 350.0 * per_sm;
      const double cost = waves * cta;
      if (cost < best) { best = cost; *block_n_out = bn; *mt_out = mt; }
    }
  }
}

static int conv3x3_generic(const void* x0, int C0, const void* x1, int C1, const void* wpack, const float* bias,
                           void* out0, int outc0, int acc0, void* out1, int outc1, int acc1, int N, int H, int W, int dil,
                           cudaStream_t stream, double* stats, int groups, const ConvAffine* affine);

// ----------------------------------------------------------------------------------------------
// Kernel choice for the wide layers. Candidates: the generic per-tap TMA kernel above (its own (BLOCK_N, MT) model)
// and the best few tilings of the shared-memory-resident kernel (conv_rows.cu: single CTA / CTA pair, BLOCK_N, MT).
// No closed-form model ranks them reliably across the UNet's shapes (wave quantisation against 148 SMs, pad-column
// waste, shared-memory operand bandwidth and L2 traffic all move with the shape), so the FIRST call with a new shape
// measures them — cuDNN's "benchmark" mode: each candidate runs the real convolution into the caller's buffers (same
// result every time, BatchNorm statistics switched off for the trial runs), timed with CUDA events on the caller's
// stream, ONE stream synchronisation per new shape (documented exception to "never syncs", like pp_init). The winner is
// cached per (shape, epilogue kind). Launches that ACCUMULATE into their destination cannot be repeated and launches
// inside a CUDA-graph capture cannot be timed: they use the cached winner of the same shape if there is one, else the
// cost models. PP_CONV_AUTOTUNE=0 uses the models only; PP_CONV_TUNE_DEBUG=1 prints the trials.
// ----------------------------------------------------------------------------------------------
struct ConvChoice {
  bool rows = false;
  RowsPlan plan{};
};

static double generic_model_cost(int N, int H, int W, int C0, int C1, int cout) {
  int bw, bh, bn;
  pixel_box(W, H, 128, &bw, &bh, &bn);
  const int m_tiles = ceil_div(W, bw) * ceil_div(H, bh) * ceil_div(N, bn);
  const int ktot = C0 + C1, bk = (C0 % 64 == 0 && C1 % 64 == 0) ? 64 : 32;
  int block_n, mt;
  conv_tc_pick_tile(cout, ktot, bk, m_tiles, &block_n, &mt);
  const double ctas = static_cast<double>(ceil_div(m_tiles, mt)) * (cout / block_n);
  const int per_sm = (block_n <= 96 && mt == 1) ? 2 : 1;
  const double waves = ceil(ctas / (sm_count() * per_sm));
  const double n_mma = static_cast<double>(mt) * 9.0 * ktot / 16.0;
  const double mma = n_mma * block_n / 2.0;
  const double tma_bytes = 9.0 * ktot * 2.0 * (128.0 * mt + block_n);
  const double smem_cyc = (n_mma * (4096.0 + 32.0 * block_n) + tma_bytes) / 120.0;
  const double l2 = tma_bytes / 40.0;
  double body = mma > smem_cyc ? mma : smem_cyc;
  if (l2 > body) body = l2;
  return waves * (body * per_sm + 6000.0 + mt * ((block_n + 63) / 64) * 450.0 * per_sm);
}

static int conv_choose(const void* x0, int C0, const void* x1, int C1, const void* wpack, const float* bias, void* out0,
                       int outc0, int acc0, void* out1, int outc1, int acc1, int N, int H, int W, int dil,
                       cudaStream_t stream, double* stats, int groups, const ConvAffine* affine, ConvChoice* choice) {
  static std::mutex mu;
  static std::map<std::array<int, 9>, ConvChoice> cache;
  static const int autotune = [] { const char* e = getenv("PP_CONV_AUTOTUNE"); return (e && e[0] == '0') ? 0 : 1; }();
  static const int debug = [] { const char* e = getenv("PP_CONV_TUNE_DEBUG"); return (e && e[0] == '1') ? 1 : 0; }();
  const int cout = outc0 + outc1;
  const std::array<int, 9> key = {N, H, W, C0, C1, outc0, outc1, dil,
                                  (bias != nullptr ? 1 : 0) | (stats != nullptr ? 2 : 0) | (affine != nullptr ? 4 : 0)};
  // PP_CONV_TUNE_FILE=<path>: measured choices are appended there and read back at start-up, so that a second process
  // (e.g. the same command under ncu) replays the first one's choices instead of measuring under the profiler
  static const char* tune_file = getenv("PP_CONV_TUNE_FILE");
  {
    std::lock_guard<std::mutex> lk(mu);
    static bool loaded = false;
    if (!loaded) {
      loaded = true;
      if (tune_file != nullptr) {
        if (FILE* f = fopen(tune_file, "r")) {
          std::array<int, 9> k;
          int rows;
          RowsPlan pl{};
          while (fscanf(f, "%d %d %d %d %d %d %d %d %d %d %d %d %d %d %d %d %d %d %d", &k[0], &k[1], &k[2], &k[3], &k[4], &k[5],
                        &k[6], &k[7], &k[8], &rows, &pl.block_n, &pl.bk, &pl.mt, &pl.pair, &pl.R, &pl.rbox, &pl.a_bytes,
                        &pl.nb, &pl.smem) == 19) {
            ConvChoice c;
            c.rows = rows != 0;
            c.plan = pl;
            cache[k] = c;
          }
          fclose(f);
        }
      }
    }
    auto it = cache.find(key);
    if (it != cache.end()) { *choice = it->second; return PP_OK; }
  }
  RowsPlan plans[6];
  const int np = conv3x3_rows_plans(N, H, W, dil, C0, C1, cout, plans, 6);
  ConvChoice best;
  if (np == 0) { *choice = best; return PP_OK; }
  static const int force = [] {   // PP_CONV_FORCE=rows | generic: profiling / A-B runs
    const char* e = getenv("PP_CONV_FORCE");
    return e == nullptr ? 0 : (e[0] == 'r' ? 1 : (e[0] == 'g' ? 2 : 0));
  }();
  if (force) {
    static const int force_plan = [] { const char* e = getenv("PP_CONV_FORCE_PLAN"); return e ? atoi(e) : 0; }();
    if (force == 1) { best.rows = true; best.plan = plans[force_plan < np ? force_plan : np - 1]; }
    *choice = best;
    return PP_OK;
  }
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(stream, &cap);
  const bool can_time = autotune && acc0 == 0 && acc1 == 0 && cap == cudaStreamCaptureStatusNone;
  if (!can_time) {   // cost models only; not cached, a later repeatable launch of the same shape may still measure
    if (plans[0].cost < generic_model_cost(N, H, W, C0, C1, cout)) { best.rows = true; best.plan = plans[0]; }
    *choice = best;
    return PP_OK;
  }
  cudaEvent_t e0, e1;
  PP_CHECK_CUDA(cudaEventCreate(&e0));
  PP_CHECK_CUDA(cudaEventCreate(&e1));
  float best_ms = 1e30f;
  for (int c = -1; c < np; ++c) {   // -1: the generic kernel
    int rc = PP_OK;
    for (int it = 0; it < 4 && rc == PP_OK; ++it) {   // one warm-up + three timed launches
      if (it == 1) cudaEventRecord(e0, stream);
      rc = c < 0 ? conv3x3_generic(x0, C0, x1, C1, wpack, bias, out0, outc0, 0, out1, outc1, 0, N, H, W, dil, stream, nullptr,
                                   groups, affine)
                 : conv3x3_rows_tc(x0, C0, x1, C1, wpack, bias, out0, outc0, 0, out1, outc1, 0, N, H, W, dil, stream, nullptr,
                                   groups, affine, &plans[c]);
    }
    if (rc) { cudaEventDestroy(e0); cudaEventDestroy(e1); return rc; }
    cudaEventRecord(e1, stream);
    PP_CHECK_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (debug) {
      if (c < 0) fprintf(stderr, "conv tune N=%d %dx%d C=%d+%d->%d+%d dil=%d: generic %.1f us\n", N, H, W, C0, C1, outc0, outc1,
                         dil, ms / 3 * 1e3);
      else fprintf(stderr, "   rows BLOCK_N=%d BK=%d MT=%d PAIR=%d R=%d: %.1f us (model %.0f)\n", plans[c].block_n, plans[c].bk,
                   plans[c].mt, plans[c].pair, plans[c].R, ms / 3 * 1e3, plans[c].cost);
    }
    if (ms < best_ms) {
      best_ms = ms;
      best.rows = c >= 0;
      if (c >= 0) best.plan = plans[c];
    }
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  {
    std::lock_guard<std::mutex> lk(mu);
    cache[key] = best;
    if (tune_file != nullptr) {
      if (FILE* f = fopen(tune_file, "a")) {
        fprintf(f, "%d %d %d %d %d %d %d %d %d %d %d %d %d %d %d %d %d %d %d\n", key[0], key[1], key[2], key[3], key[4], key[5],
                key[6], key[7], key[8], best.rows ? 1 : 0, best.plan.block_n, best.plan.bk, best.plan.mt, best.plan.pair,
                best.plan.R, best.plan.rbox, best.plan.a_bytes, best.plan.nb, best.plan.smem);
        fclose(f);
      }
    }
  }
  *choice = best;
  return PP_OK;
}

// x0:[N,H,W,C0] x1:[N,H,W,C1] (or null), wpack:[9][outc0+outc1][C0+C1] bf16.
int conv3x3_tc(const void* x0, int C0, const void* x1, int C1, const void* wpack, const float* bias, void* out0,
               int outc0, int acc0, void* out1, int outc1, int acc1, int N, int H, int W, int dil,
               cudaStream_t stream, double* stats, int groups, const ConvAffine* affine) {
  const int cout = outc0 + outc1;
  if (affine != nullptr && affine->scale == nullptr) affine = nullptr;
  if (affine != nullptr)
    PP_REQUIRE(affine->shift != nullptr && bias == nullptr && stats == nullptr && outc1 == 0 && acc0 == 0 &&
                   (reinterpret_cast<uintptr_t>(affine->scale) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(affine->shift) & 15) == 0,
               "conv3x3_tc: the affine epilogue needs 16-byte aligned scale/shift, one destination, no bias / "
               "statistics / accumulation");
  const int ctot = C0 + C1;
  PP_REQUIRE(N > 0 && H > 0 && W > 0 && dil >= 1, "conv3x3_tc: bad shape N=%d H=%d W=%d dil=%d", N, H, W, dil);
  PP_REQUIRE(C0 % 32 == 0 && C1 % 32 == 0 && C0 > 0, "conv3x3_tc: input channels must be multiples of 32 (C0=%d C1=%d)",
             C0, C1);
  PP_REQUIRE((x1 == nullptr) == (C1 == 0), "conv3x3_tc: x1/C1 mismatch");
  PP_REQUIRE((out1 == nullptr) == (outc1 == 0), "conv3x3_tc: out1/outc1 mismatch");
  PP_REQUIRE(outc0 % 32 == 0 && outc1 % 32 == 0, "conv3x3_tc: output channels must be multiples of 32 (outc0=%d outc1=%d)",
             outc0, outc1);
  if (conv3x3_halo_applicable(C0, C1, cout, outc0, outc1, H, W, dil))   // narrow high-resolution layers
    return conv3x3_halo_tc(x0, C0, x1, C1, wpack, bias, out0, outc0, acc0, out1, outc1, acc1, N, H, W, stream, stats,
                           groups, affine);
  if (conv3x3_rows_applicable(C0, C1, cout, N, H, W, dil)) {            // wide layers: generic kernel or conv_rows.cu
    ConvChoice ch;
    int rc = conv_choose(x0, C0, x1, C1, wpack, bias, out0, outc0, acc0, out1, outc1, acc1, N, H, W, dil, stream, stats,
                         groups, affine, &ch);
    if (rc) return rc;
    if (ch.rows)
      return conv3x3_rows_tc(x0, C0, x1, C1, wpack, bias, out0, outc0, acc0, out1, outc1, acc1, N, H, W, dil, stream,
                             stats, groups, affine, &ch.plan);
  }
  return conv3x3_generic(x0, C0, x1, C1, wpack, bias, out0, outc0, acc0, out1, outc1, acc1, N, H, W, dil, stream, stats,
                         groups, affine);
}

static int conv3x3_generic(const void* x0, int C0, const void* x1, int C1, const void* wpack, const float* bias,
                           void* out0, int outc0, int acc0, void* out1, int outc1, int acc1, int N, int H, int W, int dil,
                           cudaStream_t stream, double* stats, int groups, const ConvAffine* affine) {
  const int cout = outc0 + outc1, ctot = C0 + C1;
  const int bk = (C0 % 64 == 0 && C1 % 64 == 0) ? 64 : 32;
  ConvTcParams p{};
  p.N = N; p.H = H; p.W = W; p.dil = dil;
  pixel_box(W, H, 128, &p.bw, &p.bh, &p.bn);
  p.tiles_w = ceil_div(W, p.bw);
  p.tiles_h = ceil_div(H, p.bh);
  const int tiles_n = ceil_div(N, p.bn);
  const int m_tiles = p.tiles_w * p.tiles_h * tiles_n;
  // N tile (a divisor of the output channels; a tile may straddle the two dgrad destinations: the epilogue picks the
  // destination per 32-column chunk) and M tiles per CTA
  int block_n, mt;
  conv_tc_pick_tile(cout, ctot, bk, m_tiles, &block_n, &mt);
  p.stats = stats;
  p.imgs_per_group = groups > 0 ? N / groups : N;
  p.groups = groups > 0 ? groups : 1;
  if (stats != nullptr)
    PP_REQUIRE(groups >= 1 && N % groups == 0 && (p.bn == 1 || p.imgs_per_group % p.bn == 0),
               "conv3x3_tc: a pixel tile (%d images) would straddle BatchNorm statistics groups (%d images each)", p.bn,
               p.imgs_per_group);
  p.kc0 = C0 / bk; p.kc1 = C1 / bk; p.ctot = ctot; p.c0 = C0;
  p.out0 = static_cast<__nv_bfloat16*>(out0); p.out1 = static_cast<__nv_bfloat16*>(out1);
  p.outc0 = outc0; p.outc1 = outc1; p.acc0 = acc0; p.acc1 = acc1; p.bias = bias;
  if (affine != nullptr) { p.ep_scale = affine->scale; p.ep_shift = affine->shift; p.ep_slope = affine->slope; }

  CUtensorMap a0, a1, b;
  int rc = encode_tmap_nhwc(&a0, x0, N, H, W, C0, bk, p.bw, p.bh, p.bn, bk == 64);
  if (rc) return rc;
  if (C1 > 0) rc = encode_tmap_nhwc(&a1, x1, N, H, W, C1, bk, p.bw, p.bh, p.bn, bk == 64);
  else a1 = a0;
  if (rc) return rc;
  rc = encode_tmap_weights(&b, wpack, 9, cout, ctot, bk, block_n, bk == 64);
  if (rc) return rc;

  const int n_tiles = cout / block_n;
#define PP_CONV_CASE(BN_, BK_, MT_) \
  if (block_n == BN_ && bk == BK_ && mt == MT_) return launch_conv_tc<BN_, BK_, MT_>(a0, a1, b, p, m_tiles, n_tiles, stream);
#define PP_CONV_CASES(BK_)                                                                                        \
  PP_CONV_CASE(256, BK_, 1) PP_CONV_CASE(256, BK_, 2) PP_CONV_CASE(192, BK_, 1) PP_CONV_CASE(192, BK_, 2)          \
  PP_CONV_CASE(128, BK_, 1) PP_CONV_CASE(128, BK_, 2) PP_CONV_CASE(128, BK_, 4) PP_CONV_CASE(96, BK_, 1)           \
  PP_CONV_CASE(96, BK_, 2) PP_CONV_CASE(96, BK_, 4) PP_CONV_CASE(64, BK_, 1) PP_CONV_CASE(64, BK_, 2)              \
  PP_CONV_CASE(64, BK_, 4) PP_CONV_CASE(32, BK_, 1) PP_CONV_CASE(32, BK_, 2) PP_CONV_CASE(32, BK_, 4)
  PP_CONV_CASES(64)
  PP_CONV_CASES(32)
#undef PP_CONV_CASES
#undef PP_CONV_CASE
  set_error("conv3x3_tc: no kernel for block_n=%d bk=%d mt=%d", block_n, bk, mt);
  return PP_ERR_INVALID;
}

template <int BLOCK_N, int MT>
static int launch_wgrad_tc(const CUtensorMap& dy, const CUtensorMap& x0, const CUtensorMap& x1, WgradTcParams p,
                           dim3 grid, cudaStream_t stream) {
  constexpr int STAGE_BYTES = (2 * MT + BLOCK_N / 64) * 64 * 128;
  int stages = (200 * 1024) / STAGE_BYTES;
  if (stages > kMaxStages) stages = kMaxStages;
  p.stages = stages;
  const int smem = stages * STAGE_BYTES + 1024 + 256;
  static bool attr_set = false;
  if (!attr_set) {
    PP_CHECK_CUDA(cudaFuncSetAttribute(conv3x3_wgrad_tc_kernel<BLOCK_N, MT>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024 + 1024 + 256));
    attr_set = true;
  }
  const double flops = 2.0 * p.N * p.H * p.W * 9.0 * (p.C0 + p.C1) * p.Cout;
  const int slot = prof_begin(PROF_WGRAD, flops, stream);
  conv3x3_wgrad_tc_kernel<BLOCK_N, MT><<<grid, kTcThreads, smem, stream>>>(dy, x0, x1, p);
  prof_end(slot, stream);
  PP_LAUNCH_CHECK();
  return PP_OK;
}

// Sum of the per-split partial gradients ws[split][tap][Cout][Cin] -> OIHW grad [Cout][Cin][3][3] (+=) for input
// channels [ci_begin, ci_begin + ci_count): fixed summation order, so the weight gradient is bit-reproducible.
// Thread = one (co, ci) pair: per split its 9 tap loads are independent and coalesced over ci across the warp; the 9
// taps of a pair are contiguous in OIHW, so the warp's read-modify-write covers one contiguous 1152-byte span.
__global__ void __launch_bounds__(256) wgrad_reduce_unpack_kernel(const float* __restrict__ ws, float* __restrict__ g,
                                                                  int splits, int Cout, int Cin, int ci_begin,
                                                                  int ci_count) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= Cout * ci_count) return;
  const int co = idx / ci_count, ci = ci_begin + idx % ci_count;
  const size_t split_stride = static_cast<size_t>(9) * Cout * Cin, tap_stride = static_cast<size_t>(Cout) * Cin;
  const float* src = ws + static_cast<size_t>(co) * Cin + ci;
  float acc[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) acc[t] = 0.f;
  // These launches are latency-bound (a few thousand threads, `splits` dependent-looking round trips to L2 of 9 loads
  // each: ~15 us whatever the layer): four splits' loads are issued together, the sums keep their fixed order.
  int sp = 0;
  for (; sp + 4 <= splits; sp += 4) {
    float v[4][9];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int t = 0; t < 9; ++t) v[u][t] = __ldg(src + (sp + u) * split_stride + t * tap_stride);
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int t = 0; t < 9; ++t) acc[t] += v[u][t];
  }
  for (; sp < splits; ++sp) {
    float v[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) v[t] = __ldg(src + sp * split_stride + t * tap_stride);
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[t] += v[t];
  }
  float* o = g + (static_cast<size_t>(co) * Cin + ci) * 9;
#pragma unroll
  for (int t = 0; t < 9; ++t) o[t] += acc[t];
}

template <int CI, int NCOUT>
static int launch_wgrad_narrow(const void* dy, int Cout, const void* x, int Csrc, int ctot, int cbase, float* dw, int N,
                               int H, int W, int dil, cudaStream_t stream) {
  WgradNarrowParams p{};
  p.N = N; p.H = H; p.W = W; p.dil = dil;
  constexpr int ROW_A = CI * 2, ROW_B = (NCOUT < 64 ? NCOUT : 64) * 2, NB_B = NCOUT > 64 ? NCOUT / 64 : 1;
  p.pixk = (64 * (9 * ROW_A + NB_B * ROW_B) <= 48 * 1024) ? 64 : 32;
  pixel_box(W, H, p.pixk, &p.bw, &p.bh, &p.bn);
  p.tiles_w = ceil_div(W, p.bw);
  p.tiles_h = ceil_div(H, p.bh);
  p.tiles_total = p.tiles_w * p.tiles_h * ceil_div(N, p.bn);
  p.Cout = Cout; p.Csrc = Csrc; p.ctot = ctot; p.cbase = cbase; p.dw = dw;
  const int stage_bytes = p.pixk * (9 * ROW_A + NB_B * ROW_B);
  p.stages = (200 * 1024) / stage_bytes;
  if (p.stages > kMaxStages) p.stages = kMaxStages;
  int ctas = sm_count() < p.tiles_total ? sm_count() : p.tiles_total;
  p.kb_per_cta = ceil_div(p.tiles_total, ctas);
  ctas = ceil_div(p.tiles_total, p.kb_per_cta);
  CUtensorMap tx, tdy;
  int rc = encode_tmap_nhwc(&tx, x, N, H, W, Csrc, CI, p.bw, p.bh, p.bn, CI == 64);
  if (rc) return rc;
  rc = encode_tmap_nhwc(&tdy, dy, N, H, W, Cout, NCOUT < 64 ? NCOUT : 64, p.bw, p.bh, p.bn, NCOUT >= 64);
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    PP_CHECK_CUDA(cudaFuncSetAttribute(conv3x3_wgrad_narrow_tc_kernel<CI, NCOUT>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024 + 1024 + 256));
    attr_set = true;
  }
  const int smem = p.stages * stage_bytes + 1024 + 256;
  const double flops = 2.0 * N * H * W * 9.0 * Csrc * Cout;
  const int slot = prof_begin(PROF_WGRAD, flops, stream);
  conv3x3_wgrad_narrow_tc_kernel<CI, NCOUT><<<ctas, kTcThreads, smem, stream>>>(tx, tdy, p);
  prof_end(slot, stream);
  PP_LAUNCH_CHECK();
  return PP_OK;
}

template <int CI, int NCOUT>
static int launch_wgrad_rows(const void* dy, int Cout, const void* x, int Csrc, int ctot, int cbase, float* dw, int N,
                             int H, int W, cudaStream_t stream) {
  WgradNarrowParams p{};
  p.N = N; p.H = H; p.W = W; p.dil = 1;
  constexpr int ROW_A = CI * 2, ROW_B = NCOUT * 2;
  constexpr int XBOX = (66 * ROW_A + 1023) / 1024 * 1024;
  constexpr int STAGE_BYTES = 3 * XBOX + 64 * ROW_B;
  p.pixk = 64;
  p.bw = 64; p.bh = 1; p.bn = 1;
  p.tiles_w = ceil_div(W, 64);
  p.tiles_h = H;
  p.tiles_total = p.tiles_w * H * N;
  p.Cout = Cout; p.Csrc = Csrc; p.ctot = ctot; p.cbase = cbase; p.dw = dw;
  p.stages = (200 * 1024 - XBOX) / STAGE_BYTES;
  if (p.stages > kMaxStages) p.stages = kMaxStages;
  int ctas = sm_count() < p.tiles_total ? sm_count() : p.tiles_total;
  p.kb_per_cta = ceil_div(p.tiles_total, ctas);
  ctas = ceil_div(p.tiles_total, p.kb_per_cta);
  CUtensorMap tx, tdy;
  int rc = encode_tmap_nhwc(&tx, x, N, H, W, Csrc, CI, 66, 1, 1, CI == 64);
  if (rc) return rc;
  rc = encode_tmap_nhwc(&tdy, dy, N, H, W, Cout, NCOUT, 64, 1, 1, NCOUT == 64);
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    PP_CHECK_CUDA(cudaFuncSetAttribute(conv3x3_wgrad_rows_tc_kernel<CI, NCOUT>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024 + 1024 + 256));
    attr_set = true;
  }
  const int smem = p.stages * STAGE_BYTES + XBOX + 1024 + 256;
  const double flops = 2.0 * N * H * W * 9.0 * Csrc * Cout;
  const int slot = prof_begin(PROF_WGRAD, flops, stream);
  conv3x3_wgrad_rows_tc_kernel<CI, NCOUT><<<ctas, kTcThreads, smem, stream>>>(tx, tdy, p);
  prof_end(slot, stream);
  PP_LAUNCH_CHECK();
  return PP_OK;
}

// x holds Cx channels per pixel; the launch covers its channels [ci0, ci0 + Csrc) (Csrc <= CI), written to the dw columns
// [cbase, cbase + Csrc)
template <int CI, int NCOUT>
static int launch_wgrad_rowsn(const void* dy, int Cout, const void* x, int Csrc, int ctot, int cbase, float* dw, int N,
                              int H, int W, cudaStream_t stream, int Cx = 0, int ci0 = 0) {
  if (Cx == 0) Cx = Csrc;
  WgradNarrowParams p{};
  p.ci0 = ci0;
  p.N = N; p.H = H; p.W = W; p.dil = 1;
  constexpr int ROW_A = CI * 2, ROW_B = NCOUT * 2;
  constexpr int R = wgrad_rowsn_rows(CI, NCOUT);
  constexpr int XBOX = 64 * ROW_A, YBOX = (66 * ROW_B + 1023) / 1024 * 1024;
  constexpr int STAGE_BYTES = (R + 2) * XBOX + R * YBOX;
  p.pixk = 64;
  p.bw = 64; p.bh = R; p.bn = 1;
  p.tiles_w = ceil_div(W, 64);
  p.tiles_h = ceil_div(H, R);
  p.tiles_total = p.tiles_w * p.tiles_h * N;
  p.Cout = Cout; p.Csrc = Csrc; p.ctot = ctot; p.cbase = cbase; p.dw = dw;
  p.stages = (200 * 1024 - 8192) / STAGE_BYTES;
  if (p.stages > kMaxStages) p.stages = kMaxStages;
  int ctas = sm_count() < p.tiles_total ? sm_count() : p.tiles_total;
  p.kb_per_cta = ceil_div(p.tiles_total, ctas);
  ctas = ceil_div(p.tiles_total, p.kb_per_cta);
  CUtensorMap tx, tdy;
  int rc = encode_tmap_nhwc(&tx, x, N, H, W, Cx, CI, 64, 1, 1, CI == 64);
  if (rc) return rc;
  rc = encode_tmap_nhwc(&tdy, dy, N, H, W, Cout, NCOUT, 66, 1, 1, NCOUT == 64);
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    PP_CHECK_CUDA(cudaFuncSetAttribute(conv3x3_wgrad_rowsn_tc_kernel<CI, NCOUT>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024 + 1024 + 256));
    attr_set = true;
  }
  const int smem = p.stages * STAGE_BYTES + 8192 + 1024 + 256;
  const double flops = 2.0 * N * H * W * 9.0 * Csrc * Cout;
  const int slot = prof_begin(PROF_WGRAD, flops, stream);
  conv3x3_wgrad_rowsn_tc_kernel<CI, NCOUT><<<ctas, kTcThreads, smem, stream>>>(tx, tdy, p);
  prof_end(slot, stream);
  PP_LAUNCH_CHECK();
  return PP_OK;
}

static int wgrad_rows_mode() {   // PP_WGRAD_ROWS: 0 = nine-box kernel, 1 = row kernel (kx as A offsets), 2 = kx packed into N
  static const int mode = [] {
    const char* e = getenv("PP_WGRAD_ROWS");
    return (e != nullptr && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 2;
  }();
  return mode;
}
// Sources that take the narrow (pixel-range per CTA, all taps) kernels: 32 / 64 channels into 32 / 64; with the
// kx-in-N row kernel also 128 ... 256 channels as 64-channel column groups (the 192 -> 64 layer at 128^2: its 128-channel
// source went through the generic kernel with a half-empty M = 128 tile and nine passes over X and dY, 131 us).
static bool narrow_ok(int Csrc, int Cout, int W, int dil) {
  if (Cout != 32 && Cout != 64) return false;
  if (Csrc == 32 || Csrc == 64) return true;
  return wgrad_rows_mode() == 2 && dil == 1 && W >= 64 && Csrc % 64 == 0 && Csrc <= 256;
}

static int wgrad_narrow(const void* dy, int Cout, const void* x, int Csrc, int ctot, int cbase, float* dw, int N, int H,
                        int W, int dil, cudaStream_t stream) {
  const int rows_on = wgrad_rows_mode();
  if (Csrc > 64) {   // 64-channel column groups of a wider source (narrow_ok() admits them for the kx-in-N kernel only)
    for (int c0 = 0; c0 < Csrc; c0 += 64) {
      const int rc = Cout == 32 ? launch_wgrad_rowsn<64, 32>(dy, Cout, x, 64, ctot, cbase + c0, dw, N, H, W, stream, Csrc, c0)
                                : launch_wgrad_rowsn<64, 64>(dy, Cout, x, 64, ctot, cbase + c0, dw, N, H, W, stream, Csrc, c0);
      if (rc) return rc;
    }
    return PP_OK;
  }
  if (rows_on == 2 && dil == 1 && W >= 64) {
    if (Csrc == 32 && Cout == 32) return launch_wgrad_rowsn<32, 32>(dy, Cout, x, Csrc, ctot, cbase, dw, N, H, W, stream);
    if (Csrc == 32 && Cout == 64) return launch_wgrad_rowsn<32, 64>(dy, Cout, x, Csrc, ctot, cbase, dw, N, H, W, stream);
    if (Csrc == 64 && Cout == 32) return launch_wgrad_rowsn<64, 32>(dy, Cout, x, Csrc, ctot, cbase, dw, N, H, W, stream);
    return launch_wgrad_rowsn<64, 64>(dy, Cout, x, Csrc, ctot, cbase, dw, N, H, W, stream);
  }
  if (rows_on && dil == 1 && W >= 64) {   // row variant: X fetched 3x instead of 9x (a ragged last 64-pixel block
                                          // is zero-filled by TMA and contributes nothing)
    if (Csrc == 32 && Cout == 32) return launch_wgrad_rows<32, 32>(dy, Cout, x, Csrc, ctot, cbase, dw, N, H, W, stream);
    if (Csrc == 32 && Cout == 64) return launch_wgrad_rows<32, 64>(dy, Cout, x, Csrc, ctot, cbase, dw, N, H, W, stream);
    if (Csrc == 64 && Cout == 32) return launch_wgrad_rows<64, 32>(dy, Cout, x, Csrc, ctot, cbase, dw, N, H, W, stream);
    return launch_wgrad_rows<64, 64>(dy, Cout, x, Csrc, ctot, cbase, dw, N, H, W, stream);
  }
  if (Csrc == 32 && Cout == 32) return launch_wgrad_narrow<32, 32>(dy, Cout, x, Csrc, ctot, cbase, dw, N, H, W, dil, stream);
  if (Csrc == 32 && Cout == 64) return launch_wgrad_narrow<32, 64>(dy, Cout, x, Csrc, ctot, cbase, dw, N, H, W, dil, stream);
  if (Csrc == 64 && Cout == 32) return launch_wgrad_narrow<64, 32>(dy, Cout, x, Csrc, ctot, cbase, dw, N, H, W, dil, stream);
  return launch_wgrad_narrow<64, 64>(dy, Cout, x, Csrc, ctot, cbase, dw, N, H, W, dil, stream);
}

// K-split count: minimise  waves x (K blocks per CTA + fixed per-CTA cost) x time per K block  +  the reduction pass
// over `splits` partial gradients (per_split_bytes each), one CTA per SM
static int wgrad_plan_splits(int base_ctas, int tiles_total, int max_splits, double us_per_kb, double per_split_bytes) {
  const int sms = sm_count();
  const double ovh = 6.0;           // prologue + epilogue of a CTA, in units of one 64-pixel K block
  const double reduce_bw = 2.5e6;   // bytes per microsecond the reduction pass sustains on these small tensors
  int best = 1;
  double best_cost = 1e300;
  for (int sp = 1; sp <= max_splits && sp <= tiles_total; ++sp) {
    const int kb = ceil_div(tiles_total, sp);
    if (ceil_div(tiles_total, kb) != sp) continue;   // would leave an empty split
    const double cost = static_cast<double>(ceil_div(base_ctas * sp, sms)) * (kb + ovh) * us_per_kb +
                        sp * per_split_bytes / reduce_bw;
    if (cost < best_cost - 1e-9) { best_cost = cost; best = sp; }
  }
  return best;
}

// generic kernel over the given sources; ctot / cbase place them inside the rows of dw.
// ws_split != nullptr (and g_oihw): every K split stores its partial gradient to ws_split[split] and a fixed-order
// reduction folds them into the OIHW gradient (no atomics, deterministic); 256-row tiles when Cout % 256 == 0.
static int wgrad_wide(const void* dy, int Cout, const void* x0, int C0, const void* x1, int C1, int ctot, int cbase0,
                      int cbase1, float* dw, int oihw, int N, int H, int W, int dil, cudaStream_t stream,
                      float* ws_split = nullptr, long long ws_floats = 0, float* g_oihw = nullptr) {
  WgradTcParams p{};
  p.oihw = oihw;
  p.N = N; p.H = H; p.W = W; p.dil = dil;
  pixel_box(W, H, 64, &p.bw, &p.bh, &p.bn);
  p.tiles_w = ceil_div(W, p.bw);
  p.tiles_h = ceil_div(H, p.bh);
  p.tiles_total = p.tiles_w * p.tiles_h * ceil_div(N, p.bn);
  p.Cout = Cout; p.C0 = C0; p.C1 = C1; p.dw = dw;
  p.ctot = ctot; p.cbase0 = cbase0; p.cbase1 = cbase1;
  const int cmax = C0 > C1 ? C0 : C1;
  const int block_n = cmax <= 64 ? 64 : (cmax <= 128 ? 128 : 256);
  p.ci_tiles0 = ceil_div(C0, block_n);
  p.ci_tiles1 = C1 > 0 ? ceil_div(C1, block_n) : 0;
  const long long per_split = 9LL * Cout * ctot;
  const bool split_mode = ws_split != nullptr && g_oihw != nullptr && ws_floats >= per_split && Cout % 32 == 0 &&
                          C0 % 32 == 0 && C1 % 32 == 0 && ctot % 4 == 0 && cbase0 % 4 == 0 && cbase1 % 4 == 0;
  const int mt = (split_mode && Cout % 256 == 0 && block_n >= 128) ? 2 : 1;
  const int co_tiles = ceil_div(Cout, 128 * mt);
  const int base_ctas = co_tiles * (p.ci_tiles0 + p.ci_tiles1) * 9;
  int splits;
  if (split_mode) {
    long long cap = ws_floats / per_split;
    if (cap > 64) cap = 64;
    splits = wgrad_plan_splits(base_ctas, p.tiles_total, static_cast<int>(cap), 0.45 * mt * block_n / 256.0 + 0.1,
                               4.0 * per_split);
    p.ws_split = ws_split;
  } else {
    // atomic accumulation: at most ~2 full waves of one CTA per SM (fewer partial-tile reductions)
    splits = (2 * sm_count()) / base_ctas;
    if (splits > p.tiles_total) splits = p.tiles_total;
    if (splits < 1) splits = 1;
  }
  p.kb_per_split = ceil_div(p.tiles_total, splits);
  splits = ceil_div(p.tiles_total, p.kb_per_split);

  CUtensorMap tdy, tx0, tx1;
  int rc = encode_tmap_nhwc(&tdy, dy, N, H, W, Cout, 64, p.bw, p.bh, p.bn, true);
  if (rc) return rc;
  rc = encode_tmap_nhwc(&tx0, x0, N, H, W, C0, 64, p.bw, p.bh, p.bn, true);
  if (rc) return rc;
  if (C1 > 0) rc = encode_tmap_nhwc(&tx1, x1, N, H, W, C1, 64, p.bw, p.bh, p.bn, true);
  else tx1 = tx0;
  if (rc) return rc;
  dim3 grid(co_tiles * (p.ci_tiles0 + p.ci_tiles1), 9, splits);
  if (mt == 2) rc = block_n == 128 ? launch_wgrad_tc<128, 2>(tdy, tx0, tx1, p, grid, stream)
                                   : launch_wgrad_tc<256, 2>(tdy, tx0, tx1, p, grid, stream);
  else if (block_n == 64) rc = launch_wgrad_tc<64, 1>(tdy, tx0, tx1, p, grid, stream);
  else if (block_n == 128) rc = launch_wgrad_tc<128, 1>(tdy, tx0, tx1, p, grid, stream);
  else rc = launch_wgrad_tc<256, 1>(tdy, tx0, tx1, p, grid, stream);
  if (rc || !split_mode) return rc;
  // fold the splits into the OIHW gradient, source by source (their channel ranges inside a dw row)
  const int begins[2] = {cbase0, cbase1}, counts[2] = {C0, C1};
  for (int k = 0; k < 2; ++k) {
    if (counts[k] == 0) continue;
    wgrad_reduce_unpack_kernel<<<ceil_div(Cout * counts[k], 256), 256, 0, stream>>>(ws_split, g_oihw, splits, Cout, ctot,
                                                                                   begins[k], counts[k]);
    PP_LAUNCH_CHECK();
  }
  return PP_OK;
}

int unpack_wgrad_range(const float* dwp, float* g, int Cout, int Cin, int ci_begin, int ci_count, int accumulate,
                       cudaStream_t s);

bool conv3x3_wgrad_tc_uses_scratch(int Cout, int C0, int C1, int W, int dil) {
  return narrow_ok(C0, Cout, W, dil) || (C1 > 0 && narrow_ok(C1, Cout, W, dil));
}

// dy:[N,H,W,Cout] bf16, x0/x1 as in forward.
//   g_oihw == nullptr: dwp[9][Cout][C0+C1] fp32 packed gradient, accumulated (caller zeroes it).
//   g_oihw != nullptr: the OIHW gradient [Cout][C0+C1][3][3] is accumulated in place. Wide sources go there
//     directly from the epilogue; narrow sources go through dwp (caller zeroes it when
//     conv3x3_wgrad_tc_uses_scratch()) and are folded in by a column-range unpack.
//   ws_split (optional, ws_floats fp32 elements): scratch for the deterministic split-K path of the wide sources.
int conv3x3_wgrad_tc(const void* dy, int Cout, const void* x0, int C0, const void* x1, int C1, float* dwp,
                     float* g_oihw, int N, int H, int W, int dil, cudaStream_t stream, float* ws_split,
                     long long ws_floats) {
  PP_REQUIRE(N > 0 && H > 0 && W > 0 && dil >= 1, "conv3x3_wgrad_tc: bad shape");
  PP_REQUIRE(Cout % 8 == 0 && C0 % 8 == 0 && C1 % 8 == 0 && C0 > 0, "conv3x3_wgrad_tc: channels must be multiples of 8");
  PP_REQUIRE((x1 == nullptr) == (C1 == 0), "conv3x3_wgrad_tc: x1/C1 mismatch");
  const int ctot = C0 + C1;
  const bool n0 = narrow_ok(C0, Cout, W, dil), n1 = C1 > 0 && narrow_ok(C1, Cout, W, dil);
  const int oihw = g_oihw != nullptr;
  float* wide_dst = oihw ? g_oihw : dwp;
  int rc = PP_OK;
  if (n0) rc = wgrad_narrow(dy, Cout, x0, C0, ctot, 0, dwp, N, H, W, dil, stream);
  if (rc) return rc;
  if (n1) rc = wgrad_narrow(dy, Cout, x1, C1, ctot, C0, dwp, N, H, W, dil, stream);
  if (rc) return rc;
  if (!n0 && C1 > 0 && !n1) {
    rc = wgrad_wide(dy, Cout, x0, C0, x1, C1, ctot, 0, C0, wide_dst, oihw, N, H, W, dil, stream, ws_split, ws_floats,
                    g_oihw);
  } else {
    if (!n0) rc = wgrad_wide(dy, Cout, x0, C0, nullptr, 0, ctot, 0, 0, wide_dst, oihw, N, H, W, dil, stream, ws_split,
                             ws_floats, g_oihw);
    if (rc) return rc;
    if (C1 > 0 && !n1) rc = wgrad_wide(dy, Cout, x1, C1, nullptr, 0, ctot, C0, 0, wide_dst, oihw, N, H, W, dil, stream,
                                       ws_split, ws_floats, g_oihw);
  }
  if (rc) return rc;
  if (oihw) {
    if (n0) rc = unpack_wgrad_range(dwp, g_oihw, Cout, ctot, 0, C0, 1, stream);
    if (rc) return rc;
    if (n1) rc = unpack_wgrad_range(dwp, g_oihw, Cout, ctot, C0, C1, 1, stream);
  }
  return rc;
}

}  // namespace pp
