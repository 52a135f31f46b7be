// conv_rows.cu — 3x3 (dilated, stride-1, zero-padded) convolution for the WIDE layers (64..1024 channels at 64^2, 32^2,
// 28^2 ... maps): implicit GEMM on tcgen05 with a SHARED-MEMORY-RESIDENT im2col.
//
// Why: conv_tc.cu fetches the 128-pixel A tile of every tap separately (9 TMA boxes of the same pixels, shifted), so a
// 256 x 256 CTA tile pulls 64 B/cycle through TMA at full tensor rate while the chip's L2 sustains ~42 B/cycle/SM with
// all SMs streaming (ncu, round 1: 10-11.6 TB/s L2->SM with the tensor pipe only 68-75 % active). Here the input of
// one K chunk (64 or 32 channels) arrives ONCE per CTA as a single TMA box of whole zero-padded rows,
//     box = {BK ch, Wp = W + 2 dil pixels, rbox rows, 1 image} at (x, y) = (-dil, y0 - dil)   (TMA zero-fills the border)
// and lands densely in shared memory as rbox * Wp consecutive 128-byte (64-byte) rows. An output "position"
// p = y * Wp + x of that padded raster needs, for tap (ky, kx), input row index  p + ky*dil*Wp + kx*dil  of the box —
// the SAME linear shift for every position. So the A operand of tap (ky, kx) for 128 consecutive positions is the box
// itself, read through a UMMA descriptor whose start address is advanced by (ky*dil*Wp + kx*dil) rows; the UMMA swizzle
// is a function of the absolute shared-memory address (hardware experiment tests/cuda/exp_desc_shift.cu), so a K-major
// swizzled descriptor may start at any row. Positions with x >= W (the 2*dil pad columns of every row) and y >= H are
// computed and discarded: 6-29 % extra MMA work (W=32: dil 1 / 2 / 4 -> 34 / 36 / 40 columns) buys a 3-5x cut of the
// activation traffic, after which the weight stream (BLOCK_N x BK per tap) is what is left — and a CTA amortises it
// over MT = 1..4 accumulators (MT * 128 positions), chosen per launch so that the grid fills whole waves of 148 SMs
// (e.g. 512 -> 512 at 24 x 32 x 32: MT = 3, BLOCK_N = 128 -> 288 CTAs = 1.95 waves instead of 192 = 1.3).
//
// Same semantics as conv3x3_tc (two concat sources walked by the K loop, two dgrad destinations with accumulate flags,
// bias, fused BatchNorm batch statistics, or the eval-mode BatchNorm + LeakyReLU epilogue). Any W, H (no power-of-two
// boxes: the ragged 28 x 28 / 56 x 56 maps of 224^2 inputs are ordinary cases). Reference call site: nn.Conv2d in
// /root/reference/models/unet.py:188 and its autograd backward (dgrad = this kernel on the flipped, transposed pack).
#include <vector>

#include "pp_common.cuh"
#include "pp_ops.h"

namespace pp {

static constexpr int kRowsThreads = 320;   // TMA producer + MMA issuer + 8 epilogue warps (two per TMEM lane quarter)
static constexpr int kMaxBStages = 8;

struct RowsParams {
  int N, H, W, dil;
  int Wp;                      // W + dil: [dil zero columns | W pixels]; the next row's zeros are this row's right pad
  int R;                       // row-aligned items (pair mode): output rows per CTA (R * Wp <= MT * 128); 0: items are
                               // consecutive runs of MT * 128 positions of the image raster (start anywhere in a row)
  int items_per_img;           // ceil(H / R)  or  ceil(H * Wp / (MT * 128))
  int items_total;             // N * items_per_img (the grid may hold one padding CTA to complete a pair)
  int rbox;                    // rows of one A box
  int a_bytes;                 // bytes of one A stage (1 KB multiple)
  int kc0, kc1, c0, ctot;      // K chunks per source, channels of source 0, total input channels
  int nb;                      // B (weight) stages
  __nv_bfloat16* out0;
  __nv_bfloat16* out1;
  int outc0, outc1, acc0, acc1;
  const float* bias;
  double* stats;
  int imgs_per_group, groups;
  const float* ep_scale;
  const float* ep_shift;
  float ep_slope;
  int tma_store;               // 1: the epilogue stages the bf16 tile in shared memory and TMA-stores whole rows
  long long* trace;            // debug (PP_ROWS_TRACE=1): per CTA [start, first MMA, accumulators complete, end] clocks
};

// PAIR = 1: one CTA, M = 128 MMAs (cta_group::1). PAIR = 2: a cluster of two CTAs on one TPC, each with its own item
// (own activation box, own accumulators) and HALF of the weight tile; the leader issues M = 256 MMAs (cta_group::2).
template <int BLOCK_N, int BK, int MT, int PAIR>
__global__ void __launch_bounds__(kRowsThreads, 1)
conv3x3_rows_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                       const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmO0,
                       const __grid_constant__ CUtensorMap tmO1, const RowsParams p) {
  constexpr int ROW = BK * 2;                       // bytes of one pixel row of the A box
  constexpr int B_ROWS = BLOCK_N / PAIR;            // weight rows (output channels) held by this CTA
  constexpr int B_BYTES = B_ROWS * ROW;
  constexpr uint32_t SWZ = (BK == 64) ? SWZ_128B : SWZ_64B;
  constexpr uint32_t SBO = 8 * ROW;
  constexpr int TMEM_NEED = MT * BLOCK_N;
  constexpr int TMEM_COLS = TMEM_NEED <= 32 ? 32 : (TMEM_NEED <= 64 ? 64 : (TMEM_NEED <= 128 ? 128 : (TMEM_NEED <= 256 ? 256 : 512)));
  static_assert(TMEM_NEED <= 512, "accumulators do not fit TMEM");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_a = smem;                               // [2][a_bytes]
  uint8_t* s_b = smem + 2 * p.a_bytes;               // [nb][B_BYTES]
  uint64_t* a_full = reinterpret_cast<uint64_t*>(s_b + p.nb * B_BYTES);
  uint64_t* a_empty = a_full + 2;
  uint64_t* b_full = a_empty + 2;
  uint64_t* b_empty = b_full + kMaxBStages;
  uint64_t* tmem_full_bar = b_empty + kMaxBStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  float* s_stats = reinterpret_cast<float*>(tmem_slot + 2);   // [4 quarters][2][BLOCK_N]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tile = blockIdx.y;
  const long long t_start = p.trace != nullptr ? clock64() : 0;
  const uint32_t rank = PAIR == 2 ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  const int gitem = blockIdx.x;                      // consecutive items form a pair
  const bool live = gitem < p.items_total;           // a padding CTA loads out-of-range (zero) boxes and stores nothing
  const int img = live ? gitem / p.items_per_img : p.N;
  const int item = live ? gitem % p.items_per_img : 0;
  // first output row, offset of the first position inside that row, number of positions that belong to this CTA
  const int p_start = p.R > 0 ? item * p.R * p.Wp : item * (MT * 128);
  const int y0 = p.R > 0 ? item * p.R : p_start / p.Wp;
  const int item_row = p_start - y0 * p.Wp;          // 0 for row-aligned items (the pair shares ONE A descriptor)
  const int npos = p.R > 0 ? min(p.R, p.H - y0) * p.Wp : min(MT * 128, p.H * p.Wp - p_start);
  const int chunks = p.kc0 + p.kc1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    if (p.kc1 > 0) tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmB);
    if (p.tma_store) { tma_prefetch_desc(&tmO0); if (p.outc1 > 0) tma_prefetch_desc(&tmO1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < p.nb; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    mbar_init(tmem_full_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    if constexpr (PAIR == 2) tmem_alloc_2sm<TMEM_COLS>(tmem_slot);
    else tmem_alloc<TMEM_COLS>(tmem_slot);
  }
  tc_fence_before();
  if constexpr (PAIR == 2) cluster_sync_all();       // the peer's barriers are initialised before anything signals them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = warp_uniform(*tmem_slot);

  if (warp == 0) {
    if (elect_one()) {
      // ===== TMA producer: one activation box per K chunk (two stages), one weight tile per (K chunk, tap) =====
      // PAIR == 2: both CTAs load their own data; the bytes are accounted on the LEADER's full barriers.
      const uint32_t a_tx = static_cast<uint32_t>(p.rbox) * p.Wp * ROW * PAIR;
      auto load_a = [&](int kc) {
        const int sa = kc & 1;
        mbar_wait(&a_empty[sa], ((kc >> 1) & 1) ^ 1);
        if (leader) mbar_arrive_expect_tx(&a_full[sa], a_tx);
        const CUtensorMap* tm = kc < p.kc0 ? &tmA0 : &tmA1;
        const int kcoord = (kc < p.kc0 ? kc : kc - p.kc0) * BK;
        if constexpr (PAIR == 2) tma_load_4d_2sm(s_a + sa * p.a_bytes, tm, &a_full[sa], kcoord, -p.dil, y0 - p.dil, img);
        else tma_load_4d(s_a + sa * p.a_bytes, tm, &a_full[sa], kcoord, -p.dil, y0 - p.dil, img);
      };
      load_a(0);
      int stage = 0;
      uint32_t phase = 0;
      for (int kc = 0; kc < chunks; ++kc) {
        const int kofs = kc < p.kc0 ? kc * BK : p.c0 + (kc - p.kc0) * BK;
        for (int tap = 0; tap < 9; ++tap) {
          if (tap == 2 && kc + 1 < chunks) load_a(kc + 1);   // its stage was released when chunk kc-1 retired
          mbar_wait(&b_empty[stage], phase ^ 1);
          if (leader) mbar_arrive_expect_tx(&b_full[stage], B_BYTES * PAIR);
          if constexpr (PAIR == 2)
            tma_load_3d_2sm(s_b + stage * B_BYTES, &tmB, &b_full[stage], kofs, n_tile * BLOCK_N + int(rank) * B_ROWS, tap);
          else
            tma_load_3d(s_b + stage * B_BYTES, &tmB, &b_full[stage], kofs, n_tile * BLOCK_N, tap);
          if (++stage == p.nb) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (leader && elect_one()) {
      // ===== MMA issuer (the pair's leader) =====
      constexpr uint32_t idesc = make_idesc_bf16(128 * PAIR, BLOCK_N, 0, 0);
      constexpr uint32_t dhi = smem_desc_hi(SBO, SWZ);
      const uint32_t a_lo0 = smem_desc_lo(smem_u32(s_a), 16);
      const uint32_t b_lo0 = smem_desc_lo(smem_u32(s_b), 16);
      const uint32_t a_stage16 = static_cast<uint32_t>(p.a_bytes) >> 4;
      int stage = 0;
      uint32_t phase = 0;
      for (int kc = 0; kc < chunks; ++kc) {
        const int sa = kc & 1;
        mbar_wait(&a_full[sa], (kc >> 1) & 1);
        if (p.trace != nullptr && kc == 0) p.trace[(blockIdx.y * gridDim.x + blockIdx.x) * 4 + 1] = clock64() - t_start;
        const uint32_t a_base = a_lo0 + static_cast<uint32_t>(sa) * a_stage16;
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap) {
          // tap (ky, kx) of output position p reads box row p + ky*dil*Wp + kx*dil: one descriptor start offset
          const uint32_t shift = static_cast<uint32_t>((tap / 3) * p.dil * p.Wp + (tap % 3) * p.dil + item_row);
          const uint32_t a_tap = a_base + shift * (ROW >> 4);
          mbar_wait(&b_full[stage], phase);
          tc_fence_after();
          const uint32_t b_lo = b_lo0 + static_cast<uint32_t>(stage) * (B_BYTES >> 4);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
#pragma unroll
            for (int m = 0; m < MT; ++m) {
              if constexpr (PAIR == 2)
                umma_bf16_lohi_2sm(tmem_base + m * BLOCK_N, a_tap + m * (128 * ROW >> 4) + k * 2, dhi, b_lo + k * 2, dhi,
                                   idesc, (kc | tap | k) != 0 ? 1u : 0u);
              else
                umma_bf16_lohi(tmem_base + m * BLOCK_N, a_tap + m * (128 * ROW >> 4) + k * 2, dhi, b_lo + k * 2, dhi, idesc,
                               (kc | tap | k) != 0 ? 1u : 0u);
            }
          }
          if constexpr (PAIR == 2) umma_commit_2sm(&b_empty[stage]);
          else umma_commit(&b_empty[stage]);
          if (++stage == p.nb) { stage = 0; phase ^= 1; }
        }
        // the activation box of this chunk is free once its 9 taps have retired
        if constexpr (PAIR == 2) umma_commit_2sm(&a_empty[sa]);
        else umma_commit(&a_empty[sa]);
      }
      if constexpr (PAIR == 2) umma_commit_2sm(tmem_full_bar);
      else umma_commit(tmem_full_bar);
    }
  } else {
    // ===== epilogue: TMEM -> registers -> (+bias | BN-eval affine + LeakyReLU, +old) -> bf16 NHWC (+ BN partial sums) =====
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int col0 = n_tile * BLOCK_N;
    const int et = threadIdx.x - 64;
    const int cout = p.outc0 + p.outc1;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    if (p.trace != nullptr && threadIdx.x == 64) p.trace[(blockIdx.y * gridDim.x + blockIdx.x) * 4 + 2] = clock64() - t_start;
#pragma unroll 1
    for (int m = 0; m < MT; ++m) {
      const int pos = m * 128 + q * 32 + lane;
      const int g = item_row + pos;
      const int yl = g / p.Wp, px = g - yl * p.Wp;
      const bool valid = live && (pos < npos) && (px < p.W);
      const long long pix = (static_cast<long long>(img) * p.H + (y0 + yl)) * p.W + px;
      if (m * 128 >= npos && p.stats == nullptr && !p.tma_store) break;   // nothing of this accumulator is stored
#pragma unroll 1
      for (int c = half * 32; c < BLOCK_N; c += 64) {
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
          if ((reinterpret_cast<uintptr_t>(p.bias) & 15) == 0) {
            const float4* b4 = reinterpret_cast<const float4*>(p.bias + col);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 bv = __ldg(b4 + j);
              f[4 * j] += bv.x; f[4 * j + 1] += bv.y; f[4 * j + 2] += bv.z; f[4 * j + 3] += bv.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] += __ldg(p.bias + col + j);
          }
        }
        if (p.ep_scale != nullptr) {
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
        if (p.tma_store) {
          // Staging tile for the TMA store: per 64-channel group a dense [position][64 ch] box image with the 128-byte
          // swizzle (16-byte chunk index ^ position & 7), written here with conflict-free 16-byte shared stores. The
          // pipeline stages are idle by now (every MMA of BOTH CTAs has retired), so the tile lives there. A thread's
          // direct global stores would each be their own LSU wavefront (rows are Cout*2 bytes apart): 8x the cost.
          uint8_t* srow = smem + (c >> 6) * (MT * 128 * 128) + pos * 128;
          const int j0 = (c & 63) >> 3;   // first 16-byte chunk of this 32-column half inside the 64-channel row
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            Vec8<__nv_bfloat16> pk;
            float t[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) t[j] = f[g * 8 + j];
            pk.set(t);
            *reinterpret_cast<uint4*>(srow + (((j0 + g) ^ (pos & 7)) << 4)) = pk.raw;
            pk.get(t);   // statistics are taken of the bf16-ROUNDED values (what BatchNorm will read back)
#pragma unroll
            for (int j = 0; j < 8; ++j) f[g * 8 + j] = t[j];
          }
        } else if (valid) {
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
          s_stats[(q * 2 + 0) * BLOCK_N + c + lane] = s1[0];
          s_stats[(q * 2 + 1) * BLOCK_N + c + lane] = s2[0];
        }
      }
      if (p.stats != nullptr) {
        // a CTA's positions all belong to ONE image, hence to one statistics group
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (live) {
          const int grp = img / p.imgs_per_group;
          for (int i = et; i < 2 * BLOCK_N; i += 256) {
            const int st = i / BLOCK_N, j = i % BLOCK_N, cc = col0 + j;
            const double tot = (static_cast<double>(s_stats[(0 * 2 + st) * BLOCK_N + j]) + s_stats[(1 * 2 + st) * BLOCK_N + j]) +
                               (static_cast<double>(s_stats[(2 * 2 + st) * BLOCK_N + j]) + s_stats[(3 * 2 + st) * BLOCK_N + j]);
            atomicAdd(p.stats + ((static_cast<long long>((gitem * MT + m) % kStatReplicas) * p.groups + grp) * cout + cc) * 2 + st, tot);
          }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
    }
    if (p.tma_store) {
      // whole rows of this CTA go out as one box per 64-channel group: {64 ch, Wp px from x = 0, R rows}; the pad
      // columns x >= W and the rows >= H are out of bounds of the output tensor and are clipped by the TMA unit
      fence_proxy_async_smem();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (threadIdx.x == 64 && live) {
#pragma unroll 1
        for (int cg = 0; cg < BLOCK_N / 64; ++cg) {
          const int col = col0 + cg * 64;
          if (col < p.outc0) tma_store_4d(&tmO0, smem + cg * (MT * 128 * 128), col, 0, y0, img);
          else tma_store_4d(&tmO1, smem + cg * (MT * 128 * 128), col - p.outc0, 0, y0, img);
        }
        tma_store_commit();
        tma_store_wait_read();   // the staging tile must outlive the bulk reads
      }
    }
    tc_fence_before();
    if (p.trace != nullptr && threadIdx.x == 64) p.trace[(blockIdx.y * gridDim.x + blockIdx.x) * 4 + 3] = clock64() - t_start;
  }
  __syncwarp();
  if constexpr (PAIR == 2) cluster_sync_all();       // neither CTA's shared memory / TMEM goes away under the other's MMAs
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if constexpr (PAIR == 2) tmem_dealloc_2sm<TMEM_COLS>(tmem_base);
    else tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// ----------------------------------------------------------------------------------------------
// host
// ----------------------------------------------------------------------------------------------
static constexpr int kRowsTail = 1024 + 512;   // alignment slack + barriers + TMEM slot (statistics scratch added per BLOCK_N)

// Candidate tilings of one launch, cheapest first by the cost model below; conv_tc.cu's autotuner times the best few.
//   R > 0: row-aligned items (required in pair mode: the two CTAs share one A descriptor); R == 0: items are runs of
//   MT*128 consecutive raster positions (no partially used accumulators except the image's last one).
// Cost = waves x per-CTA time, a CTA's time being the largest of
//   * its tensor-pipe cycles (MT*128 x BLOCK_N x 9*Cin MACs at 4096 MAC/cycle/SM, incl. the discarded pad positions),
//   * its shared-memory cycles: every 128 x N x 16 MMA reads 4 KB of A and 32*N bytes of B (half of that per CTA in pair
//     mode) and the TMA writes land there too, at ~120 B/cycle usable of the 128 B/cycle/SM — the bound for N <= 256
//     single-CTA tiles (128 x 128: 128 B/cycle of operand reads alone),
//   * its L2 -> shared-memory fetch cycles (~40 B/cycle/SM with every SM streaming),
// plus a fixed prologue and the epilogue. PP_CONV_ROWS_PAIR=0 / 1 forbids / forces pair mode (A/B experiments).
int conv3x3_rows_plans(int N, int H, int W, int dil, int C0, int C1, int cout, RowsPlan* out, int max_out) {
  static const int kTiles[3] = {256, 128, 64};
  static const int pair_mode = [] { const char* e = getenv("PP_CONV_ROWS_PAIR"); return e ? atoi(e) : -1; }();
  const int Wp = W + dil;
  const int ctot = C0 + C1;
  const bool k64 = (C0 % 64 == 0) && (C1 % 64 == 0);
  const double sms = sm_count();
  if (Wp > 256 || C0 % 32 != 0 || C1 % 32 != 0 || cout % 64 != 0) return 0;
  RowsPlan all[96];
  int n = 0;
  for (int pair = 1; pair <= 2; ++pair) {
    if (pair_mode == 0 && pair == 2) continue;
    if (pair_mode == 1 && pair == 1) continue;
    for (int t = 0; t < 3; ++t) {
      const int bn = kTiles[t];
      if (cout % bn != 0) continue;
      for (int mt = 1; mt <= 4; ++mt) {
        if (mt * bn > 512) break;
       for (int aligned = (pair == 2 ? 1 : 0); aligned <= 1; ++aligned) {
        int R = 0, items_img;
        if (aligned) {   // whole rows per CTA: required by pair mode, and what the TMA-store epilogue needs
          R = (mt * 128) / Wp;
          if (R < 1) continue;
          if (R > H) R = H;
          R = ceil_div(H, ceil_div(H, R));   // same item count, evenly sized items
          items_img = ceil_div(H, R);
        } else {
          items_img = ceil_div(H * Wp, mt * 128);
        }
        for (int bk = 64; bk >= 32; bk -= 32) {
          if (bk == 64 && !k64) continue;
          const int rbox = ceil_div(mt * 128 + (R > 0 ? 0 : Wp - 1) + 2 * dil * Wp + 2 * dil, Wp);
          if (rbox > 256) continue;
          const int a_bytes = (rbox * Wp * bk * 2 + 1023) / 1024 * 1024;
          const int b_bytes = (bn / pair) * bk * 2;
          int nb = (225 * 1024 - kRowsTail - 8 * bn * 4 - 2 * a_bytes) / b_bytes;
          if (nb > kMaxBStages) nb = kMaxBStages;
          if (nb < 3) continue;
          const int items = N * items_img;
          const double ctas = static_cast<double>(pair == 2 ? (items + 1) / 2 * 2 : items) * (cout / bn);
          const double waves = ceil(ctas / sms);
          const double n_mma = static_cast<double>(mt) * 9.0 * ctot / 16.0;
          const double mma = n_mma * bn / 2.0;
          const double tma_bytes = static_cast<double>(rbox) * Wp * ctot * 2.0 + 9.0 * ctot * (bn / pair) * 2.0;
          const double smem_cyc = (n_mma * (4096.0 + 32.0 * bn / pair) + tma_bytes) / 120.0;
          const double l2 = tma_bytes / 40.0;
          double body = mma > smem_cyc ? mma : smem_cyc;
          if (l2 > body) body = l2;
          if (bk == 32) body *= 1.1;   // twice the barrier round trips per byte
          // epilogue: direct global stores cost one LSU wavefront per 16 bytes; the TMA-store path ~1/4 of that
          const double epi = mt * 128.0 * bn * 2.0 / 16.0 * (R > 0 ? 0.3 : 1.0) + mt * (bn / 32) * 60.0;
          const double cta = body + 6000.0 + epi + (pair == 2 ? 1500.0 : 0.0);
          RowsPlan pl;
          pl.cost = waves * cta; pl.block_n = bn; pl.bk = bk; pl.mt = mt; pl.pair = pair; pl.R = R;
          pl.rbox = rbox; pl.a_bytes = a_bytes; pl.nb = nb;
          pl.smem = 2 * a_bytes + nb * b_bytes + kRowsTail + 8 * bn * 4;
          if (n < 96) all[n++] = pl;
        }
       }
      }
    }
  }
  for (int i = 1; i < n; ++i)   // insertion sort by cost
    for (int j = i; j > 0 && all[j].cost < all[j - 1].cost; --j) { RowsPlan t = all[j]; all[j] = all[j - 1]; all[j - 1] = t; }
  // the cheapest plan of each (pair, aligned, BLOCK_N) class first, so that a short list spans different regimes
  int m = 0;
  for (int pass = 0; pass < 2 && m < max_out; ++pass)
    for (int i = 0; i < n && m < max_out; ++i) {
      bool first_of_class = true;
      for (int j = 0; j < i; ++j)
        if (all[j].pair == all[i].pair && all[j].block_n == all[i].block_n && (all[j].R > 0) == (all[i].R > 0))
          first_of_class = false;
      if ((pass == 0) == first_of_class) out[m++] = all[i];
    }
  return m;
}

bool conv3x3_rows_applicable(int C0, int C1, int cout, int N, int H, int W, int dil) {
  static const int on = [] { const char* e = getenv("PP_CONV_ROWS"); return (e && e[0] == '0') ? 0 : 1; }();
  if (!on) return false;
  if (H * (W + dil) < 256) return false;   // tiny maps: the generic kernel packs several images into one 128-pixel tile
  if (W > 96) return false;                // full-resolution rows: the halo / generic kernels (128-wide: no gain measured)
  RowsPlan pl;
  return conv3x3_rows_plans(N, H, W, dil, C0, C1, cout, &pl, 1) > 0;
}

template <int BLOCK_N, int BK, int MT, int PAIR>
static int launch_rows(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const CUtensorMap& o0,
                       const CUtensorMap& o1, const RowsParams& p, dim3 grid, int smem, double flops, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    PP_CHECK_CUDA(cudaFuncSetAttribute(conv3x3_rows_tc_kernel<BLOCK_N, BK, MT, PAIR>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(kRowsThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = PAIR;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  static const int trace_on = [] { const char* e = getenv("PP_ROWS_TRACE"); return (e && e[0] == '1') ? 1 : 0; }();
  if (trace_on) {   // debug only: allocates, synchronises and prints the mean phase lengths of this launch
    RowsParams pt = p;
    const int ctas = grid.x * grid.y;
    PP_CHECK_CUDA(cudaMalloc(&pt.trace, sizeof(long long) * 4 * ctas));
    PP_CHECK_CUDA(cudaMemsetAsync(pt.trace, 0, sizeof(long long) * 4 * ctas, stream));
    PP_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv3x3_rows_tc_kernel<BLOCK_N, BK, MT, PAIR>, a0, a1, b, o0, o1, pt));
    std::vector<long long> h(4 * ctas);
    PP_CHECK_CUDA(cudaStreamSynchronize(stream));
    PP_CHECK_CUDA(cudaMemcpy(h.data(), pt.trace, sizeof(long long) * 4 * ctas, cudaMemcpyDeviceToHost));
    cudaFree(pt.trace);
    double s1 = 0, s2 = 0, s3 = 0; int n1 = 0, n = 0;
    for (int i = 0; i < ctas; ++i) {
      if (h[4 * i + 1] > 0) { s1 += h[4 * i + 1]; ++n1; }
      if (h[4 * i + 3] > 0) { s2 += h[4 * i + 2]; s3 += h[4 * i + 3]; ++n; }
    }
    fprintf(stderr, "rows trace BLOCK_N=%d MT=%d PAIR=%d grid=%dx%d: first MMA at %.0f, accumulators complete at %.0f, "
                    "epilogue done at %.0f cycles (means)\n", BLOCK_N, MT, PAIR, grid.x, grid.y, n1 ? s1 / n1 : 0.0,
            n ? s2 / n : 0.0, n ? s3 / n : 0.0);
    PP_LAUNCH_CHECK();
    return PP_OK;
  }
  const int slot = prof_begin(PROF_CONV, flops, stream);
  PP_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv3x3_rows_tc_kernel<BLOCK_N, BK, MT, PAIR>, a0, a1, b, o0, o1, p));
  prof_end(slot, stream);
  PP_LAUNCH_CHECK();
  return PP_OK;
}

int conv3x3_rows_tc(const void* x0, int C0, const void* x1, int C1, const void* wpack, const float* bias, void* out0,
                    int outc0, int acc0, void* out1, int outc1, int acc1, int N, int H, int W, int dil,
                    cudaStream_t stream, double* stats, int groups, const ConvAffine* affine, const RowsPlan* plan) {
  const int cout = outc0 + outc1, ctot = C0 + C1;
  RowsPlan pl;
  if (plan != nullptr) pl = *plan;
  else PP_REQUIRE(conv3x3_rows_plans(N, H, W, dil, C0, C1, cout, &pl, 1) > 0, "conv3x3_rows_tc: no tile plan for this shape");
  RowsParams p{};
  p.N = N; p.H = H; p.W = W; p.dil = dil;
  p.Wp = W + dil;
  p.R = pl.R;
  p.items_per_img = pl.R > 0 ? ceil_div(H, pl.R) : ceil_div(H * p.Wp, pl.mt * 128);
  p.items_total = N * p.items_per_img;
  p.rbox = pl.rbox; p.a_bytes = pl.a_bytes; p.nb = pl.nb;
  p.kc0 = C0 / pl.bk; p.kc1 = C1 / pl.bk; p.c0 = C0; p.ctot = ctot;
  p.out0 = static_cast<__nv_bfloat16*>(out0); p.out1 = static_cast<__nv_bfloat16*>(out1);
  p.outc0 = outc0; p.outc1 = outc1; p.acc0 = acc0; p.acc1 = acc1; p.bias = bias;
  p.stats = stats;
  p.groups = groups > 0 ? groups : 1;
  p.imgs_per_group = N / p.groups;
  if (stats != nullptr) PP_REQUIRE(N % p.groups == 0, "conv3x3_rows_tc: %d images do not split into %d groups", N, p.groups);
  if (affine != nullptr) { p.ep_scale = affine->scale; p.ep_shift = affine->shift; p.ep_slope = affine->slope; }

  CUtensorMap a0, a1, b;
  int rc = encode_tmap_nhwc(&a0, x0, N, H, W, C0, pl.bk, p.Wp, pl.rbox, 1, pl.bk == 64);
  if (rc) return rc;
  if (C1 > 0) rc = encode_tmap_nhwc(&a1, x1, N, H, W, C1, pl.bk, p.Wp, pl.rbox, 1, pl.bk == 64);
  else a1 = a0;
  if (rc) return rc;
  rc = encode_tmap_weights(&b, wpack, 9, cout, ctot, pl.bk, pl.block_n / pl.pair, pl.bk == 64);
  if (rc) return rc;
  // TMA-store epilogue: row-aligned items, plain (non-accumulating) destinations made of whole 64-channel groups, and a
  // staging tile (MT*128 positions x BLOCK_N channels of bf16) that fits the idle pipeline stages
  static const int tma_store_on = [] { const char* e = getenv("PP_ROWS_TMA_STORE"); return (e && e[0] == '0') ? 0 : 1; }();
  CUtensorMap o0 = a0, o1 = a0;
  p.tma_store = tma_store_on && pl.R > 0 && acc0 == 0 && acc1 == 0 && outc0 % 64 == 0 && outc1 % 64 == 0 &&
                pl.mt * 128 * pl.block_n * 2 <= 2 * pl.a_bytes + pl.nb * (pl.block_n / pl.pair) * pl.bk * 2;
  if (p.tma_store) {
    rc = encode_tmap_nhwc(&o0, out0, N, H, W, outc0, 64, p.Wp, pl.R, 1, true);
    if (rc) return rc;
    if (outc1 > 0) rc = encode_tmap_nhwc(&o1, out1, N, H, W, outc1, 64, p.Wp, pl.R, 1, true);
    if (rc) return rc;
  }
  const int gx = pl.pair == 2 ? (p.items_total + 1) / 2 * 2 : p.items_total;
  const dim3 grid(gx, cout / pl.block_n);
  const double flops = 2.0 * N * H * W * 9.0 * ctot * cout;
  static const int debug = [] { const char* e = getenv("PP_CONV_ROWS_DEBUG"); return (e && e[0] == '1') ? 1 : 0; }();
  if (debug)
    fprintf(stderr, "conv_rows: N=%d %dx%d C=%d+%d->%d dil=%d: BLOCK_N=%d BK=%d MT=%d PAIR=%d R=%d rbox=%d a_bytes=%d nb=%d "
                    "smem=%d grid=%dx%d\n", N, H, W, C0, C1, cout, dil, pl.block_n, pl.bk, pl.mt, pl.pair, pl.R, pl.rbox,
            pl.a_bytes, pl.nb, pl.smem, grid.x, grid.y);
#define PP_ROWS_CASE(BN_, BK_, MT_) \
  if (pl.block_n == BN_ && pl.bk == BK_ && pl.mt == MT_) {                                                   \
    if (pl.pair == 2) return launch_rows<BN_, BK_, MT_, 2>(a0, a1, b, o0, o1, p, grid, pl.smem, flops, stream); \
    return launch_rows<BN_, BK_, MT_, 1>(a0, a1, b, o0, o1, p, grid, pl.smem, flops, stream);                   \
  }
#define PP_ROWS_CASES(BK_)                                                                             \
  PP_ROWS_CASE(256, BK_, 1) PP_ROWS_CASE(256, BK_, 2) PP_ROWS_CASE(128, BK_, 1) PP_ROWS_CASE(128, BK_, 2) \
  PP_ROWS_CASE(128, BK_, 3) PP_ROWS_CASE(128, BK_, 4) PP_ROWS_CASE(64, BK_, 1) PP_ROWS_CASE(64, BK_, 2)    \
  PP_ROWS_CASE(64, BK_, 3) PP_ROWS_CASE(64, BK_, 4)
  PP_ROWS_CASES(64)
  PP_ROWS_CASES(32)
#undef PP_ROWS_CASES
#undef PP_ROWS_CASE
  set_error("conv3x3_rows_tc: no kernel for block_n=%d bk=%d mt=%d", pl.block_n, pl.bk, pl.mt);
  return PP_ERR_INVALID;
}

}  // namespace pp
