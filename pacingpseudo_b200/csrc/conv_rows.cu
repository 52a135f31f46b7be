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
#include "pp_common.cuh"
#include "pp_ops.h"

namespace pp {

static constexpr int kRowsThreads = 320;   // TMA producer + MMA issuer + 8 epilogue warps (two per TMEM lane quarter)
static constexpr int kMaxBStages = 8;

struct RowsParams {
  int N, H, W, dil;
  int Wp;                      // W + 2 * dil
  int img_pos;                 // H * Wp output positions per image (incl. the discarded pad columns)
  int items_per_img;           // ceil(img_pos / (MT * 128))
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
};

template <int BLOCK_N, int BK, int MT>
__global__ void __launch_bounds__(kRowsThreads, 1)
conv3x3_rows_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                       const __grid_constant__ CUtensorMap tmB, const RowsParams p) {
  constexpr int ROW = BK * 2;                       // bytes of one pixel row of the A box
  constexpr int B_BYTES = BLOCK_N * ROW;
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
  const int img = blockIdx.x / p.items_per_img;
  const int item = blockIdx.x % p.items_per_img;
  const int p_start = item * (MT * 128);             // first output position of this CTA inside the image raster
  const int y0 = p_start / p.Wp;                     // first output row touched
  const int chunks = p.kc0 + p.kc1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    if (p.kc1 > 0) tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < 2; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < p.nb; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
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
      // ===== TMA producer: one activation box per K chunk (two stages), one weight tile per (K chunk, tap) =====
      const uint32_t a_tx = static_cast<uint32_t>(p.rbox) * p.Wp * ROW;
      auto load_a = [&](int kc) {
        const int sa = kc & 1;
        mbar_wait(&a_empty[sa], ((kc >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&a_full[sa], a_tx);
        if (kc < p.kc0) tma_load_4d(s_a + sa * p.a_bytes, &tmA0, &a_full[sa], kc * BK, -p.dil, y0 - p.dil, img);
        else tma_load_4d(s_a + sa * p.a_bytes, &tmA1, &a_full[sa], (kc - p.kc0) * BK, -p.dil, y0 - p.dil, img);
      };
      load_a(0);
      int stage = 0;
      uint32_t phase = 0;
      for (int kc = 0; kc < chunks; ++kc) {
        const int kofs = kc < p.kc0 ? kc * BK : p.c0 + (kc - p.kc0) * BK;
        for (int tap = 0; tap < 9; ++tap) {
          if (tap == 2 && kc + 1 < chunks) load_a(kc + 1);   // its stage was released when chunk kc-1 retired
          mbar_wait(&b_empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&b_full[stage], B_BYTES);
          tma_load_3d(s_b + stage * B_BYTES, &tmB, &b_full[stage], kofs, n_tile * BLOCK_N, tap);
          if (++stage == p.nb) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // ===== MMA issuer =====
      constexpr uint32_t idesc = make_idesc_bf16(128, BLOCK_N, 0, 0);
      constexpr uint32_t dhi = smem_desc_hi(SBO, SWZ);
      const uint32_t a_lo0 = smem_desc_lo(smem_u32(s_a), 16);
      const uint32_t b_lo0 = smem_desc_lo(smem_u32(s_b), 16);
      const uint32_t a_stage16 = static_cast<uint32_t>(p.a_bytes) >> 4;
      // first box row of the item's first position: (p_start - y0 * Wp) rows into the box (x offset inside row y0)
      const uint32_t item_row = static_cast<uint32_t>(p_start - y0 * p.Wp);
      int stage = 0;
      uint32_t phase = 0;
      for (int kc = 0; kc < chunks; ++kc) {
        const int sa = kc & 1;
        mbar_wait(&a_full[sa], (kc >> 1) & 1);
        const uint32_t a_base = a_lo0 + static_cast<uint32_t>(sa) * a_stage16;
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap) {
          const uint32_t shift = static_cast<uint32_t>((tap / 3) * p.dil * p.Wp + (tap % 3) * p.dil) + item_row;
          const uint32_t a_tap = a_base + shift * (ROW >> 4);
          mbar_wait(&b_full[stage], phase);
          tc_fence_after();
          const uint32_t b_lo = b_lo0 + static_cast<uint32_t>(stage) * (B_BYTES >> 4);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
#pragma unroll
            for (int m = 0; m < MT; ++m)
              umma_bf16_lohi(tmem_base + m * BLOCK_N, a_tap + m * (128 * ROW >> 4) + k * 2, dhi, b_lo + k * 2, dhi, idesc,
                             (kc | tap | k) != 0 ? 1u : 0u);
          }
          umma_commit(&b_empty[stage]);
          if (++stage == p.nb) { stage = 0; phase ^= 1; }
        }
        umma_commit(&a_empty[sa]);   // the activation box of this chunk is free once its 9 taps have retired
      }
      umma_commit(tmem_full_bar);
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
#pragma unroll 1
    for (int m = 0; m < MT; ++m) {
      const int pos = p_start + m * 128 + q * 32 + lane;
      const int py = pos / p.Wp, px = pos - py * p.Wp;
      const bool valid = (px < p.W) && (py < p.H);
      const long long pix = (static_cast<long long>(img) * p.H + py) * p.W + px;
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
        const int grp = img / p.imgs_per_group;
        for (int i = et; i < 2 * BLOCK_N; i += 256) {
          const int st = i / BLOCK_N, j = i % BLOCK_N, cc = col0 + j;
          const double tot = (static_cast<double>(s_stats[(0 * 2 + st) * BLOCK_N + j]) + s_stats[(1 * 2 + st) * BLOCK_N + j]) +
                             (static_cast<double>(s_stats[(2 * 2 + st) * BLOCK_N + j]) + s_stats[(3 * 2 + st) * BLOCK_N + j]);
          atomicAdd(p.stats + ((static_cast<long long>((blockIdx.x * MT + m) % kStatReplicas) * p.groups + grp) * cout + cc) * 2 + st, tot);
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

// ----------------------------------------------------------------------------------------------
// host
// ----------------------------------------------------------------------------------------------
struct RowsPlan {
  int block_n, bk, mt, rbox, a_bytes, nb, smem;
  double cost;
};

static int rows_rbox(int W, int dil, int mt) {
  const int Wp = W + 2 * dil;
  return ceil_div(mt * 128 + Wp - 1 + 2 * dil * Wp + 2 * dil, Wp);
}

static constexpr int kRowsTail = 1024 + 512;   // alignment slack + barriers + TMEM slot (statistics scratch added per BLOCK_N)

// Picks (BLOCK_N, BK, MT): minimise waves x (per-CTA time), where a CTA's time is the larger of its tensor-pipe cycles
// and its L2 -> shared-memory fetch cycles (~40 B/cycle/SM with every SM streaming) plus prologue / epilogue.
static bool rows_plan(int N, int H, int W, int dil, int C0, int C1, int cout, RowsPlan* best) {
  static const int kTiles[3] = {256, 128, 64};
  const int Wp = W + 2 * dil;
  const int img_pos = H * Wp;
  const int ctot = C0 + C1;
  const bool k64 = (C0 % 64 == 0) && (C1 % 64 == 0);
  const double sms = sm_count();
  best->cost = 1e300;
  for (int t = 0; t < 3; ++t) {
    const int bn = kTiles[t];
    if (cout % bn != 0) continue;
    for (int mt = 1; mt <= 4; ++mt) {
      if (mt * bn > 512) break;
      for (int bk = 64; bk >= 32; bk -= 32) {
        if (bk == 64 && !k64) continue;
        const int rbox = rows_rbox(W, dil, mt);
        if (rbox > 256 || Wp > 256) continue;
        const int a_bytes = (rbox * Wp * bk * 2 + 1023) / 1024 * 1024;
        const int b_bytes = bn * bk * 2;
        int nb = (225 * 1024 - kRowsTail - 8 * bn * 4 - 2 * a_bytes) / b_bytes;
        if (nb > kMaxBStages) nb = kMaxBStages;
        if (nb < 3) continue;
        const int items = N * ceil_div(img_pos, mt * 128);
        const double ctas = static_cast<double>(items) * (cout / bn);
        const double waves = ceil(ctas / sms);
        const double mma = static_cast<double>(mt) * 128.0 * bn * 9.0 * ctot / 4096.0;
        const double l2 = (static_cast<double>(rbox) * Wp * ctot * 2.0 + 9.0 * ctot * bn * 2.0) / 40.0;
        const double issue = (bk == 32 ? 1.15 : 1.0);   // twice the barrier round trips per byte
        const double cta = (mma > l2 ? mma : l2) * issue + 5000.0 + mt * ((bn + 63) / 64) * 450.0;
        const double cost = waves * cta;
        if (cost < best->cost) {
          best->cost = cost; best->block_n = bn; best->bk = bk; best->mt = mt; best->rbox = rbox; best->a_bytes = a_bytes;
          best->nb = nb; best->smem = 2 * a_bytes + nb * b_bytes + kRowsTail + 8 * bn * 4;
        }
      }
    }
  }
  return best->cost < 1e299;
}

bool conv3x3_rows_applicable(int C0, int C1, int cout, int N, int H, int W, int dil) {
  static const int on = [] { const char* e = getenv("PP_CONV_ROWS"); return (e && e[0] == '0') ? 0 : 1; }();
  if (!on) return false;
  if (C0 % 32 != 0 || C1 % 32 != 0 || cout % 64 != 0) return false;
  const int Wp = W + 2 * dil;
  if (H * Wp < 256) return false;   // tiny maps: the generic kernel packs several images into one 128-pixel tile
  if (W > 96) return false;         // full-resolution rows: the halo / generic kernels
  RowsPlan pl;
  return rows_plan(N, H, W, dil, C0, C1, cout, &pl);
}

template <int BLOCK_N, int BK, int MT>
static int launch_rows(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const RowsParams& p, dim3 grid,
                       int smem, double flops, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    PP_CHECK_CUDA(cudaFuncSetAttribute(conv3x3_rows_tc_kernel<BLOCK_N, BK, MT>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const int slot = prof_begin(PROF_CONV, flops, stream);
  conv3x3_rows_tc_kernel<BLOCK_N, BK, MT><<<grid, kRowsThreads, smem, stream>>>(a0, a1, b, p);
  prof_end(slot, stream);
  PP_LAUNCH_CHECK();
  return PP_OK;
}

int conv3x3_rows_tc(const void* x0, int C0, const void* x1, int C1, const void* wpack, const float* bias, void* out0,
                    int outc0, int acc0, void* out1, int outc1, int acc1, int N, int H, int W, int dil,
                    cudaStream_t stream, double* stats, int groups, const ConvAffine* affine) {
  const int cout = outc0 + outc1, ctot = C0 + C1;
  RowsPlan pl;
  PP_REQUIRE(rows_plan(N, H, W, dil, C0, C1, cout, &pl), "conv3x3_rows_tc: no tile plan for this shape");
  RowsParams p{};
  p.N = N; p.H = H; p.W = W; p.dil = dil;
  p.Wp = W + 2 * dil;
  p.img_pos = H * p.Wp;
  p.items_per_img = ceil_div(p.img_pos, pl.mt * 128);
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
  rc = encode_tmap_weights(&b, wpack, 9, cout, ctot, pl.bk, pl.block_n, pl.bk == 64);
  if (rc) return rc;
  const dim3 grid(N * p.items_per_img, cout / pl.block_n);
  const double flops = 2.0 * N * H * W * 9.0 * ctot * cout;
#define PP_ROWS_CASE(BN_, BK_, MT_) \
  if (pl.block_n == BN_ && pl.bk == BK_ && pl.mt == MT_) \
    return launch_rows<BN_, BK_, MT_>(a0, a1, b, p, grid, pl.smem, flops, stream);
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
