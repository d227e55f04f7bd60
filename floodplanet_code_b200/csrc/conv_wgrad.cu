// Weight gradient of the 3x3 / pad 1 convolution on the sm_100a tensor cores.
// (autograd of nn.Conv2d in st_water_seg/models/unet.py:14,16)
//
//   dW[co, (r,s), ci] = sum_{n,h,w} dy[n,h,w,co] * x[n, h+r-1, w+s-1, ci]
//
// is a GEMM whose reduction dimension is the PIXEL axis (up to 16.8 M long) and whose
// output is tiny, so it is split-K: every CTA owns one output tile and one contiguous range
// of 64-pixel tiles (4 image rows x 16 pixels), accumulates in TMEM, and writes fp32
// partials that a second kernel sums deterministically into the OIHW fp32 gradient.
//
// Both GEMM operands are "MN-major": in NHWC memory the channel axis (the GEMM M resp. N
// axis) is the contiguous one and the pixel axis (GEMM K) is strided.  TMA drops pixel boxes
// into 128B-swizzled smem (one 128-byte row per pixel, 64 channels) and tcgen05.mma consumes
// them directly with the transposed-operand bits of the instruction descriptor set -- no
// transposes anywhere.  One K step of an MMA is one image row of the tile (16 pixels = two
// 8-row swizzle groups).
//
// The 3x3 taps are shifted views of ONE box that carries the halo (zero filled by TMA outside
// the image): a tap is only a different descriptor start address, exactly as in the fprop
// kernel, so each operand byte enters shared memory once per tile instead of once per tap.
//
// Operand arrangements (every layer of the UNet keeps the tensor cores on full-height M = 128 MMAs
// except the first one, which is shared-memory bound either way):
//   MODE_X_SHIFT  (0, Cout % 128 == 0): A = 128 output channels of dy (no halo); B = one filter
//                  row of x, box 18 px wide, the three taps s stacked on N (N = 192 per 64-channel
//                  block of x, the N blocks one pixel apart), so dy is read once per K step and
//                  channel block instead of once per tap (24 -> 20 KB of operand reads per K step
//                  at NB = 2, which sat exactly on the 128 B/clk shared-memory limit: +3 %).
//                  Work item = (128 co, 64*NB ci, filter row r): NB accumulators of 192 cols.
//   MODE_POINTWISE (2, 1x1 convolution, the late-fusion concat_convs): no halo, one tap:
//                  A = 128 output channels of dy, B = NB blocks of 64 input channels of x.
//   MODE_RS_SPLIT (4, Cout == 64 blocks, Cin % 64 == 0): the vertical tap offset is carried by dy and
//                  the horizontal one by x, dW[(r,s)] = sum_q dy[q - (r-1, 0)] x[q + (0, s-1)]:
//                  A = dy box with a vertical halo (16 x 6 px), two filter rows stacked on M (second
//                  64-row block LBO = one box row further), B = x box with a horizontal halo
//                  (18 x 4 px), the three taps s stacked on N (N = 192, blocks one pixel apart).
//                  One MMA covers 6 taps: 2 MMAs and 20 KB of operand reads per K step.  (Its
//                  predecessor shifted only dy and stacked tap PAIRS on M: 5 MMAs and 30 KB per K
//                  step, shared-memory bound at 60 % tensor pipe; 64->64 @512^2 1.24 -> 1.15 ms.)
//   MODE_X_STACK  (3, Cout == 64 blocks with < 64 input channels, i.e. the first layer): with N = 16
//                  an M128 MMA is 8 tensor cycles of work for 4.5 KB of operands (the tap-pair
//                  arrangement re-read the dy operand for each of its 20 MMAs per stage: 90 KB,
//                  17 % tensor-pipe utilisation, ncu).  Here A = dy unshifted (M = 64), B = x box
//                  with halo, and the three taps of a filter row are stacked on N (N blocks one
//                  pixel apart, LBO = one pixel row): 12 MMAs and 42 KB per stage, 0.905 -> 0.652 ms.
#include <stdlib.h>
#include "host_common.h"
#include "ptx.cuh"

namespace fp {

constexpr int kWgThreads = 192;
constexpr int kWgTW = 16, kWgTH = 4;        // pixel tile: 4 image rows x 16 pixels
constexpr int kWgBK = kWgTW * kWgTH;        // 64 pixels per pipeline stage
constexpr int kWgBoxW = kWgTW + 2;

struct WgradParams {
  int N, H, W;
  int tiles_w, tiles_h;  // per image
  int num_pix_tiles;
  int ksplit;
  int n_items;      // output tiles
  int items_ci;     // number of ci groups (item = co_grp * items_ci * items_r + ci_grp * items_r + r)
  int items_r;      // 3 in MODE_X_SHIFT, 1 otherwise
  int Cout, Cin;    // Cin = padded input channels (layout of the partials)
  int taps;         // 9 (3x3) or 1 (pointwise)
  float* ws;        // [ksplit][Cout][taps][Cin]
};

template <int MODE, int NBW, int NB>
struct WgCfg {
  // A = dy.  modes 0, 2: two [64 px][64 co] blocks; mode 3: one; mode 4: box 16 x 6 px (vertical halo)
  static constexpr int kABlock = MODE == 4 ? kWgTW * (kWgTH + 2) * 128 : kWgBK * 128;
  static constexpr int kABytes = (MODE == 0 || MODE == 2) ? 2 * kABlock : kABlock;          // TMA bytes
  static constexpr int kASlot = (kABytes + 1023) / 1024 * 1024;
  // B = x.  modes 0, 4: (NB blocks of) one filter row with a horizontal halo [18 x 4 px][64 ci];
  // mode 2: NB blocks [64 px][64 ci]; mode 3: haloed box [18 x 6 px][NBW ci]
  static constexpr int kBBlock = (MODE == 0 || MODE == 4) ? kWgBoxW * kWgTH * 128
                               : (MODE == 3 ? kWgBoxW * (kWgTH + 2) * NBW * 2 : kWgBK * NBW * 2);
  static constexpr int kBBytes = (MODE == 0 || MODE == 2) ? NB * kBBlock : kBBlock;
  static constexpr int kBSlot = (kBBytes + 1023) / 1024 * 1024;
  static constexpr int kStageBytes = kASlot + kBSlot;
  static constexpr int kTxBytes = kABytes + kBBytes;
  static constexpr int kStagesRaw = (200 * 1024) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 6 ? 6 : kStagesRaw;
  // mode 0: one group per 64-channel block of x, the three taps of the filter row stacked on N
  static constexpr int kGroups = MODE == 0 ? NB : (MODE == 3 ? 3 : (MODE == 4 ? 2 : 1));
  static constexpr int kN = (MODE == 0 || MODE == 3 || MODE == 4) ? 3 * NBW : NBW * NB;   // UMMA N per group
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;
};

template <int MODE, int NBW, int NB, bool ACOLL = false>
__global__ void __launch_bounds__(kWgThreads, 1)
conv3x3_wgrad_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX,
                     const WgradParams p) {
  static_assert(MODE == 0 || MODE == 2 || MODE == 3 || MODE == 4, "unknown wgrad mode");
  using Cfg = WgCfg<MODE, NBW, NB>;
  constexpr int kStages = Cfg::kStages;
  constexpr uint32_t kIdesc = make_idesc_bf16(MODE == 3 ? 64 : 128, Cfg::kN, 1, 1);
  constexpr uint32_t kBRow = (MODE == 0 || MODE == 2 || MODE == 4) ? 128 : NBW * 2;   // bytes per pixel row of B
  constexpr uint32_t kBSwz = kBRow;
  constexpr uint32_t kBSBO = 8 * kBRow;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + kStages * Cfg::kStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 64u + 8u * s; };
  const uint32_t accum_bar = bar_base + 128u;
  const uint32_t tmem_slot = bar_base + 160u;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // work item
  const int item = blockIdx.x % p.n_items;
  const int split = blockIdx.x / p.n_items;
  const int r_idx = item % p.items_r;
  const int ci_grp = (item / p.items_r) % p.items_ci;
  const int co_grp = item / (p.items_r * p.items_ci);
  const int co0 = co_grp * ((MODE == 0 || MODE == 2) ? 128 : 64);
  // second 64-channel block of A; a 64-channel pointwise layer re-reads the first block (its
  // duplicate accumulator rows are discarded)
  const int co1 = (MODE == 2 && co0 + 64 >= p.Cout) ? co0 : co0 + 64;
  const int ci0 = ci_grp * ((MODE == 3 || MODE == 4) ? NBW : NBW * NB);
  const int t_begin = (int)(((long)p.num_pix_tiles * split) / p.ksplit);
  const int t_end = (int)(((long)p.num_pix_tiles * (split + 1)) / p.ksplit);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmDY);
    tma_prefetch_desc(&tmX);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(accum_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base =
      *reinterpret_cast<volatile uint32_t*>(smem_al + kStages * Cfg::kStageBytes + 160);

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = t_begin; t < t_end; ++t) {
        const int twi = t % p.tiles_w;
        const int t2 = t / p.tiles_w;
        const int thi = t2 % p.tiles_h;
        const int img = t2 / p.tiles_h;
        const int w0 = twi * kWgTW;
        const int h0 = thi * kWgTH;
        mbar_wait(empty_bar(stage), phase ^ 1u);
        const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
        const uint32_t sb = sa + Cfg::kASlot;
        mbar_arrive_expect_tx(full_bar(stage), Cfg::kTxBytes);
        if (MODE == 0) {
          tma_load_4d(sa, &tmDY, full_bar(stage), co0, w0, h0, img);
          tma_load_4d(sa + Cfg::kABlock, &tmDY, full_bar(stage), co1, w0, h0, img);
#pragma unroll
          for (int b = 0; b < NB; ++b)
            tma_load_4d(sb + b * Cfg::kBBlock, &tmX, full_bar(stage), ci0 + b * 64, w0 - 1,
                        h0 + r_idx - 1, img);
        } else if (MODE == 2) {
          tma_load_4d(sa, &tmDY, full_bar(stage), co0, w0, h0, img);
          tma_load_4d(sa + Cfg::kABlock, &tmDY, full_bar(stage), co1, w0, h0, img);
#pragma unroll
          for (int b = 0; b < NB; ++b)
            tma_load_4d(sb + b * Cfg::kBBlock, &tmX, full_bar(stage), ci0 + b * 64, w0, h0, img);
        } else if (MODE == 3) {
          tma_load_4d(sa, &tmDY, full_bar(stage), co0, w0, h0, img);
          tma_load_4d(sb, &tmX, full_bar(stage), ci0, w0 - 1, h0 - 1, img);
        } else {
          tma_load_4d(sa, &tmDY, full_bar(stage), co0, w0, h0 - 1, img);
          tma_load_4d(sb, &tmX, full_bar(stage), ci0, w0 - 1, h0, img);
        }
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // warp-uniform loop (descriptors in uniform registers); one elected lane issues tcgen05
    const bool leader = elect_one();
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = t_begin; t < t_end; ++t) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
        const uint32_t sb = sa + Cfg::kASlot;
        constexpr uint32_t kAHi = smem_desc_hi(1024, 128);
        constexpr uint32_t kBHi = smem_desc_hi(kBSBO, kBSwz);
        const uint32_t a_addr16 = (sa >> 4) & 0x3FFF;
        const uint32_t b_addr16 = (sb >> 4) & 0x3FFF;
        const uint32_t acc = (t > t_begin) ? 1u : 0u;
        // MODE_X_SHIFT with two input-channel blocks: both groups multiply the SAME dy slice of K step k, so the loops
        // run K-step-major and the first group parks A in the collector for the second (A is read from shared memory
        // once per K step: 16 instead of 20 KB of operand reads per pair); ACOLL = false keeps the group-major plain form
        constexpr bool kShareA = MODE == 0 && NB == 2 && ACOLL;
#pragma unroll
        for (int o = 0; o < (kShareA ? kWgTH : Cfg::kGroups); ++o) {
#pragma unroll
          for (int i = 0; i < (kShareA ? Cfg::kGroups : kWgTH); ++i) {
            const int g = kShareA ? i : o;     // accumulator group
            const int k = kShareA ? o : i;     // K step k = image row k of the tile (16 pixels)
            uint32_t a_lo, b_lo;
            if (MODE == 0) {
              // A: dy rows k*16.. of both 64-co blocks (LBO = block pitch)
              a_lo = (a_addr16 + ((k * kWgTW * 128) >> 4)) | ((uint32_t(Cfg::kABlock) >> 4) << 16);
              // B: x row k of the haloed filter-row box of channel block g, the taps s = 0..2 as N
              // blocks one pixel apart (dy is read once per K step and channel block, not once per tap)
              b_lo = (b_addr16 + ((g * Cfg::kBBlock + k * kWgBoxW * 128) >> 4)) | ((128u >> 4) << 16);
            } else if (MODE == 2) {
              a_lo = (a_addr16 + ((k * kWgTW * 128) >> 4)) | ((uint32_t(Cfg::kABlock) >> 4) << 16);
              b_lo = (b_addr16 + ((k * kWgTW * 128) >> 4)) | ((uint32_t(Cfg::kBBlock) >> 4) << 16);
            } else if (MODE == 4) {
              // A: dy box row k + 1 - r with r = 1 (group 0) or r = 2 (group 1) in M rows 0..63 and
              // the box row below it (r - 1) in M rows 64..127 (group 1: a discarded duplicate).
              // B: x box row k, taps s = 0..2 as N blocks one pixel apart.
              a_lo = (a_addr16 + ((((k + 1 - g) * kWgTW) * 128) >> 4)) | ((uint32_t(kWgTW * 128) >> 4) << 16);
              b_lo = (b_addr16 + ((k * kWgBoxW * 128) >> 4)) | ((128u >> 4) << 16);
            } else {
              // MODE 3.  A: dy row k of the tile, one 64-channel block.  B: x row k + r of the haloed
              // box (r = g), the three taps s = 0..2 as N blocks one pixel row (kBRow bytes) apart
              a_lo = (a_addr16 + ((k * kWgTW * 128) >> 4)) | ((uint32_t(Cfg::kABlock) >> 4) << 16);
              b_lo = (b_addr16 + ((((k + g) * kWgBoxW) * kBRow) >> 4)) | ((kBRow >> 4) << 16);
            }
            if (leader) {
              if constexpr (kShareA && ACOLL) {
                if (g == 0)
                  umma_bf16_acoll<0>(tmem_base + g * Cfg::kN, smem_desc_join(a_lo, kAHi),
                                     smem_desc_join(b_lo, kBHi), kIdesc, k == 0 ? acc : 1u);
                else
                  umma_bf16_acoll<2>(tmem_base + g * Cfg::kN, smem_desc_join(a_lo, kAHi),
                                     smem_desc_join(b_lo, kBHi), kIdesc, k == 0 ? acc : 1u);
              } else {
                umma_bf16(tmem_base + g * Cfg::kN, smem_desc_join(a_lo, kAHi),
                          smem_desc_join(b_lo, kBHi), kIdesc, k == 0 ? acc : 1u);
              }
            }
          }
        }
        if (leader) umma_commit(empty_bar(stage));
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
      }
      if (leader) umma_commit(accum_bar);
    }
  } else {
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const bool have = t_end > t_begin;
    float* ws = p.ws + (size_t)split * p.Cout * p.taps * p.Cin;
#pragma unroll
    for (int g = 0; g < Cfg::kGroups; ++g) {
      int co, tap;
      if (MODE == 4) {
        co = co0 + (row & 63);
        // group 0: filter rows 1 (M rows 0..63) and 0 (rows 64..127); group 1: filter row 2 and a discard
        tap = g == 0 ? (row < 64 ? 3 : 0) : (row < 64 ? 6 : -1);   // + s per column chunk below
      } else if (MODE == 3) {
        // M = 64 accumulator: rows 16q .. 16q+15 live in lanes 0..15 of TMEM lane quadrant q
        co = co0 + quad * 16 + (lane & 15);
        tap = lane < 16 ? g * 3 : -1;   // + s per column chunk below
      } else if (MODE == 0) {
        co = co0 + row;
        tap = r_idx * 3;          // + s per column chunk below; group g = 64-channel block g
      } else {
        co = co0 + row;
        tap = co < p.Cout ? 0 : -1;
        if (tap < 0) co = 0;
      }
#pragma unroll
      for (int c = 0; c < Cfg::kN / 16; ++c) {
        uint32_t r[16];
        tmem_ld_32x16(tmem_base + (uint32_t(quad * 32) << 16) + g * Cfg::kN + c * 16, r);
        tmem_ld_wait();
        if (tap >= 0) {
          // modes 3 / 4: the accumulator columns are [tap s][NBW channels]; column chunk c belongs to
          // tap s = 16c / NBW, channels 16c % NBW ..; modes 0 / 2: one tap per group, kN channels
          constexpr bool kTapsOnN = MODE == 0 || MODE == 3 || MODE == 4;
          const int tap_c = kTapsOnN ? tap + (c * 16) / NBW : tap;
          const int ch_c = kTapsOnN ? (c * 16) % NBW + (MODE == 0 ? g * NBW : 0) : c * 16;
          float* dst = ws + ((size_t)co * p.taps + tap_c) * p.Cin + ci0 + ch_c;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float4 o;
            o.x = have ? __uint_as_float(r[q * 4 + 0]) : 0.f;
            o.y = have ? __uint_as_float(r[q * 4 + 1]) : 0.f;
            o.z = have ? __uint_as_float(r[q * 4 + 2]) : 0.f;
            o.w = have ? __uint_as_float(r[q * 4 + 3]) : 0.f;
            *reinterpret_cast<float4*>(dst + q * 4) = o;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// dw[co][ci][tap] = sum_split ws[split][co][tap][ci_pad]
__global__ void wgrad_reduce_kernel(const float* __restrict__ ws, float* __restrict__ dw, int ksplit,
                                    int Cout, int Cin, int cin_real, int taps) {
  const long total = (long)Cout * taps * Cin;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total;
       i += (long)gridDim.x * blockDim.x) {
    const int ci = i % Cin;
    const int tap = (i / Cin) % taps;
    const int co = i / ((long)taps * Cin);
    if (ci >= cin_real) continue;
    float acc = 0.f;
    for (int s = 0; s < ksplit; ++s) acc += ws[(size_t)s * total + i];
    dw[((long)co * cin_real + ci) * taps + tap] = acc;
  }
}

struct WgPlan {
  int mode, nbw, nb;
  int n_items, items_ci, items_r, ksplit;
  int tiles_w, tiles_h, num_pix_tiles;
};

static int plan_wgrad(int N, int H, int W, int Cin, int Cout, WgPlan* pl, int taps = 9) {
  if (Cout % 64 != 0 || Cin % 16 != 0) return FPB200_ERR_SHAPE;
  if (taps == 1) {
    if (Cin % 64 != 0) return FPB200_ERR_SHAPE;
    pl->mode = 2; pl->nbw = 64; pl->nb = (Cin % 128 == 0) ? 2 : 1;
    pl->items_r = 1;
    pl->items_ci = Cin / (64 * pl->nb);
    pl->n_items = ((Cout + 127) / 128) * pl->items_ci;
  } else if (Cout % 128 == 0 && Cin % 64 == 0) {
    pl->mode = 0; pl->nbw = 64; pl->nb = (Cin % 128 == 0) ? 2 : 1;
    pl->items_r = 3;
    pl->items_ci = Cin / (64 * pl->nb);
    pl->n_items = (Cout / 128) * pl->items_ci * 3;
  } else if (Cin % 64 == 0) {
    pl->mode = 4; pl->nb = 1; pl->nbw = 64;
    pl->items_r = 1;
    pl->items_ci = Cin / 64;
    pl->n_items = (Cout / 64) * pl->items_ci;
  } else {
    pl->mode = 3; pl->nb = 1;
    pl->nbw = (Cin % 32 == 0) ? 32 : 16;
    pl->items_r = 1;
    pl->items_ci = Cin / pl->nbw;
    pl->n_items = (Cout / 64) * pl->items_ci;
  }
  pl->tiles_w = (W + kWgTW - 1) / kWgTW;
  pl->tiles_h = (H + kWgTH - 1) / kWgTH;
  pl->num_pix_tiles = N * pl->tiles_h * pl->tiles_w;
  // split-K: one CTA per SM (smem-limited), so pick the split count whose CTA total fills
  // whole waves of the 148 SMs best, preferring fewer waves (less partial-sum traffic)
  const int sms = sm_count();
  int best_ks = 1;
  double best_eff = 0.0;
  for (int ks = 1; ks <= 512 && ks <= pl->num_pix_tiles; ++ks) {
    const long ctas = (long)pl->n_items * ks;
    const long waves = (ctas + sms - 1) / sms;
    if (waves > 3) break;
    const double eff = (double)ctas / (double)(waves * sms);
    if (eff > best_eff + 0.02) { best_eff = eff; best_ks = ks; }
  }
  pl->ksplit = best_ks;
  return FPB200_OK;
}

template <int MODE, int NBW, int NB, bool ACOLL = false>
static int launch_wgrad(const CUtensorMap& tmDY, const CUtensorMap& tmX, const WgradParams& p,
                        cudaStream_t stream) {
  using Cfg = WgCfg<MODE, NBW, NB>;
  auto kern = conv3x3_wgrad_kernel<MODE, NBW, NB, ACOLL>;
  static bool attr_set[kMaxDevices] = {false};   // cudaFuncSetAttribute is per device
  const int dev_ = current_device();
  if (!attr_set[dev_]) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes) !=
        cudaSuccess)
      return check_launch("wgrad smem attribute");
    attr_set[dev_] = true;
  }
  kern<<<p.n_items * p.ksplit, kWgThreads, Cfg::kSmemBytes, stream>>>(tmDY, tmX, p);
  return check_launch("conv3x3_wgrad");
}

}  // namespace fp

using namespace fp;

extern "C" {

long fpb200_conv3x3_wgrad_workspace_bytes(int N, int H, int W, int Cin, int Cout) {
  WgPlan pl;
  if (plan_wgrad(N, H, W, Cin, Cout, &pl) != FPB200_OK) return -1;
  return (long)pl.ksplit * Cout * 9 * Cin * (long)sizeof(float);
}

int fpb200_conv3x3_wgrad_bf16_nhwc(const void* x, long ldx, const void* dy, long lddy,
                                   float* dw_oihw, void* workspace, int N, int H, int W, int Cin,
                                   int cin_real, int Cout, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  WgPlan pl;
  int rc = plan_wgrad(N, H, W, Cin, Cout, &pl);
  if (rc != FPB200_OK) return rc;
  if (ldx % 8 != 0 || lddy % 8 != 0 || ldx < Cin || lddy < Cout || cin_real > Cin)
    return FPB200_ERR_SHAPE;
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(dy) & 15) ||
      (reinterpret_cast<uintptr_t>(workspace) & 15))
    return FPB200_ERR_ALIGN;
  CUtensorMap tmDY, tmX;
  if (pl.mode == 0) {
    rc = make_tmap_act(&tmDY, dy, N, H, W, Cout, lddy, 64, kWgTW, kWgTH);
    if (rc != FPB200_OK) return rc;
    rc = make_tmap_act(&tmX, x, N, H, W, Cin, ldx, 64, kWgBoxW, kWgTH);
  } else if (pl.mode == 3) {
    rc = make_tmap_act(&tmDY, dy, N, H, W, Cout, lddy, 64, kWgTW, kWgTH);
    if (rc != FPB200_OK) return rc;
    rc = make_tmap_act(&tmX, x, N, H, W, Cin, ldx, pl.nbw, kWgBoxW, kWgTH + 2);
  } else {
    rc = make_tmap_act(&tmDY, dy, N, H, W, Cout, lddy, 64, kWgTW, kWgTH + 2);
    if (rc != FPB200_OK) return rc;
    rc = make_tmap_act(&tmX, x, N, H, W, Cin, ldx, 64, kWgBoxW, kWgTH);
  }
  if (rc != FPB200_OK) return rc;
  WgradParams p;
  p.N = N; p.H = H; p.W = W;
  p.tiles_w = pl.tiles_w; p.tiles_h = pl.tiles_h;
  p.num_pix_tiles = pl.num_pix_tiles;
  p.ksplit = pl.ksplit; p.n_items = pl.n_items; p.items_ci = pl.items_ci; p.items_r = pl.items_r;
  p.Cout = Cout; p.Cin = Cin; p.taps = 9;
  p.ws = reinterpret_cast<float*>(workspace);
  // FPB200_WGRAD_ACOLL (experiment switch, read once; default on): A collector in the two-block MODE_X_SHIFT kernel.
  // Tensor-bound either way (97 % pipe): +0.5 % over the six layers that use it, 20 % fewer operand reads
  // (alternating A/B on one box: 6.78 / 6.80 vs 6.76 / 6.74 ms, profiles/r02_wgrad_acoll_ab.txt).
  static const bool acoll = [] { const char* e = getenv("FPB200_WGRAD_ACOLL"); return e == nullptr || atoi(e) != 0; }();
  if (pl.mode == 0 && pl.nb == 2 && acoll) rc = launch_wgrad<0, 64, 2, true>(tmDY, tmX, p, stream);
  else if (pl.mode == 0 && pl.nb == 2) rc = launch_wgrad<0, 64, 2>(tmDY, tmX, p, stream);
  else if (pl.mode == 0) rc = launch_wgrad<0, 64, 1>(tmDY, tmX, p, stream);
  else if (pl.mode == 4) rc = launch_wgrad<4, 64, 1>(tmDY, tmX, p, stream);
  else if (pl.mode == 3 && pl.nbw == 32) rc = launch_wgrad<3, 32, 1>(tmDY, tmX, p, stream);
  else rc = launch_wgrad<3, 16, 1>(tmDY, tmX, p, stream);
  if (rc != FPB200_OK) return rc;
  const long total = (long)Cout * 9 * Cin;
  long g = (total + 255) / 256;
  if (g > 148L * 8) g = 148L * 8;
  wgrad_reduce_kernel<<<(int)g, 256, 0, stream>>>(p.ws, dw_oihw, pl.ksplit, Cout, Cin, cin_real, 9);
  return check_launch("wgrad_reduce");
}

long fpb200_conv1x1_wgrad_workspace_bytes(int N, int H, int W, int Cin, int Cout) {
  WgPlan pl;
  if (plan_wgrad(N, H, W, Cin, Cout, &pl, 1) != FPB200_OK) return -1;
  return (long)pl.ksplit * Cout * Cin * (long)sizeof(float);
}

int fpb200_conv1x1_wgrad_bf16_nhwc(const void* x, long ldx, const void* dy, long lddy, float* dw_oi,
                                   void* workspace, int N, int H, int W, int Cin, int Cout,
                                   void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  WgPlan pl;
  int rc = plan_wgrad(N, H, W, Cin, Cout, &pl, 1);
  if (rc != FPB200_OK) return rc;
  if (ldx % 8 != 0 || lddy % 8 != 0 || ldx < Cin || lddy < Cout) return FPB200_ERR_SHAPE;
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(dy) & 15) ||
      (reinterpret_cast<uintptr_t>(workspace) & 15))
    return FPB200_ERR_ALIGN;
  CUtensorMap tmDY, tmX;
  rc = make_tmap_act(&tmDY, dy, N, H, W, Cout, lddy, 64, kWgTW, kWgTH);
  if (rc != FPB200_OK) return rc;
  rc = make_tmap_act(&tmX, x, N, H, W, Cin, ldx, 64, kWgTW, kWgTH);
  if (rc != FPB200_OK) return rc;
  WgradParams p;
  p.N = N; p.H = H; p.W = W;
  p.tiles_w = pl.tiles_w; p.tiles_h = pl.tiles_h;
  p.num_pix_tiles = pl.num_pix_tiles;
  p.ksplit = pl.ksplit; p.n_items = pl.n_items; p.items_ci = pl.items_ci; p.items_r = pl.items_r;
  p.Cout = Cout; p.Cin = Cin; p.taps = 1;
  p.ws = reinterpret_cast<float*>(workspace);
  rc = pl.nb == 2 ? launch_wgrad<2, 64, 2>(tmDY, tmX, p, stream)
                  : launch_wgrad<2, 64, 1>(tmDY, tmX, p, stream);
  if (rc != FPB200_OK) return rc;
  const long total = (long)Cout * Cin;
  long g = (total + 255) / 256;
  if (g > 148L * 8) g = 148L * 8;
  wgrad_reduce_kernel<<<(int)g, 256, 0, stream>>>(p.ws, dw_oi, pl.ksplit, Cout, Cin, Cin, 1);
  return check_launch("wgrad_reduce");
}

}  // extern "C"
