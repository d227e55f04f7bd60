// Weight gradient of the 3x3 / pad 1 convolution on the sm_100a tensor cores.
// (autograd of nn.Conv2d in st_water_seg/models/unet.py:14,16)
//
//   dW[co, (r,s), ci] = sum_{n,h,w} dy[n,h,w,co] * x[n, h+r-1, w+s-1, ci]
//
// is a GEMM whose reduction dimension is the PIXEL axis (up to 16.8 M long) and whose
// output is tiny, so it is split-K: every CTA owns one output tile and one contiguous range
// of 64-pixel tiles, accumulates in TMEM, and writes fp32 partials that a second kernel
// sums deterministically into the OIHW fp32 gradient.
//
// Both GEMM operands are "MN-major": in NHWC memory the channel axis (the GEMM M resp. N
// axis) is the contiguous one and the pixel axis (GEMM K) is strided.  TMA drops a
// [64 pixels][64 channels] box into 128B-swizzled smem (one 128-byte row per pixel) and
// tcgen05.mma consumes it directly with the transposed-operand bits of the instruction
// descriptor set -- no transposes anywhere.  The 3x3 taps are nine shifted views of the same
// tensor; the shift is only a TMA coordinate offset and the conv halo is the TMA
// out-of-bounds zero fill.
//
// Two operand arrangements keep UMMA_M = 128 for every layer of the UNet:
//   MODE_X_SHIFT  (Cout >= 128): A = 128 output channels of dy (unshifted),
//                  B = three column-shifted x boxes of one filter row, N = 64 or 128 each.
//   MODE_DY_SHIFT (Cout == 64 per block): A = TWO differently shifted dy boxes stacked on M
//                  (2 taps x 64 channels), B = unshifted x; 5 such pairs cover the 9 taps.
#include "host_common.h"
#include "ptx.cuh"

namespace fp {

constexpr int kWgThreads = 192;
constexpr int kWgBK = 64;             // pixels per pipeline stage
constexpr int kASlotBytes = kWgBK * 128;  // [64 px][64 ch] bf16

struct WgradParams {
  int N, H, W;
  int tw_log2;           // pixel tile = (64 >> tw_log2) rows x (1 << tw_log2) cols
  int tiles_w, tiles_h;  // per image
  int num_pix_tiles;
  int ksplit;
  int n_items;      // output tiles
  int items_ci;     // number of ci groups (item = co_grp * items_ci * items_r + ci_grp * items_r + r)
  int items_r;      // 3 in MODE_X_SHIFT, 1 in MODE_DY_SHIFT
  int Cout, Cin;    // Cin = padded input channels (layout of the partials)
  float* ws;        // [ksplit][Cout][9][Cin]
};

template <int MODE, int NBW, int NB>
struct WgCfg {
  static constexpr int kNA = MODE == 0 ? 2 : 10;          // A slots (64-channel dy boxes)
  static constexpr int kNBS = MODE == 0 ? 3 * NB : 1;     // B slots
  static constexpr int kBSlotBytes = kWgBK * NBW * 2;
  static constexpr int kStageBytes = kNA * kASlotBytes + kNBS * kBSlotBytes;
  static constexpr int kStagesRaw = (200 * 1024) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 6 ? 6 : kStagesRaw;
  static constexpr int kGroups = MODE == 0 ? 3 : 5;
  static constexpr int kN = NBW * NB;                     // UMMA N per group
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;
};

template <int MODE, int NBW, int NB>
__global__ void __launch_bounds__(kWgThreads, 1)
conv3x3_wgrad_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX,
                     const WgradParams p) {
  using Cfg = WgCfg<MODE, NBW, NB>;
  constexpr int kStages = Cfg::kStages;
  constexpr uint32_t kIdesc = make_idesc_bf16(128, Cfg::kN, 1, 1);
  constexpr uint32_t kBSwz = NBW * 2;
  constexpr uint32_t kBSBO = 8 * NBW * 2;

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
  const int co0 = co_grp * (MODE == 0 ? 128 : 64);
  const int ci0 = ci_grp * Cfg::kN;
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
        const int w0 = twi << p.tw_log2;
        const int h0 = thi * (kWgBK >> p.tw_log2);
        mbar_wait(empty_bar(stage), phase ^ 1u);
        const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
        const uint32_t sb = sa + Cfg::kNA * kASlotBytes;
        mbar_arrive_expect_tx(full_bar(stage), Cfg::kStageBytes);
        if (MODE == 0) {
          tma_load_4d(sa, &tmDY, full_bar(stage), co0, w0, h0, img);
          tma_load_4d(sa + kASlotBytes, &tmDY, full_bar(stage), co0 + 64, w0, h0, img);
#pragma unroll
          for (int s = 0; s < 3; ++s)
#pragma unroll
            for (int b = 0; b < NB; ++b)
              tma_load_4d(sb + (s * NB + b) * Cfg::kBSlotBytes, &tmX, full_bar(stage),
                          ci0 + b * 64, w0 + s - 1, h0 + r_idx - 1, img);
        } else {
#pragma unroll
          for (int a = 0; a < 10; ++a) {
            const int tap = a < 9 ? a : 8;
            const int r = tap / 3, s = tap - 3 * r;
            tma_load_4d(sa + a * kASlotBytes, &tmDY, full_bar(stage), co0, w0 - (s - 1),
                        h0 - (r - 1), img);
          }
          tma_load_4d(sb, &tmX, full_bar(stage), ci0, w0, h0, img);
        }
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = t_begin; t < t_end; ++t) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
        const uint32_t sb = sa + Cfg::kNA * kASlotBytes;
        constexpr uint32_t kAHi = smem_desc_hi(1024, 128);
        constexpr uint32_t kBHi = smem_desc_hi(kBSBO, kBSwz);
        const uint32_t a_lo = smem_desc_lo(sa, kASlotBytes);
        const uint32_t b_lo = smem_desc_lo(sb, Cfg::kBSlotBytes);
        const uint32_t acc = (t > t_begin) ? 1u : 0u;
#pragma unroll
        for (int g = 0; g < Cfg::kGroups; ++g) {
          // compile-time operand offsets (bytes): slot of the group, then 16 pixels per K step
          constexpr uint32_t kAStep = 2048, kBStep = 2 * kBSBO;
          const uint32_t a_g = MODE == 0 ? 0u : uint32_t(2 * g) * kASlotBytes;
          const uint32_t b_g = MODE == 0 ? uint32_t(g * NB) * Cfg::kBSlotBytes : 0u;
#pragma unroll
          for (int k = 0; k < kWgBK / 16; ++k) {
            umma_bf16(tmem_base + g * Cfg::kN, smem_desc_join(a_lo + ((a_g + k * kAStep) >> 4), kAHi),
                      smem_desc_join(b_lo + ((b_g + k * kBStep) >> 4), kBHi), kIdesc,
                      k == 0 ? acc : 1u);
          }
        }
        umma_commit(empty_bar(stage));
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
      }
      umma_commit(accum_bar);
    }
  } else {
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const bool have = t_end > t_begin;
    float* ws = p.ws + (size_t)split * p.Cout * 9 * p.Cin;
#pragma unroll
    for (int g = 0; g < Cfg::kGroups; ++g) {
      int co, tap;
      if (MODE == 0) {
        co = co0 + row;
        tap = r_idx * 3 + g;
      } else {
        co = co0 + (row & 63);
        tap = 2 * g + (row >> 6);
      }
      float* dst = ws + ((size_t)co * 9 + tap) * p.Cin + ci0;
#pragma unroll
      for (int c = 0; c < Cfg::kN / 16; ++c) {
        uint32_t r[16];
        tmem_ld_32x16(tmem_base + (uint32_t(quad * 32) << 16) + g * Cfg::kN + c * 16, r);
        tmem_ld_wait();
        if (tap < 9) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float4 o;
            o.x = have ? __uint_as_float(r[q * 4 + 0]) : 0.f;
            o.y = have ? __uint_as_float(r[q * 4 + 1]) : 0.f;
            o.z = have ? __uint_as_float(r[q * 4 + 2]) : 0.f;
            o.w = have ? __uint_as_float(r[q * 4 + 3]) : 0.f;
            *reinterpret_cast<float4*>(dst + c * 16 + q * 4) = o;
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
                                    int Cout, int Cin, int cin_real) {
  const long total = (long)Cout * 9 * Cin;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total;
       i += (long)gridDim.x * blockDim.x) {
    const int ci = i % Cin;
    const int tap = (i / Cin) % 9;
    const int co = i / (9L * Cin);
    if (ci >= cin_real) continue;
    float acc = 0.f;
    for (int s = 0; s < ksplit; ++s) acc += ws[(size_t)s * total + i];
    dw[((long)co * cin_real + ci) * 9 + tap] = acc;
  }
}

struct WgPlan {
  int mode, nbw, nb;
  int n_items, items_ci, items_r, ksplit;
  int tw_log2, tiles_w, tiles_h, num_pix_tiles;
};

static int plan_wgrad(int N, int H, int W, int Cin, int Cout, WgPlan* pl) {
  if (Cout % 64 != 0 || Cin % 16 != 0) return FPB200_ERR_SHAPE;
  if (Cout % 128 == 0 && Cin % 64 == 0) {
    pl->mode = 0; pl->nbw = 64; pl->nb = (Cin % 128 == 0) ? 2 : 1;
    pl->items_r = 3;
    pl->items_ci = Cin / (64 * pl->nb);
    pl->n_items = (Cout / 128) * pl->items_ci * 3;
  } else {
    pl->mode = 1; pl->nb = 1;
    pl->nbw = (Cin % 64 == 0) ? 64 : ((Cin % 32 == 0) ? 32 : 16);
    pl->items_r = 1;
    pl->items_ci = Cin / pl->nbw;
    pl->n_items = (Cout / 64) * pl->items_ci;
  }
  // pixel tiles of 64: pick the shape with least padding (ties -> wider)
  int best_l = 3;
  long best_area = -1;
  for (int l = 3; l <= 6; ++l) {
    const int tw = 1 << l, th = kWgBK >> l;
    const long area = (long)((W + tw - 1) / tw) * tw * (long)((H + th - 1) / th) * th;
    if (best_area < 0 || area <= best_area) { best_area = area; best_l = l; }
  }
  pl->tw_log2 = best_l;
  const int TW = 1 << best_l, TH = kWgBK >> best_l;
  pl->tiles_w = (W + TW - 1) / TW;
  pl->tiles_h = (H + TH - 1) / TH;
  pl->num_pix_tiles = N * pl->tiles_h * pl->tiles_w;
  // split-K so that ~2 waves of CTAs exist, but never more splits than pixel tiles
  int ks = (2 * sm_count() + pl->n_items - 1) / pl->n_items;
  if (ks < 1) ks = 1;
  if (ks > pl->num_pix_tiles) ks = pl->num_pix_tiles;
  if (ks > 512) ks = 512;
  pl->ksplit = ks;
  return FPB200_OK;
}

template <int MODE, int NBW, int NB>
static int launch_wgrad(const CUtensorMap& tmDY, const CUtensorMap& tmX, const WgradParams& p,
                        cudaStream_t stream) {
  using Cfg = WgCfg<MODE, NBW, NB>;
  auto kern = conv3x3_wgrad_kernel<MODE, NBW, NB>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes) !=
        cudaSuccess)
      return check_launch("wgrad smem attribute");
    attr_set = true;
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
  const int TW = 1 << pl.tw_log2, TH = kWgBK >> pl.tw_log2;
  CUtensorMap tmDY, tmX;
  rc = make_tmap_act(&tmDY, dy, N, H, W, Cout, lddy, 64, TW, TH);
  if (rc != FPB200_OK) return rc;
  rc = make_tmap_act(&tmX, x, N, H, W, Cin, ldx, pl.nbw, TW, TH);
  if (rc != FPB200_OK) return rc;
  WgradParams p;
  p.N = N; p.H = H; p.W = W;
  p.tw_log2 = pl.tw_log2; p.tiles_w = pl.tiles_w; p.tiles_h = pl.tiles_h;
  p.num_pix_tiles = pl.num_pix_tiles;
  p.ksplit = pl.ksplit; p.n_items = pl.n_items; p.items_ci = pl.items_ci; p.items_r = pl.items_r;
  p.Cout = Cout; p.Cin = Cin;
  p.ws = reinterpret_cast<float*>(workspace);
  if (pl.mode == 0 && pl.nb == 2) rc = launch_wgrad<0, 64, 2>(tmDY, tmX, p, stream);
  else if (pl.mode == 0) rc = launch_wgrad<0, 64, 1>(tmDY, tmX, p, stream);
  else if (pl.nbw == 64) rc = launch_wgrad<1, 64, 1>(tmDY, tmX, p, stream);
  else if (pl.nbw == 32) rc = launch_wgrad<1, 32, 1>(tmDY, tmX, p, stream);
  else rc = launch_wgrad<1, 16, 1>(tmDY, tmX, p, stream);
  if (rc != FPB200_OK) return rc;
  const long total = (long)Cout * 9 * Cin;
  long g = (total + 255) / 256;
  if (g > 148L * 8) g = 148L * 8;
  wgrad_reduce_kernel<<<(int)g, 256, 0, stream>>>(p.ws, dw_oihw, pl.ksplit, Cout, Cin, cin_real);
  return check_launch("wgrad_reduce");
}

}  // extern "C"
