// 3x3 / pad 1 / stride 1 convolution as an implicit GEMM on the sm_100a tensor cores,
// "halo" formulation.
//
// Replaces what nn.Conv2d(k=3, padding=1) dispatches to in the reference
// (st_water_seg/models/unet.py:14,16) -- forward (fprop) and, with the transposed +
// rotated weight packing, the data gradient (dgrad, the autograd of the same line).
//
//   D[m, co] = sum_{tap, ci} A[m + tap, ci] * Wp[co, tap, ci]
//
// Work item of a CTA: a 16x16 pixel patch of one image (= two 128-row MMA tiles, the left
// and right 16x8 halves) x BN output channels.  For every 64-channel chunk of the input the
// TMA unit fetches ONE (16+2) x (16+2) x KCH box -- the patch plus its 1-pixel halo, zero
// filled outside the image -- into swizzled shared memory, one 128-byte row per pixel.  All
// nine filter taps and both MMA tiles then read that single box: a tap shift (r, s) is just
// a different start address (r*18 + s rows further) in the tcgen05 shared-memory descriptor,
// and the 8-row core-matrix groups of a tile are one image row each, a constant 18 rows
// apart (descriptor stride).  (Measured on B200: the 128B/64B/32B swizzle is a function of
// the absolute shared-memory address, so operand start addresses that are only 128-byte --
// not 1024-byte -- aligned read back exactly what TMA wrote, with descriptor base_offset 0.)  Activation traffic into the SM drops from 9 boxes per chunk to
// 1.27 (halo overhead) and every weight tile [BN x KCH] is used by two MMA tiles, which is
// what lifts the kernel from smem-fill bound to tensor-pipe bound.
//
// Items are ordered patch-major / channel-block-minor: with 148 CTAs and 1, 2 or 4 channel blocks
// every CTA keeps one channel block for the whole launch (its BatchNorm partials stay in
// registers) while neighbouring CTAs work on the same patch, so the activation box is fetched
// from HBM once and re-served from L2 for the other channel blocks.
//
// Pipeline: persistent CTAs (one per SM), warp specialised
//   warp 0    TMA producer: ring A (activation boxes), ring B (weight tiles per tap)
//   warp 1    tcgen05.mma issuer + TMEM allocator
//   warps 2.. epilogue (4, or 8 for BN = 64): tcgen05.ld -> registers -> fused epilogue -> global
// TMEM holds 2 (double buffer) x 2 (tiles) x BN fp32 accumulator columns, so the epilogue
// of one patch overlaps the MMAs of the next.
//
// Fused epilogues (runtime flags, warp-uniform):
//   * per-channel affine (+ReLU): eval-mode BatchNorm folded to scale/shift
//   * bf16 cast, staged per warp in swizzled smem and written with TMA stores (8 px x 4 rows
//     x 64 ch boxes; ragged image edges are clipped by the TMA unit) into an NHWC view with
//     arbitrary pixel pitch
//   * BatchNorm batch-statistic partials: per-channel sum / sum of squares of the stored bf16
//     values, read back conflict-free from the staging tile, accumulated in registers across
//     all patches of the persistent CTA.
#include <stdlib.h>
#include "host_common.h"
#include "ptx.cuh"

namespace fp {


struct HaloParams {
  int N, H, W;
  int Cin;   // padded input channels (multiple of KCH)
  int Cout;  // multiple of BN
  int patches_w, patches_h;
  int num_patches, num_n_blks;
  __nv_bfloat16* y;
  long ldy;
  const float* scale;    // stats_mode 0/1: optional epilogue affine; stats_mode 2: BN scale of the
  const float* shift;    //   layer whose gradient this dgrad produces (with mean / invstd below)
  const float* bn_mean;
  const float* bn_invstd;
  int relu;
  int stats_mode;        // 0 none, 1 BatchNorm forward sums (y, y^2), 2 BatchNorm backward sums (g, g*xhat)
  float* stat_partials;  // [fpb200_conv_stat_rows()][2][Cout], row = CTA * epilogue warps + warp
};

constexpr int kPatch = 16;        // patch edge in pixels

// buf / op are compile-time constants after the issue loop is unrolled: the switch folds to one instruction
__device__ __forceinline__ void umma_bf16_ws_sel(int buf, int op, uint32_t d, uint64_t a, uint64_t b, uint32_t idesc,
                                                 uint32_t acc) {
  switch (buf * 3 + op) {
    case 0: umma_bf16_ws<0, 0>(d, a, b, idesc, acc); break;
    case 2: umma_bf16_ws<0, 2>(d, a, b, idesc, acc); break;
    case 3: umma_bf16_ws<1, 0>(d, a, b, idesc, acc); break;
    case 5: umma_bf16_ws<1, 2>(d, a, b, idesc, acc); break;
    case 6: umma_bf16_ws<2, 0>(d, a, b, idesc, acc); break;
    case 8: umma_bf16_ws<2, 2>(d, a, b, idesc, acc); break;
    case 9: umma_bf16_ws<3, 0>(d, a, b, idesc, acc); break;
    default: umma_bf16_ws<3, 2>(d, a, b, idesc, acc); break;
  }
}

// TAPS = 9: 3x3 / pad 1 (box = patch + 1-pixel halo); TAPS = 1: pointwise (1x1) convolution,
// the same pipeline with a halo-free 16 x 16 box and a single "tap".
template <int BN, int KCH, int TAPS, int EW = (BN >= 128 ? 4 : 8)>
struct HaloCfg {
  static_assert(TAPS == 9 || TAPS == 1, "3x3 or 1x1");
  static_assert(EW == 4 || EW == 8, "4 or 8 epilogue warps");
  static constexpr int kHalo = TAPS == 9 ? 1 : 0;
  static constexpr int kBox = kPatch + 2 * kHalo;                   // box edge in pixels
  static constexpr int kBoxRows = kBox * kBox;
  static constexpr int kRowBytes = KCH * 2;
  static constexpr int kABytes = kBoxRows * kRowBytes;             // bytes the TMA writes
  static constexpr int kASlot = (kABytes + 1023) / 1024 * 1024;    // ring pitch
  static constexpr int kBBytes = BN * kRowBytes;
  // epilogue warps: one per (TMEM lane quadrant, MMA tile) when the epilogue is co-critical
  // (BN = 64: the MMAs are 32 tensor cycles each; BN = 128 with a single 64-channel K chunk: only
  // 36 MMAs per tile pair), one per quadrant otherwise
  static constexpr int kEpiWarps = EW;
  static constexpr int kNB = BN >= 128 ? (EW == 8 ? 4 : 5) : 8;    // weight ring depth (<= 16)
  static constexpr int kThreads = 64 + 32 * kEpiWarps;
  static constexpr int kStageOut = kEpiWarps * 4096;               // epilogue output staging, one 4 KB tile per warp
  // staging for the fused BatchNorm-backward reduction (dgrad): the y tile of the layer being
  // differentiated, TMA-loaded per (tile, unit); ping-pong when a warp walks several units
  static constexpr int kYBufs = (BN >= 128 && EW == 4) ? 2 : 1;
  static constexpr int kStageY = kEpiWarps * kYBufs * 4096;
  static constexpr int kBudget = 212 * 1024;
  static constexpr int kNARaw = (kBudget - kStageOut - kStageY - kNB * kBBytes) / kASlot;
  static constexpr int kNA = kNARaw > 4 ? 4 : kNARaw;
  static constexpr int kTmemCols = 4 * BN <= 256 ? 256 : 512;
  static constexpr int kRingBytes = kNA * kASlot + ((kNB * kBBytes + 1023) / 1024 * 1024);
  static constexpr int kDataBytes = kRingBytes + kStageOut + kStageY;
  static constexpr int kSmemBytes = kDataBytes + 1024 + 1024 + 4 * 512 * 4;
  static_assert(kNA >= 2, "need at least two activation slots");
  static_assert(4 * BN <= 512, "TMEM: 2 tiles x 2 stages x BN columns");
};

// WS: weight-stationary MMA issue (see the MMA issuer below); a template parameter because the issue loop is one
// lane's dependent instruction chain -- a runtime switch inside it cost every kernel 25 % (measured).
template <int BN, int KCH, int TAPS, int EW = (BN >= 128 ? 4 : 8), bool WS = false>
__global__ void __launch_bounds__(HaloCfg<BN, KCH, TAPS, EW>::kThreads, 1)
conv3x3_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmYL,
                    const HaloParams p) {
  using Cfg = HaloCfg<BN, KCH, TAPS, EW>;
  constexpr int kNA = Cfg::kNA, kNB = Cfg::kNB;
  constexpr int kBox = Cfg::kBox, kHalo = Cfg::kHalo;
  constexpr uint32_t kRB = Cfg::kRowBytes;
  constexpr uint32_t kSwz = kRB;
  constexpr uint32_t kIdesc = make_idesc_bf16(128, BN, 0, 0);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t a_base = smem_base;
  const uint32_t b_base = smem_base + kNA * Cfg::kASlot;
  const uint32_t stg_base = smem_base + Cfg::kRingBytes;
  const uint32_t ystg_base = stg_base + Cfg::kStageOut;
  const uint32_t bar_base = smem_base + Cfg::kDataBytes;
  auto afull = [&](int s) { return bar_base + 8u * s; };            // 4
  auto aempty = [&](int s) { return bar_base + 32u + 8u * s; };     // 4
  auto bfull = [&](int s) { return bar_base + 64u + 8u * s; };      // 16
  auto bempty = [&](int s) { return bar_base + 192u + 8u * s; };    // 16
  auto tfull = [&](int s) { return bar_base + 320u + 8u * s; };     // 2
  auto tempty = [&](int s) { return bar_base + 336u + 8u * s; };    // 2
  const uint32_t tmem_slot = bar_base + 352u;
  auto ybar = [&](int w, int b) { return bar_base + 384u + 8u * (w * 2 + b); };   // 16
  float* s_scale = reinterpret_cast<float*>(smem_al + Cfg::kDataBytes + 1024);
  float* s_shift = s_scale + 512;
  float* s_mean = s_scale + 1024;
  float* s_invstd = s_scale + 1536;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_items = p.num_patches * p.num_n_blks;
  const int k_chunks = p.Cin / KCH;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmY);
    if (p.stats_mode == 2) tma_prefetch_desc(&tmYL);
    for (int s = 0; s < kNA; ++s) { mbar_init(afull(s), 1); mbar_init(aempty(s), 1); }
    for (int s = 0; s < kNB; ++s) { mbar_init(bfull(s), 1); mbar_init(bempty(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull(s), 1); mbar_init(tempty(s), Cfg::kEpiWarps); }
    for (int s = 0; s < 16; ++s) mbar_init(bar_base + 384u + 8u * s, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  if (p.scale != nullptr) {
    for (int c = threadIdx.x; c < p.Cout; c += Cfg::kThreads) {
      s_scale[c] = p.scale[c];
      s_shift[c] = p.shift[c];
      if (p.stats_mode == 2) {
        s_mean[c] = p.bn_mean[c];
        s_invstd[c] = p.bn_invstd[c];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base =
      *reinterpret_cast<volatile uint32_t*>(smem_al + Cfg::kDataBytes + 352);

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const int patch = item / p.num_n_blks;   // n_blk-minor: CTAs b, b+1 share a patch (L2 hit)
        const int n_blk = item - patch * p.num_n_blks;
        const int pw = patch % p.patches_w;
        const int t2 = patch / p.patches_w;
        const int ph = t2 % p.patches_h;
        const int img = t2 / p.patches_h;
        const int w0 = pw * kPatch, h0 = ph * kPatch;
        for (int kc = 0; kc < k_chunks; ++kc) {
          mbar_wait(aempty(sa), pa ^ 1u);
          mbar_arrive_expect_tx(afull(sa), Cfg::kABytes);
          tma_load_4d(a_base + sa * Cfg::kASlot, &tmA, afull(sa), kc * KCH, w0 - kHalo, h0 - kHalo, img);
          if (++sa == kNA) { sa = 0; pa ^= 1u; }
          for (int tap = 0; tap < TAPS; ++tap) {
            mbar_wait(bempty(sb), pb ^ 1u);
            mbar_arrive_expect_tx(bfull(sb), Cfg::kBBytes);
            tma_load_2d(b_base + sb * Cfg::kBBytes, &tmB, bfull(sb), tap * p.Cin + kc * KCH,
                        n_blk * BN);
            if (++sb == kNB) { sb = 0; pb ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp walks the loop with warp-uniform values (so addresses/descriptors live in
    // uniform registers); only the tcgen05 instructions are issued by one elected lane.
    const bool leader = elect_one();
    // Weight-stationary issue: both 128-row tiles of the patch multiply the same weight slice, so tile 0 issues
    // tcgen05.mma.ws with collector::b<k>::fill (K slice k of the tap parks in collector buffer k) and tile 1 with
    // ::lastuse -- the 2 KB (BN = 64) slice is read from shared memory once per tap instead of twice.  An
    // M128 x N64 x K16 MMA needs 6 KB of operands per 32 tensor cycles against 128 B/clk of shared-memory operand
    // bandwidth: 48.0 cycles per MMA measured for the plain form (50.2 with the tap-shifted descriptors), 43.5 with
    // the collector (scripts/umma_ws_microbench.cu, profiles/r03_umma_ws_microbench.txt).
    static_assert(!WS || KCH / 16 <= 4, "one collector buffer per K slice");
    {
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      int it = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(tempty(as), aphase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * 2 * BN;
        for (int kc = 0; kc < k_chunks; ++kc) {
          mbar_wait(afull(sa), pa);
          tc_fence_after();
          const uint32_t a_lo = smem_desc_lo(a_base + sa * Cfg::kASlot, 16);
#pragma unroll
          for (int tap = 0; tap < TAPS; ++tap) {
            constexpr uint32_t kAHi = smem_desc_hi(kBox * kRB, kSwz);
            constexpr uint32_t kBHi = smem_desc_hi(8 * kRB, kSwz);
            const int r = tap / 3, s = tap - 3 * r;
            mbar_wait(bfull(sb), pb);
            tc_fence_after();
            const uint32_t b_lo = smem_desc_lo(b_base + sb * Cfg::kBBytes, 16);
#pragma unroll
            for (int t = 0; t < 2; ++t) {
#pragma unroll
              for (int k = 0; k < KCH / 16; ++k) {
                // tap shift (r, s), tile half t and K slice k are compile-time byte offsets
                const uint32_t a_off = (uint32_t(r * kBox + t * 8 + s) * kRB + k * 32) >> 4;
                const uint32_t acc = (tap == 0 && k == 0) ? (kc != 0 ? 1u : 0u) : 1u;
                if (leader) {
                  if constexpr (WS)
                    umma_bf16_ws_sel(k, t == 0 ? 0 : 2, d_tmem + t * BN, smem_desc_join(a_lo + a_off, kAHi),
                                     smem_desc_join(b_lo + ((k * 32) >> 4), kBHi), kIdesc, acc);
                  else
                    umma_bf16(d_tmem + t * BN, smem_desc_join(a_lo + a_off, kAHi),
                              smem_desc_join(b_lo + ((k * 32) >> 4), kBHi), kIdesc, acc);
                }
              }
            }
            if (leader) umma_commit(bempty(sb));
            if (++sb == kNB) { sb = 0; pb ^= 1u; }
          }
          if (leader) umma_commit(aempty(sa));
          if (++sa == kNA) { sa = 0; pa ^= 1u; }
        }
        if (leader) umma_commit(tfull(as));
      }
    }
  } else {
    // ===================== epilogue warps (2..5) =====================
    // Per (MMA tile, 64-channel unit): TMEM -> registers -> (affine/ReLU) -> bf16 -> this warp's
    // 32-row x 128-byte staging tile in smem (128B-swizzled) -> one TMA store of the
    // 8 px x 4 rows x 64 ch box.  BatchNorm statistics are column sums over the staged tile
    // (i.e. of exactly the bf16 values that are stored and later normalised).
    const int quad = warp & 3;
    const int row = quad * 32 + lane;  // MMA tile row = (th, tw) = (row >> 3, row & 7)
    const int ew = warp - 2;
    constexpr int kTilesPerWarp = Cfg::kEpiWarps == 8 ? 1 : 2;
    const int t_first = Cfg::kEpiWarps == 8 ? (ew >> 2) : 0;   // 8 warps: warps 2-5 tile 0, 6-9 tile 1
    const int stats_mode = p.stat_partials != nullptr ? p.stats_mode : 0;
    const bool do_stats = stats_mode != 0;
    const bool do_affine = p.scale != nullptr && stats_mode != 2;
    const uint32_t stg = stg_base + ew * 4096;                      // this warp's output staging tile
    const uint32_t ystg0 = ystg_base + ew * (Cfg::kYBufs * 4096);   // ... and its y tile(s)
    uint32_t ybuf_issue = 0, ybuf_use = 0, yphase = 0;              // y-tile ping-pong state (bit b = parity of buffer b)
    constexpr int kUnits = BN / 64;
    constexpr int kUnitsPerItem = kTilesPerWarp * kUnits;
    float acc_sum[kUnits][2], acc_sq[kUnits][2];
#pragma unroll
    for (int u = 0; u < kUnits; ++u) { acc_sum[u][0] = acc_sum[u][1] = acc_sq[u][0] = acc_sq[u][1] = 0.f; }
    int cur_n_blk = -1;
    auto flush_stats = [&]() {
      if (do_stats && cur_n_blk >= 0) {
        float* dst = p.stat_partials + (size_t)(blockIdx.x * Cfg::kEpiWarps + ew) * 2 * p.Cout + cur_n_blk * BN;
#pragma unroll
        for (int u = 0; u < kUnits; ++u) {
          // this (row, column) slot is owned by exactly this thread and was zeroed by the host,
          // so a CTA that comes back to an n_blk accumulates without atomics
          float2* d0 = reinterpret_cast<float2*>(dst + u * 64 + 2 * lane);
          float2* d1 = reinterpret_cast<float2*>(dst + p.Cout + u * 64 + 2 * lane);
          const float2 o0 = *d0, o1 = *d1;
          *d0 = make_float2(o0.x + acc_sum[u][0], o0.y + acc_sum[u][1]);
          *d1 = make_float2(o1.x + acc_sq[u][0], o1.y + acc_sq[u][1]);
          acc_sum[u][0] = acc_sum[u][1] = acc_sq[u][0] = acc_sq[u][1] = 0.f;
        }
      }
    };
    int it = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
      const int patch = item / p.num_n_blks;
      const int n_blk = item - patch * p.num_n_blks;
      const int pw = patch % p.patches_w;
      const int t2 = patch / p.patches_w;
      const int ph = t2 % p.patches_h;
      const int img = t2 / p.patches_h;
      if (n_blk != cur_n_blk) { flush_stats(); cur_n_blk = n_blk; }
      // fused BatchNorm-backward: fetch the y tile of unit `k` of this item (same box as the
      // store) into the next ping-pong buffer; issued one unit ahead so the latency hides behind
      // the MMA wait / the previous unit's work
      auto issue_y = [&](int k) {
        const int t = t_first + k / kUnits, u = k % kUnits;
        if (lane == 0) {
          mbar_arrive_expect_tx(ybar(ew, ybuf_issue), 4096);
          tma_load_4d(ystg0 + ybuf_issue * 4096, &tmYL, ybar(ew, ybuf_issue), n_blk * BN + u * 64,
                      pw * kPatch + t * 8, ph * kPatch + quad * 4, img);
        }
        ybuf_issue = (ybuf_issue + 1) % Cfg::kYBufs;
      };
      if (stats_mode == 2) issue_y(0);   // all lanes finished reading this buffer: loop-top __syncwarp of the last unit
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(tfull(as), aphase);
      tc_fence_after();
      const int py = ph * kPatch + (row >> 3);
#pragma unroll
      for (int tt = 0; tt < kTilesPerWarp; ++tt) {
        const int t = t_first + tt;
        const int px = pw * kPatch + t * 8 + (row & 7);
        const bool valid = (px < p.W) && (py < p.H);
        const uint32_t vmask = __ballot_sync(0xffffffffu, valid);
#pragma unroll
        for (int u = 0; u < kUnits; ++u) {
          // the previous TMA store must have finished reading the staging tile (a second,
          // ping-pong tile was measured to buy nothing: the store drains long before the next
          // unit's accumulator has been converted)
          const uint32_t stg_row = stg + lane * 128;
          if (lane == 0) tma_store_wait_read();
          __syncwarp();
          if (stats_mode == 2 && Cfg::kYBufs == 2) {
            const int k = tt * kUnits + u;
            if (k + 1 < kUnitsPerItem) issue_y(k + 1);   // the other buffer was released by the __syncwarp above
          }
#pragma unroll
          for (int hlf = 0; hlf < 2; ++hlf) {
            uint32_t r[32];
            tmem_ld_32x32(tmem_base + (uint32_t(quad * 32) << 16) + as * 2 * BN + t * BN + u * 64 +
                              hlf * 32, r);
            tmem_ld_wait();
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
            if (do_affine) {
              const int cb = n_blk * BN + u * 64 + hlf * 32;
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                v[j] = fmaf(v[j], s_scale[cb + j], s_shift[cb + j]);
                if (p.relu) v[j] = fmaxf(v[j], 0.f);
              }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint4 o;
              o.x = pack_bf16x2(v[q * 8 + 0], v[q * 8 + 1]);
              o.y = pack_bf16x2(v[q * 8 + 2], v[q * 8 + 3]);
              o.z = pack_bf16x2(v[q * 8 + 4], v[q * 8 + 5]);
              o.w = pack_bf16x2(v[q * 8 + 6], v[q * 8 + 7]);
              st_shared_v4(stg_row + (uint32_t((hlf * 4 + q) ^ (lane & 7)) << 4), o);
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_4d(&tmY, stg, n_blk * BN + u * 64, pw * kPatch + t * 8,
                         ph * kPatch + quad * 4, img);
            tma_store_commit();
          }
          if (stats_mode == 2) {
            // BatchNorm-backward sums of the layer whose activation gradient this dgrad writes:
            //   g = dx * [y*scale+shift > 0],  sum g  and  sum g*xhat,  xhat = (y-mean)*invstd
            // dx = the staged bf16 tile, y = the TMA-loaded tile of the same box.
            const uint32_t ybuf = ystg0 + ybuf_use * 4096;
            mbar_wait(ybar(ew, ybuf_use), (yphase >> ybuf_use) & 1u);
            yphase ^= 1u << ybuf_use;
            ybuf_use = (ybuf_use + 1) % Cfg::kYBufs;
            const int cb = n_blk * BN + u * 64 + 2 * lane;
            const float sc0 = s_scale[cb], sc1 = s_scale[cb + 1], sh0 = s_shift[cb], sh1 = s_shift[cb + 1];
            const float mu0 = s_mean[cb], mu1 = s_mean[cb + 1], is0 = s_invstd[cb], is1 = s_invstd[cb + 1];
            float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
            // rows are read in batches of 8 independent loads (a branch per row serialises the
            // shared-memory latency: measured ~35 cycles per row on the epilogue's critical path);
            // rows outside the image contribute through a 0/1 factor instead of a branch
#pragma unroll
            for (int r0 = 0; r0 < 32; r0 += 8) {
              uint32_t gv[8], yv[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const int rr = r0 + j;
                const uint32_t off = rr * 128 + (uint32_t((lane >> 2) ^ (rr & 7)) << 4) + (lane & 3) * 4;
                gv[j] = ld_shared_u32(stg + off);
                yv[j] = ld_shared_u32(ybuf + off);
              }
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const bool ok = (vmask >> (r0 + j)) & 1u;
                const float y0 = bf16_lo(yv[j]), y1 = bf16_hi(yv[j]);
                const float g0 = (ok && fmaf(y0, sc0, sh0) > 0.f) ? bf16_lo(gv[j]) : 0.f;
                const float g1 = (ok && fmaf(y1, sc1, sh1) > 0.f) ? bf16_hi(gv[j]) : 0.f;
                s0 += g0; s1 += g1;
                q0 = fmaf(g0, (y0 - mu0) * is0, q0);
                q1 = fmaf(g1, (y1 - mu1) * is1, q1);
              }
            }
            acc_sum[u][0] += s0; acc_sum[u][1] += s1;
            acc_sq[u][0] += q0; acc_sq[u][1] += q1;
            if (Cfg::kYBufs == 1 && kUnitsPerItem > 1) {
              // single y buffer, several units per item: fetch the next unit's tile once every lane
              // has finished reading this one
              const int k = tt * kUnits + u;
              __syncwarp();
              if (k + 1 < kUnitsPerItem) issue_y(k + 1);
            }
          } else if (do_stats) {
            // lane l owns channels 2l, 2l+1 of this unit: walk the 32 staged rows.  The channel pair is summed with
            // packed fp32 instructions (FADD2 / FFMA2: same bits as two scalar operations, half the issue slots -- the
            // epilogue of the 64-channel layers is instruction bound); tiles that lie entirely inside the image, i.e.
            // nearly all of them, skip the per-row validity select.
            float2 sp = make_float2(0.f, 0.f), qp = make_float2(0.f, 0.f);
            const bool all_valid = vmask == 0xffffffffu;
#pragma unroll
            for (int r0 = 0; r0 < 32; r0 += 8) {
              uint32_t wv[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const int rr = r0 + j;
                wv[j] = ld_shared_u32(stg + rr * 128 + (uint32_t((lane >> 2) ^ (rr & 7)) << 4) + (lane & 3) * 4);
              }
              if (all_valid) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float2 v = make_float2(bf16_lo(wv[j]), bf16_hi(wv[j]));
                  sp = add2(sp, v);
                  qp = fma2(v, v, qp);
                }
              } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const bool ok = (vmask >> (r0 + j)) & 1u;
                  const float2 v = make_float2(ok ? bf16_lo(wv[j]) : 0.f, ok ? bf16_hi(wv[j]) : 0.f);
                  sp = add2(sp, v);
                  qp = fma2(v, v, qp);
                }
              }
            }
            acc_sum[u][0] += sp.x; acc_sum[u][1] += sp.y;
            acc_sq[u][0] += qp.x; acc_sq[u][1] += qp.y;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty(as));
    }
    flush_stats();
    if (lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// FPB200_HALO_WS (experiment switch, read once): bit 0 = weight-stationary MMAs in the 64 -> 64 / 128 -> 64 channel
// kernel, bit 1 = in the 128-output-channel kernels with 64-channel chunks; both on by default.  At N = 128 the MMA
// stream itself is no faster (64 cycles either way) but operand reads drop from 128 to 96 B/clk, which leaves the
// epilogue's staging / statistics traffic room beside it: +1.5 % fprop, +0.7 % dgrad over the Cout >= 128 layers,
// +0.6 % on the step (alternating A/B, profiles/r02_ws_halo64_ncu.md).
static int halo_ws_mask() {
  static const int mask = [] {
    const char* e = getenv("FPB200_HALO_WS");
    return e != nullptr ? atoi(e) : 3;
  }();
  return mask;
}

template <int BN, int KCH, int TAPS, int EW = (BN >= 128 ? 4 : 8), bool WS = false>
static int launch_halo(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmY,
                       const CUtensorMap& tmYL, const HaloParams& p, cudaStream_t stream) {
  using Cfg = HaloCfg<BN, KCH, TAPS, EW>;
  auto kern = conv3x3_halo_kernel<BN, KCH, TAPS, EW, WS>;
  static bool attr_set[kMaxDevices] = {false};   // cudaFuncSetAttribute is per device
  const int dev_ = current_device();
  if (!attr_set[dev_]) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes) !=
        cudaSuccess)
      return check_launch("conv3x3_halo smem attribute");
    attr_set[dev_] = true;
  }
  const int items = p.num_patches * p.num_n_blks;
  int grid = sm_count();
  if (grid > items) grid = items;
  kern<<<grid, Cfg::kThreads, Cfg::kSmemBytes, stream>>>(tmA, tmB, tmY, tmYL, p);
  return check_launch("conv3x3_halo");
}

// bn_y (nullable): dgrad only -- raw conv output of the layer whose activation gradient is being
// produced; with it scale/shift/bn_mean/bn_invstd are that layer's BatchNorm coefficients and
// stat_partials receives its backward sums (sum g, sum g*xhat).
static int conv3x3_dispatch(const void* x, long ldx, const void* w_packed, void* y, long ldy, int N,
                            int H, int W, int Cin, int Cout, const float* scale, const float* shift,
                            int relu, float* stat_partials, const void* bn_y, long ld_bn_y,
                            const float* bn_mean, const float* bn_invstd, cudaStream_t stream,
                            int taps = 9) {
  if (N <= 0 || H <= 0 || W <= 0) return FPB200_ERR_SHAPE;
  if (taps == 1 && (Cin % 64 != 0 || bn_y != nullptr)) return FPB200_ERR_SHAPE;
  if (Cin % 16 != 0 || Cout % 64 != 0 || Cout > 4096) return FPB200_ERR_SHAPE;
  if ((ldx % 8) != 0 || (ldy % 8) != 0 || ldx < Cin || ldy < Cout) return FPB200_ERR_SHAPE;
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(y) & 15) ||
      (reinterpret_cast<uintptr_t>(w_packed) & 15))
    return FPB200_ERR_ALIGN;
  if (scale != nullptr && Cout > 512) return FPB200_ERR_SHAPE;
  if (bn_y != nullptr && (!scale || !shift || !bn_mean || !bn_invstd || !stat_partials || ld_bn_y % 8 != 0 ||
                          ld_bn_y < Cout || (reinterpret_cast<uintptr_t>(bn_y) & 15)))
    return FPB200_ERR_SHAPE;
  const int KCH = (Cin % 64 == 0) ? 64 : ((Cin % 32 == 0) ? 32 : 16);
  const int BN = (Cout % 128 == 0) ? 128 : 64;

  HaloParams p;
  p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout;
  p.patches_w = (W + kPatch - 1) / kPatch;
  p.patches_h = (H + kPatch - 1) / kPatch;
  p.num_patches = N * p.patches_h * p.patches_w;
  p.num_n_blks = Cout / BN;
  p.y = reinterpret_cast<__nv_bfloat16*>(y);
  p.ldy = ldy;
  p.scale = scale; p.shift = shift; p.relu = relu;
  p.bn_mean = bn_mean; p.bn_invstd = bn_invstd;
  p.stats_mode = stat_partials == nullptr ? 0 : (bn_y != nullptr ? 2 : 1);
  p.stat_partials = stat_partials;

  CUtensorMap tmA, tmB, tmY;
  const int box = kPatch + (taps == 9 ? 2 : 0);
  int rc = make_tmap_act(&tmA, x, N, H, W, Cin, ldx, KCH, box, box);
  if (rc != FPB200_OK) return rc;
  rc = make_tmap_mat(&tmB, w_packed, Cout, (long)taps * Cin, KCH, BN);
  if (rc != FPB200_OK) return rc;
  rc = make_tmap_act(&tmY, y, N, H, W, Cout, ldy, 64, 8, 4);  // epilogue store box: 8 px x 4 rows x 64 ch
  if (rc != FPB200_OK) return rc;
  CUtensorMap tmYL = tmY;
  if (bn_y != nullptr) {
    rc = make_tmap_act(&tmYL, bn_y, N, H, W, Cout, ld_bn_y, 64, 8, 4);
    if (rc != FPB200_OK) return rc;
  }
  if (stat_partials != nullptr) {
    if (cudaMemsetAsync(stat_partials, 0, (size_t)fpb200_conv_stat_rows() * 2 * Cout * sizeof(float),
                        stream) != cudaSuccess)
      return check_launch("conv3x3 stat memset");
  }
  // one 64-channel K chunk under a 128-wide N tile: 36 MMAs per tile pair, so an epilogue that does
  // more than convert (statistics or affine) is the critical path with 4 warps -> 8 epilogue warps,
  // one per (quadrant, tile).  Measured (B = 64, 64 -> 128 @ 256^2 fprop with statistics): 0.619 ->
  // 0.476 ms; with two or more K chunks the shallower rings of this variant cost more than the
  // epilogue gains (128 -> 128: 0.834 -> 0.853 ms), and a plain dgrad epilogue has slack either way.
  if (BN == 128 && KCH == 64 && taps == 9 && Cin == 64 && (p.stats_mode != 0 || scale != nullptr)) {
    if (halo_ws_mask() & 2) return launch_halo<128, 64, 9, 8, true>(tmA, tmB, tmY, tmYL, p, stream);
    return launch_halo<128, 64, 9, 8>(tmA, tmB, tmY, tmYL, p, stream);
  }
  if (BN == 128 && KCH == 64 && taps == 9 && (halo_ws_mask() & 2))
    return launch_halo<128, 64, 9, 4, true>(tmA, tmB, tmY, tmYL, p, stream);
  if (BN == 64 && KCH == 64 && taps == 9 && (halo_ws_mask() & 1))
    return launch_halo<64, 64, 9, 8, true>(tmA, tmB, tmY, tmYL, p, stream);
#define FP_HALO_CASE(bn, kch, tp) \
  if (BN == bn && KCH == kch && taps == tp) return launch_halo<bn, kch, tp>(tmA, tmB, tmY, tmYL, p, stream);
  FP_HALO_CASE(128, 64, 9)
  FP_HALO_CASE(64, 64, 9)
  FP_HALO_CASE(128, 32, 9)
  FP_HALO_CASE(64, 32, 9)
  FP_HALO_CASE(128, 16, 9)
  FP_HALO_CASE(64, 16, 9)
  FP_HALO_CASE(128, 64, 1)
  FP_HALO_CASE(64, 64, 1)
#undef FP_HALO_CASE
  return FPB200_ERR_SHAPE;
}

}  // namespace fp

extern "C" {

int fpb200_conv_stat_rows(void) { return 8 * fp::sm_count(); }

int fpb200_conv3x3_fprop_bf16_nhwc(const void* x, long ldx, const void* w_packed, void* y, long ldy,
                                   int N, int H, int W, int Cin, int Cout, const float* scale,
                                   const float* shift, int relu, float* stat_partials,
                                   void* stream) {
  return fp::conv3x3_dispatch(x, ldx, w_packed, y, ldy, N, H, W, Cin, Cout, scale, shift, relu,
                              stat_partials, nullptr, 0, nullptr, nullptr,
                              static_cast<cudaStream_t>(stream));
}

int fpb200_conv3x3_dgrad_bf16_nhwc(const void* dy, long lddy, const void* w_packed_dgrad, void* dx,
                                   long lddx, int N, int H, int W, int Cout, int Cin,
                                   const void* bn_y, long ld_bn_y, const float* bn_scale,
                                   const float* bn_shift, const float* bn_mean,
                                   const float* bn_invstd, float* bn_partials, void* stream) {
  // dgrad of a 3x3/pad-1 conv is the same convolution with the channel roles swapped and
  // the filter rotated by 180 degrees; the rotation/transposition lives in the packing.
  return fp::conv3x3_dispatch(dy, lddy, w_packed_dgrad, dx, lddx, N, H, W, Cout, Cin,
                              bn_y ? bn_scale : nullptr, bn_y ? bn_shift : nullptr, 0,
                              bn_y ? bn_partials : nullptr, bn_y, ld_bn_y, bn_mean, bn_invstd,
                              static_cast<cudaStream_t>(stream));
}

/* pointwise (1x1) convolution: forward with w_packed = [Cout][Cin], data gradient with the
 * transposed packing [Cin][Cout] (then "Cin"/"Cout" swap roles, as for the 3x3 dgrad). */
int fpb200_conv1x1_bf16_nhwc(const void* x, long ldx, const void* w_packed, void* y, long ldy, int N,
                             int H, int W, int Cin, int Cout, const float* scale, const float* shift,
                             int relu, void* stream) {
  if ((scale == nullptr) != (shift == nullptr)) return FPB200_ERR_SHAPE;
  return fp::conv3x3_dispatch(x, ldx, w_packed, y, ldy, N, H, W, Cin, Cout, scale, shift, relu,
                              nullptr, nullptr, 0, nullptr, nullptr,
                              static_cast<cudaStream_t>(stream), 1);
}

}  // extern "C"
