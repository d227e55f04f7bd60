// TEST-ONLY cross-check, NOT part of the product library: the first-generation conv kernel (one TMA box
// per filter tap), built into tests/lib/libfpb200_crosscheck.so by floodplanet_code_b200.build
// .build_test_library() and loaded only by tests/test_kernels_gpu.py, which runs every fprop case
// through the product's halo kernel AND this independent implementation.
//
// 3x3 / pad 1 / stride 1 convolution as an implicit GEMM on the sm_100a tensor cores.
//
// Replaces what nn.Conv2d(k=3, padding=1) dispatches to in the reference
// (st_water_seg/models/unet.py:14,16) -- forward (fprop) and, with the transposed +
// rotated weight packing, the data gradient (dgrad, the autograd of the same line).
//
//   D[m, co] = sum_{tap, ci} A[m + tap, ci] * Wp[co, tap, ci]
//     m   = output pixel (n, h, w)            GEMM M = N*H*W, tiled 128 pixels = TH x TW patch
//     co  = output channel                    GEMM N = Cout,  tiled BN
//     tap = (r, s) in 3x3, ci = input chan.   GEMM K = 9*Cin, tiled KCH per pipeline stage
//
// Data movement: activations are NHWC bf16.  For every (tap, channel chunk) the TMA unit
// fetches the shifted TH x TW x KCH box straight into 128B/64B/32B-swizzled shared memory;
// the 1-pixel halo of the conv is the TMA out-of-bounds zero fill, so no im2col buffer and
// no boundary branches exist anywhere.  Weights [Cout][9*Cin] are a plain K-major matrix.
//
// Execution: persistent CTAs (one per SM), warp specialised:
//   warp 0    TMA producer           (one elected lane)
//   warp 1    tcgen05.mma issuer     (one elected lane) + TMEM allocator
//   warps 2-5 epilogue: tcgen05.ld accumulator -> registers -> fused epilogue -> global
// Accumulators live in TMEM, double buffered (2 x BN fp32 columns), so the epilogue of
// tile i overlaps the MMAs of tile i+1.
//
// Fused epilogues (runtime flags, warp-uniform):
//   * per-channel affine (+ReLU): eval-mode BatchNorm folded to scale/shift, or bias add
//   * BatchNorm batch-statistic partials: per-channel sum and sum of squares of the fp32
//     accumulators, reduced over the 32 rows of each warp with a shuffle transpose-reduce
//     and accumulated in registers across all tiles of the persistent CTA
//   * bf16 cast + 16-byte vector stores into an NHWC view with an arbitrary pixel pitch
//     (so the output can land directly inside a concat buffer).
#include "host_common.h"
#include "ptx.cuh"

namespace fp {
namespace v1 {

struct ConvParams {
  int N, H, W;
  int Cin;   // padded input channels (multiple of KCH)
  int Cout;  // multiple of BN
  int tw_log2;
  int tiles_w, tiles_h;
  int num_m_tiles, num_n_blks;
  __nv_bfloat16* y;
  long ldy;
  const float* scale;  // nullable
  const float* shift;  // nullable
  int relu;
  float* stat_partials;  // nullable; [gridDim.x*4][2][Cout]
};

constexpr int kBM = 128;
constexpr int kNumThreads = 192;

template <int BN, int KCH>
struct ConvCfg {
  static constexpr int kABytes = kBM * KCH * 2;
  static constexpr int kBBytes = BN * KCH * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kSmemBudget = 200 * 1024;
  static constexpr int kStagesRaw = kSmemBudget / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kTmemCols = (2 * BN <= 32) ? 32 : (2 * BN <= 64 ? 64 : (2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512)));
  // stage buffers + 1024 alignment slack + barriers/scale/shift
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256 + 2 * 512 * 4;
};

template <int BN, int KCH>
__global__ void __launch_bounds__(kNumThreads, 1)
conv3x3_igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const ConvParams p) {
  using Cfg = ConvCfg<BN, KCH>;
  constexpr int kStages = Cfg::kStages;
  constexpr uint32_t kSwz = KCH * 2;       // swizzle span in bytes == bytes per smem row
  constexpr uint32_t kSBO = 8 * KCH * 2;   // 8-row core-matrix group pitch
  constexpr uint32_t kIdesc = make_idesc_bf16(kBM, BN, 0, 0);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  // layout: [stages x (A | B)] [barriers 256 B] [scale 512 f32] [shift 512 f32]
  const uint32_t bar_base = smem_base + kStages * Cfg::kStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 64u + 8u * s; };
  auto tfull_bar = [&](int s) { return bar_base + 128u + 8u * s; };
  auto tempty_bar = [&](int s) { return bar_base + 144u + 8u * s; };
  const uint32_t tmem_slot = bar_base + 160u;
  float* s_scale = reinterpret_cast<float*>(smem_al + kStages * Cfg::kStageBytes + 256);
  float* s_shift = s_scale + 512;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int TW = 1 << p.tw_log2;
  const int num_tiles = p.num_m_tiles * p.num_n_blks;
  const int k_chunks = p.Cin / KCH;
  const int k_iters = 9 * k_chunks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  if (p.scale != nullptr) {
    for (int c = threadIdx.x; c < p.Cout; c += kNumThreads) {
      s_scale[c] = p.scale[c];
      s_shift[c] = p.shift[c];
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_al + kStages * Cfg::kStageBytes + 160);

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int n_blk = tile / p.num_m_tiles;
        const int m_tile = tile - n_blk * p.num_m_tiles;
        const int twi = m_tile % p.tiles_w;
        const int t2 = m_tile / p.tiles_w;
        const int thi = t2 % p.tiles_h;
        const int img = t2 / p.tiles_h;
        const int w0 = twi << p.tw_log2;
        const int h0 = thi * (kBM >> p.tw_log2);
        for (int tap = 0; tap < 9; ++tap) {
          const int r = tap / 3, s = tap - 3 * r;
          for (int kc = 0; kc < k_chunks; ++kc) {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            const uint32_t a_dst = smem_base + stage * Cfg::kStageBytes;
            const uint32_t b_dst = a_dst + Cfg::kABytes;
            mbar_arrive_expect_tx(full_bar(stage), Cfg::kStageBytes);
            tma_load_4d(a_dst, &tmA, full_bar(stage), kc * KCH, w0 + s - 1, h0 + r - 1, img);
            tma_load_2d(b_dst, &tmB, full_bar(stage), tap * p.Cin + kc * KCH, n_blk * BN);
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(tempty_bar(as), aphase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int ki = 0; ki < k_iters; ++ki) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t a_addr = smem_base + stage * Cfg::kStageBytes;
          const uint32_t b_addr = a_addr + Cfg::kABytes;
#pragma unroll
          for (int k = 0; k < KCH / 16; ++k) {
            const uint64_t da = make_smem_desc(a_addr + k * 32, 16, kSBO, kSwz);
            const uint64_t db = make_smem_desc(b_addr + k * 32, 16, kSBO, kSwz);
            umma_bf16(d_tmem, da, db, kIdesc, (ki | k) != 0 ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));  // frees the smem slot once these MMAs retire
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(tfull_bar(as));  // accumulator ready for the epilogue
      }
    }
  } else {
    // ===================== epilogue warps (2..5) =====================
    const int quad = warp & 3;  // TMEM lane quadrant this warp may touch
    const int row = quad * 32 + lane;
    const int ew = warp - 2;
    const bool do_stats = p.stat_partials != nullptr;
    const bool do_affine = p.scale != nullptr;
    float acc_sum[BN / 32], acc_sq[BN / 32];
#pragma unroll
    for (int c = 0; c < BN / 32; ++c) { acc_sum[c] = 0.f; acc_sq[c] = 0.f; }
    int cur_n_blk = -1;
    auto flush_stats = [&]() {
      if (do_stats && cur_n_blk >= 0) {
        float* dst = p.stat_partials + (size_t)(blockIdx.x * 4 + ew) * 2 * p.Cout + cur_n_blk * BN;
#pragma unroll
        for (int c = 0; c < BN / 32; ++c) {
          dst[c * 32 + lane] = acc_sum[c];
          dst[p.Cout + c * 32 + lane] = acc_sq[c];
          acc_sum[c] = 0.f;
          acc_sq[c] = 0.f;
        }
      }
    };
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int n_blk = tile / p.num_m_tiles;
      const int m_tile = tile - n_blk * p.num_m_tiles;
      const int twi = m_tile % p.tiles_w;
      const int t2 = m_tile / p.tiles_w;
      const int thi = t2 % p.tiles_h;
      const int img = t2 / p.tiles_h;
      const int pw = (twi << p.tw_log2) + (row & (TW - 1));
      const int ph = thi * (kBM >> p.tw_log2) + (row >> p.tw_log2);
      const bool valid = (pw < p.W) && (ph < p.H);
      if (n_blk != cur_n_blk) { flush_stats(); cur_n_blk = n_blk; }
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
      __nv_bfloat16* yrow = p.y + ((size_t)(img * p.H + ph) * p.W + pw) * p.ldy + n_blk * BN;
#pragma unroll
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + (uint32_t(quad * 32) << 16) + as * BN + c * 32, r);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        if (do_affine) {
          const int cb = n_blk * BN + c * 32;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            v[j] = fmaf(v[j], s_scale[cb + j], s_shift[cb + j]);
            if (p.relu) v[j] = fmaxf(v[j], 0.f);
          }
        }
        if (valid) {
          uint4* dst = reinterpret_cast<uint4*>(yrow + c * 32);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 o;
            o.x = pack_bf16x2(v[q * 8 + 0], v[q * 8 + 1]);
            o.y = pack_bf16x2(v[q * 8 + 2], v[q * 8 + 3]);
            o.z = pack_bf16x2(v[q * 8 + 4], v[q * 8 + 5]);
            o.w = pack_bf16x2(v[q * 8 + 6], v[q * 8 + 7]);
            dst[q] = o;
          }
        }
        if (do_stats) {
          float q[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            v[j] = valid ? v[j] : 0.f;
            q[j] = v[j] * v[j];
          }
          // transpose-reduce over the 32 lanes: lane l ends with column l's total
#pragma unroll
          for (int off = 16; off >= 1; off >>= 1) {
            const bool upper = (lane & off) != 0;
#pragma unroll
            for (int i = 0; i < off; ++i) {
              const float send = upper ? v[i] : v[i + off];
              const float keep = upper ? v[i + off] : v[i];
              v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
              const float send2 = upper ? q[i] : q[i + off];
              const float keep2 = upper ? q[i + off] : q[i];
              q[i] = keep2 + __shfl_xor_sync(0xffffffffu, send2, off);
            }
          }
          acc_sum[c] += v[0];
          acc_sq[c] += q[0];
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(as));
    }
    flush_stats();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

template <int BN, int KCH>
static int launch_conv(const CUtensorMap& tmA, const CUtensorMap& tmB, const ConvParams& p,
                       cudaStream_t stream) {
  using Cfg = ConvCfg<BN, KCH>;
  auto kern = conv3x3_igemm_kernel<BN, KCH>;
  static bool attr_set[kMaxDevices] = {false};   // cudaFuncSetAttribute is per device
  const int dev_ = current_device();
  if (!attr_set[dev_]) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes) !=
        cudaSuccess)
      return check_launch("conv3x3 smem attribute");
    attr_set[dev_] = true;
  }
  const int tiles = p.num_m_tiles * p.num_n_blks;
  int grid = sm_count();
  if (grid > tiles) grid = tiles;
  kern<<<grid, kNumThreads, Cfg::kSmemBytes, stream>>>(tmA, tmB, p);
  return check_launch("conv3x3_igemm");
}

int conv3x3_dispatch(const void* x, long ldx, const void* w_packed, void* y, long ldy, int N,
                            int H, int W, int Cin, int Cout, const float* scale, const float* shift,
                            int relu, float* stat_partials, cudaStream_t stream) {
  if (N <= 0 || H <= 0 || W <= 0) return FPB200_ERR_SHAPE;
  if (Cin % 16 != 0 || Cout % 64 != 0 || Cout > 2048) return FPB200_ERR_SHAPE;
  if ((ldx % 8) != 0 || (ldy % 8) != 0 || ldx < Cin || ldy < Cout) return FPB200_ERR_SHAPE;
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(y) & 15) ||
      (reinterpret_cast<uintptr_t>(w_packed) & 15))
    return FPB200_ERR_ALIGN;
  if (scale != nullptr && Cout > 512) return FPB200_ERR_SHAPE;
  const int KCH = (Cin % 64 == 0) ? 64 : ((Cin % 32 == 0) ? 32 : 16);
  const int BN = (Cout % 256 == 0) ? 256 : ((Cout % 128 == 0) ? 128 : 64);

  // tile geometry: choose TW minimising padded area (ties -> wider rows)
  int best_l = 3;
  long best_area = -1;
  for (int l = 3; l <= 7; ++l) {
    const int tw = 1 << l, th = kBM >> l;
    const long area = (long)((W + tw - 1) / tw) * tw * (long)((H + th - 1) / th) * th;
    if (best_area < 0 || area <= best_area) { best_area = area; best_l = l; }
  }
  ConvParams p;
  p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout;
  p.tw_log2 = best_l;
  const int TW = 1 << best_l, TH = kBM >> best_l;
  p.tiles_w = (W + TW - 1) / TW;
  p.tiles_h = (H + TH - 1) / TH;
  p.num_m_tiles = N * p.tiles_h * p.tiles_w;
  p.num_n_blks = Cout / BN;
  p.y = reinterpret_cast<__nv_bfloat16*>(y);
  p.ldy = ldy;
  p.scale = scale; p.shift = shift; p.relu = relu;
  p.stat_partials = stat_partials;

  CUtensorMap tmA, tmB;
  int rc = make_tmap_act(&tmA, x, N, H, W, Cin, ldx, KCH, TW, TH);
  if (rc != FPB200_OK) return rc;
  rc = make_tmap_mat(&tmB, w_packed, Cout, 9L * Cin, KCH, BN);
  if (rc != FPB200_OK) return rc;
  if (stat_partials != nullptr) {
    if (cudaMemsetAsync(stat_partials, 0, (size_t)(8 * sm_count()) * 2 * Cout * sizeof(float),
                        stream) != cudaSuccess)
      return check_launch("conv3x3 stat memset");
  }
#define FP_CONV_CASE(bn, kch) \
  if (BN == bn && KCH == kch) return launch_conv<bn, kch>(tmA, tmB, p, stream);
  FP_CONV_CASE(256, 64)
  FP_CONV_CASE(128, 64)
  FP_CONV_CASE(64, 64)
  FP_CONV_CASE(256, 32)
  FP_CONV_CASE(128, 32)
  FP_CONV_CASE(64, 32)
  FP_CONV_CASE(256, 16)
  FP_CONV_CASE(128, 16)
  FP_CONV_CASE(64, 16)
#undef FP_CONV_CASE
  return FPB200_ERR_SHAPE;
}

}  // namespace v1
}  // namespace fp

extern "C" int fpb200_test_conv3x3_pertap_bf16_nhwc(const void* x, long ldx, const void* w_packed, void* y,
                                                    long ldy, int N, int H, int W, int Cin, int Cout,
                                                    const float* scale, const float* shift, int relu,
                                                    float* stat_partials, void* stream) {
  return fp::v1::conv3x3_dispatch(x, ldx, w_packed, y, ldy, N, H, W, Cin, Cout, scale, shift, relu,
                                  stat_partials, static_cast<cudaStream_t>(stream));
}
