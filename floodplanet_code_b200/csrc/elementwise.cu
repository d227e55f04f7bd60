// Bandwidth-bound kernels of the UNet hot path: layout ingest, weight repack, BatchNorm
// finalise / apply (+ReLU, +MaxPool2d) and its backward, max-pool backward, bilinear
// upsample + pad + concat (forward and gather-form backward) and Adam.
//
// All activation traffic is NHWC bf16 moved as 16-byte vectors (8 channels per access), one
// read and one write per element; per-channel coefficients are fp32 and amortised over
// several pixels per thread.  Reductions are two-stage (per-block partials, then a tiny
// finalise kernel summing in fp64) so results are deterministic run to run.
//
// Reference lines each kernel replaces are cited in include/floodplanet_b200.h.
#include "host_common.h"
#include "ptx.cuh"

namespace fp {

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x);
  f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z);
  f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]);
  u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]);
  u.w = pack_bf16x2(f[6], f[7]);
  return u;
}
__device__ __forceinline__ void load8f(const float* p, float (&f)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
  f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ uint4 ld_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

static inline int grid_for(long work, int block, int max_blocks) {
  long g = (work + block - 1) / block;
  if (g > max_blocks) g = max_blocks;
  if (g < 1) g = 1;
  return (int)g;
}

// ---------------------------------------------------------------------------
// ingest: NCHW f32 (x n_src, concatenated along C) -> NHWC bf16, C zero-padded
// ---------------------------------------------------------------------------
constexpr int kIngestMaxC = 256;   // padded input channels of the first convolution (early fusion of many sensors)
struct IngestArgs {
  const float* src[8];
  unsigned char ch_src[kIngestMaxC];  // for padded channel c: which source
  unsigned char ch_idx[kIngestMaxC];  // ... and which channel inside it (255 = zero pad)
  int src_c[8];
};

__global__ void ingest_kernel(IngestArgs a, __nv_bfloat16* __restrict__ dst, int c_pad, int N,
                              long hw) {
  const long total = (long)N * hw;
  for (long px = blockIdx.x * (long)blockDim.x + threadIdx.x; px < total;
       px += (long)gridDim.x * blockDim.x) {
    const long n = px / hw;
    const long o = px - n * hw;
    for (int g = 0; g < c_pad; g += 8) {
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = g + j;
        const int ci = a.ch_idx[c];
        float v = 0.f;
        if (ci != 255) {
          const int s = a.ch_src[c];
          v = __ldg(a.src[s] + ((long)n * a.src_c[s] + ci) * hw + o);
        }
        f[j] = v;
      }
      *reinterpret_cast<uint4*>(dst + px * c_pad + g) = pack8(f);
    }
  }
}

// Tiles of one resident scene [C][H][W] f32 -> NHWC bf16 [n][th][tw][c_pad]; pixels outside the
// scene or outside the tile's valid extent are zero (the dataset pads remainder crops).
__global__ void ingest_scene_tiles_kernel(const float* __restrict__ scene, int C, long H, long W,
                                          const int* __restrict__ tiles, int n, int th, int tw,
                                          __nv_bfloat16* __restrict__ dst, int c_pad) {
  const long hw = (long)th * tw;
  const long total = (long)n * hw;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total;
       i += (long)gridDim.x * blockDim.x) {
    const int t = (int)(i / hw);
    const long o = i - (long)t * hw;
    const int y = (int)(o / tw), x = (int)(o - (long)y * tw);
    const int h0 = tiles[t * 4 + 0], w0 = tiles[t * 4 + 1], hh = tiles[t * 4 + 2], ww = tiles[t * 4 + 3];
    const long gy = h0 + y, gx = w0 + x;
    const bool in = y < hh && x < ww && gy < H && gx < W;
    for (int g = 0; g < c_pad; g += 8) {
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = g + j;
        f[j] = (in && c < C) ? __ldg(scene + ((long)c * H + gy) * W + gx) : 0.f;
      }
      *reinterpret_cast<uint4*>(dst + i * c_pad + g) = pack8(f);
    }
  }
}

// ---------------------------------------------------------------------------
// weight repack
// ---------------------------------------------------------------------------
__global__ void repack_fprop_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out,
                                    int Cout, int Cin, int cin_pad) {
  const long total = (long)Cout * 9 * cin_pad;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total;
       i += (long)gridDim.x * blockDim.x) {
    const int ci = i % cin_pad;
    const int tap = (i / cin_pad) % 9;
    const int co = i / (9L * cin_pad);
    const float v = ci < Cin ? w[((long)co * Cin + ci) * 9 + tap] : 0.f;
    out[i] = __float2bfloat16(v);
  }
}
// out[ci][tap'][co] = w[co][ci][2-r'][2-s']  (tap' = 3 r' + s'  ->  source tap = 8 - tap')
__global__ void repack_dgrad_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out,
                                    int Cout, int Cin) {
  const long total = (long)Cin * 9 * Cout;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total;
       i += (long)gridDim.x * blockDim.x) {
    const int co = i % Cout;
    const int tap = (i / Cout) % 9;
    const int ci = i / (9L * Cout);
    out[i] = __float2bfloat16(w[((long)co * Cin + ci) * 9 + (8 - tap)]);
  }
}

// All 3x3 operand copies of a model in ONE launch (after an optimiser step every master weight
// changed: 18 + 15 per-layer launches of a few microseconds each became one).  Each table entry is one
// packed buffer and owns a contiguous range of blocks: kind 0 (fprop layout [co][tap][ci_pad]) one
// block per output channel, which stages w[co][:][:] (contiguous) in shared memory and writes the
// transposed rows; kind 1 (dgrad layout [ci][tap'][co], tap' = 8 - tap) one block per input channel,
// whose threads read the 9 contiguous taps of (co, ci) and write co-contiguous rows.
struct RepackEntry {
  const float* w;        // fp32 OIHW master
  __nv_bfloat16* out;    // packed copy
  int cout, cin, cin_pad, kind;
  long first_block;      // first block of this entry; entries are sorted by it
};
constexpr int kRepackMaxCin = 1024;   // kind 0 stages Cin * 9 floats

__global__ void __launch_bounds__(256)
repack_batch_kernel(const RepackEntry* __restrict__ table, int n_entries) {
  __shared__ float s_w[kRepackMaxCin * 9];
  __shared__ int s_e;
  if (threadIdx.x == 0) {
    int e = 0;
    while (e + 1 < n_entries && table[e + 1].first_block <= (long)blockIdx.x) ++e;
    s_e = e;
  }
  __syncthreads();
  const RepackEntry en = table[s_e];
  const int unit = (int)((long)blockIdx.x - en.first_block);
  if (en.kind == 0) {
    const int co = unit;
    const float* src = en.w + (long)co * en.cin * 9;
    for (int i = threadIdx.x; i < en.cin * 9; i += blockDim.x) s_w[i] = src[i];
    __syncthreads();
    __nv_bfloat16* dst = en.out + (long)co * 9 * en.cin_pad;
    for (int i = threadIdx.x; i < 9 * en.cin_pad; i += blockDim.x) {
      const int tap = i / en.cin_pad, ci = i - tap * en.cin_pad;
      dst[i] = __float2bfloat16(ci < en.cin ? s_w[ci * 9 + tap] : 0.f);
    }
  } else {
    const int ci = unit;
    __nv_bfloat16* dst = en.out + (long)ci * 9 * en.cout;
    for (int co = threadIdx.x; co < en.cout; co += blockDim.x) {
      const float* src = en.w + ((long)co * en.cin + ci) * 9;
      float t[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) t[k] = src[k];
#pragma unroll
      for (int k = 0; k < 9; ++k) dst[(long)k * en.cout + co] = __float2bfloat16(t[8 - k]);
    }
  }
}

// ---------------------------------------------------------------------------
// BatchNorm statistics finalise (32 channels per block, fp64 accumulation)
// ---------------------------------------------------------------------------

// Column sums of a [P][2][C] fp32 partial array for 32 consecutive channels per block: lane = channel
// (every warp load is one coalesced 128-byte row segment), the 32 warps of the block stride over the
// P rows with all their loads independent, fp64 accumulation, one shared-memory pass across warps.
// (The first version used one warp per channel with the lanes striding over rows: 32 sectors per
// request and a serial dependent loop, 14-18 us per layer on the critical path between a conv and
// its normalisation pass; ncu, C = 64..512, P = 1184.)
constexpr int kFinThreads = 1024;

__device__ __forceinline__ bool column_sums_32(const float* __restrict__ partials, int P, int C,
                                               double& s, double& q) {
  __shared__ double red[2][kFinThreads / 32][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
  if (c < C) {
    int p = warp;
    for (; p + (kFinThreads / 32) < P; p += 2 * (kFinThreads / 32)) {
      const float* r0 = partials + (size_t)p * 2 * C + c;
      const float* r1 = r0 + (size_t)(kFinThreads / 32) * 2 * C;
      const float x0 = r0[0], y0 = r0[C], x1 = r1[0], y1 = r1[C];
      a0 += (double)x0; b0 += (double)y0; a1 += (double)x1; b1 += (double)y1;
    }
    if (p < P) {
      const float* r0 = partials + (size_t)p * 2 * C + c;
      a0 += (double)r0[0]; b0 += (double)r0[C];
    }
  }
  red[0][warp][lane] = a0 + a1;
  red[1][warp][lane] = b0 + b1;
  __syncthreads();
  if (warp != 0 || c >= C) return false;
  s = 0.0; q = 0.0;
  for (int w = 0; w < kFinThreads / 32; ++w) { s += red[0][w][lane]; q += red[1][w][lane]; }
  return true;
}

__global__ void __launch_bounds__(kFinThreads)
bn_stats_finalize_kernel(const float* __restrict__ partials, int P, int C,
                         double count, const float* __restrict__ gamma,
                         const float* __restrict__ beta,
                         const float* __restrict__ conv_bias, float eps,
                         float momentum, float* running_mean, float* running_var,
                         float* scale, float* shift, float* save_mean,
                         float* save_invstd) {
  double s, q;
  if (!column_sums_32(partials, P, C, s, q)) return;
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const double mean = s / count;
  double var = q / count - mean * mean;
  if (var < 0.0) var = 0.0;
  const float invstd = (float)(1.0 / sqrt(var + (double)eps));
  const float sc = gamma[c] * invstd;
  scale[c] = sc;
  shift[c] = beta[c] - (float)mean * sc;
  if (save_mean) save_mean[c] = (float)mean;
  if (save_invstd) save_invstd[c] = invstd;
  if (running_mean) {
    const float m_full = (float)mean + (conv_bias ? conv_bias[c] : 0.f);
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * m_full;
  }
  if (running_var) {
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

__global__ void bn_fold_eval_kernel(const float* gamma, const float* beta, const float* conv_bias,
                                    const float* rm, const float* rv, float eps, int C,
                                    float* scale, float* shift) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float sc = gamma[c] / sqrtf(rv[c] + eps);
  scale[c] = sc;
  shift[c] = beta[c] + ((conv_bias ? conv_bias[c] : 0.f) - rm[c]) * sc;
}

// eval-mode ("frozen") BatchNorm statistics in the form the backward kernels take: the conv bias is not in
// the GEMM, so xhat = (y + bias - running_mean) * invstd = (y - mean_eff) * invstd
__global__ void bn_eval_stats_kernel(const float* conv_bias, const float* rm, const float* rv, float eps, int C,
                                     float* mean_eff, float* invstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  mean_eff[c] = rm[c] - (conv_bias ? conv_bias[c] : 0.f);
  invstd[c] = 1.f / sqrtf(rv[c] + eps);
}

// ---------------------------------------------------------------------------
// BN apply + ReLU
// ---------------------------------------------------------------------------
constexpr int kPixPerThread = 4;

__global__ void bn_apply_relu_kernel(const __nv_bfloat16* __restrict__ y, long ldy,
                                     __nv_bfloat16* __restrict__ a, long lda,
                                     const float* __restrict__ scale,
                                     const float* __restrict__ shift, long num_pixels, int CG) {
  // work item = (pixel group of kPixPerThread pixels strided by `pstride`, channel group)
  const long pgroups = (num_pixels + kPixPerThread - 1) / kPixPerThread;
  const long total = pgroups * CG;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total;
       i += (long)gridDim.x * blockDim.x) {
    const int cg = i % CG;
    const long pg = i / CG;
    float sc[8], sh[8];
    load8f(scale + cg * 8, sc);
    load8f(shift + cg * 8, sh);
    uint4 in[kPixPerThread];
#pragma unroll
    for (int k = 0; k < kPixPerThread; ++k) {
      const long px = pg + k * pgroups;
      if (px < num_pixels) in[k] = ld_stream(y + px * ldy + cg * 8);
    }
#pragma unroll
    for (int k = 0; k < kPixPerThread; ++k) {
      const long px = pg + k * pgroups;
      if (px < num_pixels) {
        float f[8];
        unpack8(in[k], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = fmaxf(fmaf(f[j], sc[j], sh[j]), 0.f);
        *reinterpret_cast<uint4*>(a + px * lda + cg * 8) = pack8(f);
      }
    }
  }
}

// BN apply + ReLU fused with MaxPool2d(2).  One block = one row of 2x2 windows (n, ho); one
// thread = one window x 8 channels; 32-bit index math.
__global__ void __launch_bounds__(256)
bn_apply_relu_maxpool2_kernel(const __nv_bfloat16* __restrict__ y, long ldy,
                              __nv_bfloat16* __restrict__ a, long lda,
                              __nv_bfloat16* __restrict__ pooled, long ldp,
                              uint8_t* __restrict__ pool_idx, const float* __restrict__ scale,
                              const float* __restrict__ shift, int H, int W, int CG) {
  const int Hc = (H + 1) >> 1, Wc = (W + 1) >> 1;  // windows incl. the ragged last row/col
  const int Hp = H >> 1, Wp = W >> 1;
  const int n = blockIdx.x / Hc;
  const int ho = blockIdx.x - n * Hc;
  const bool affine = scale != nullptr;
  const int work = Wc * CG;
  for (int i = threadIdx.x; i < work; i += blockDim.x) {
    const int wo = i / CG;
    const int cg = i - wo * CG;
    float sc[8], sh[8];
    if (affine) {
      load8f(scale + cg * 8, sc);
      load8f(shift + cg * 8, sh);
    }
    uint4 in[4];
    bool ok[4];
#pragma unroll
    for (int pos = 0; pos < 4; ++pos) {
      const int h = 2 * ho + (pos >> 1), w = 2 * wo + (pos & 1);
      ok[pos] = h < H && w < W;
      if (ok[pos]) in[pos] = ld_stream(y + (((long)n * H + h) * W + w) * ldy + cg * 8);
    }
    float best[8];
    int bidx[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { best[j] = -INFINITY; bidx[j] = 0; }
#pragma unroll
    for (int pos = 0; pos < 4; ++pos) {
      if (ok[pos]) {
        const int h = 2 * ho + (pos >> 1), w = 2 * wo + (pos & 1);
        float f[8];
        unpack8(in[pos], f);
        if (affine) {
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = fmaxf(fmaf(f[j], sc[j], sh[j]), 0.f);
          // round through bf16 so the pooled value equals the stored activation bit for bit
          const uint4 packed = pack8(f);
          if (a != nullptr)
            *reinterpret_cast<uint4*>(a + (((long)n * H + h) * W + w) * lda + cg * 8) = packed;
          unpack8(packed, f);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (f[j] > best[j] || f[j] != f[j]) { best[j] = f[j]; bidx[j] = pos; }
        }
      }
    }
    if (ho < Hp && wo < Wp) {
      const long pp = ((long)n * Hp + ho) * Wp + wo;
      *reinterpret_cast<uint4*>(pooled + pp * ldp + cg * 8) = pack8(best);
      uint2 ib;
      ib.x = bidx[0] | (bidx[1] << 8) | (bidx[2] << 16) | (bidx[3] << 24);
      ib.y = bidx[4] | (bidx[5] << 8) | (bidx[6] << 16) | (bidx[7] << 24);
      *reinterpret_cast<uint2*>(pool_idx + pp * (long)(CG * 8) + cg * 8) = ib;
    }
  }
}

// FUSED: dx is the activation gradient of a conv+BN+ReLU layer (the skip tensor), so that
// layer's BatchNorm-backward sums (sum g, sum g*xhat, g = dx*[y*scale+shift > 0]) are reduced
// right here, where dx is produced, from one extra read of y -- the separate reduction pass
// (which would read dx and y again) is skipped.  Requires CG | 256, so a thread keeps one
// channel group for all its windows; blocks walk window rows with a grid stride.
template <bool FUSED>
__global__ void __launch_bounds__(256)
maxpool2_bwd_kernel(const __nv_bfloat16* __restrict__ dpooled, long lddp,
                    const uint8_t* __restrict__ pool_idx, const __nv_bfloat16* __restrict__ dskip,
                    long ldds, __nv_bfloat16* __restrict__ dx, long lddx, int N, int H, int W, int CG,
                    const __nv_bfloat16* __restrict__ y, long ldy, const float* __restrict__ scale,
                    const float* __restrict__ shift, const float* __restrict__ mean,
                    const float* __restrict__ invstd, float* __restrict__ partials) {
  __shared__ float red[FUSED ? 256 * 16 : 1];
  const int Hc = (H + 1) >> 1, Wc = (W + 1) >> 1;
  const int Hp = H >> 1, Wp = W >> 1;
  const int work = Wc * CG;
  float sc[8], sh[8], mu[8], is[8], s1[8], s2[8];
  if (FUSED) {
    const int cg0 = threadIdx.x % CG;
    load8f(scale + cg0 * 8, sc);
    load8f(shift + cg0 * 8, sh);
    load8f(mean + cg0 * 8, mu);
    load8f(invstd + cg0 * 8, is);
#pragma unroll
    for (int j = 0; j < 8; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
  }
  for (int rowi = blockIdx.x; rowi < N * Hc; rowi += gridDim.x) {
    const int n = rowi / Hc;
    const int ho = rowi - n * Hc;
    for (int i = threadIdx.x; i < work; i += blockDim.x) {
      const int wo = i / CG;
      const int cg = i - wo * CG;
      const bool has_pool = ho < Hp && wo < Wp;
      float g[8];
      uint2 ib = make_uint2(0, 0);
      uint4 sk[4], yv[4];
      bool ok[4];
      if (has_pool) {
        const long pp = ((long)n * Hp + ho) * Wp + wo;
        unpack8(ld_stream(dpooled + pp * lddp + cg * 8), g);
        ib = *reinterpret_cast<const uint2*>(pool_idx + pp * (long)(CG * 8) + cg * 8);
      }
#pragma unroll
      for (int pos = 0; pos < 4; ++pos) {
        const int h = 2 * ho + (pos >> 1), w = 2 * wo + (pos & 1);
        ok[pos] = h < H && w < W;
        if (ok[pos]) {
          const long px = ((long)n * H + h) * W + w;
          if (dskip != nullptr) sk[pos] = ld_stream(dskip + px * ldds + cg * 8);
          if (FUSED) yv[pos] = ld_stream(y + px * ldy + cg * 8);
        }
      }
#pragma unroll
      for (int pos = 0; pos < 4; ++pos) {
        if (ok[pos]) {
          const int h = 2 * ho + (pos >> 1), w = 2 * wo + (pos & 1);
          float f[8];
          if (dskip != nullptr) {
            unpack8(sk[pos], f);
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = 0.f;
          }
          if (has_pool) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint32_t word = j < 4 ? ib.x : ib.y;
              const int sel = (word >> (8 * (j & 3))) & 0xFF;
              if (sel == pos) f[j] += g[j];
            }
          }
          const uint4 packed = pack8(f);
          *reinterpret_cast<uint4*>(dx + (((long)n * H + h) * W + w) * lddx + cg * 8) = packed;
          if (FUSED) {
            float gr[8], v[8];
            unpack8(packed, gr);          // the bf16 values the apply pass will read
            unpack8(yv[pos], v);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float gg = fmaf(v[j], sc[j], sh[j]) > 0.f ? gr[j] : 0.f;
              s1[j] += gg;
              s2[j] = fmaf(gg, (v[j] - mu[j]) * is[j], s2[j]);
            }
          }
        }
      }
    }
  }
  if (FUSED) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      red[threadIdx.x * 16 + j] = s1[j];
      red[threadIdx.x * 16 + 8 + j] = s2[j];
    }
    __syncthreads();
    const int TP = 256 / CG;
    const int C = CG * 8;
    for (int o = threadIdx.x; o < CG * 16; o += 256) {
      const int ocg = o >> 4, oj = o & 15;
      float acc = 0.f;
      for (int pidx = 0; pidx < TP; ++pidx) acc += red[(pidx * CG + ocg) * 16 + oj];
      partials[(size_t)blockIdx.x * 2 * C + (oj >> 3) * C + ocg * 8 + (oj & 7)] = acc;
    }
  }
}

// ---------------------------------------------------------------------------
// BN + ReLU backward
// ---------------------------------------------------------------------------
constexpr int kBnBwdThreads = 256;

__global__ void __launch_bounds__(kBnBwdThreads)
bn_relu_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ da, long ldda,
                          const __nv_bfloat16* __restrict__ y, long ldy,
                          const float* __restrict__ scale, const float* __restrict__ shift,
                          const float* __restrict__ mean, const float* __restrict__ invstd,
                          float* __restrict__ partials, long num_pixels, int C) {
  __shared__ float red[kBnBwdThreads * 16];
  const int CG = C >> 3;            // channel groups (8..64), divides 256
  const int TP = kBnBwdThreads / CG;  // pixel lanes per block
  const int cg = threadIdx.x % CG;
  const int pl = threadIdx.x / CG;
  float sc[8], sh[8], mu[8], is[8];
  load8f(scale + cg * 8, sc);
  load8f(shift + cg * 8, sh);
  load8f(mean + cg * 8, mu);
  load8f(invstd + cg * 8, is);
  float s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
  // four pixels per iteration, their eight 16-byte loads issued before the first use (one pixel per iteration left
  // two loads in flight per thread: 5.15 TB/s); the pixels are still accumulated in the same order, so the sums keep
  // their bits
  const long stride = (long)gridDim.x * TP;
  for (long px0 = (long)blockIdx.x * TP + pl; px0 < num_pixels; px0 += 4 * stride) {
    uint4 gq[4], yq[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long px = px0 + k * stride;
      if (px < num_pixels) {
        gq[k] = ld_stream(da + px * ldda + cg * 8);
        yq[k] = ld_stream(y + px * ldy + cg * 8);
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (px0 + k * stride < num_pixels) {
        float g[8], v[8];
        unpack8(gq[k], g);
        unpack8(yq[k], v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float gg = fmaf(v[j], sc[j], sh[j]) > 0.f ? g[j] : 0.f;
          s1[j] += gg;
          s2[j] = fmaf(gg, (v[j] - mu[j]) * is[j], s2[j]);
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    red[threadIdx.x * 16 + j] = s1[j];
    red[threadIdx.x * 16 + 8 + j] = s2[j];
  }
  __syncthreads();
  // thread (cg, j16) sums over the TP pixel lanes
  for (int o = threadIdx.x; o < CG * 16; o += kBnBwdThreads) {
    const int ocg = o >> 4, oj = o & 15;
    float acc = 0.f;
    for (int p = 0; p < TP; ++p) acc += red[(p * CG + ocg) * 16 + oj];
    const int which = oj >> 3, ch = ocg * 8 + (oj & 7);
    partials[(size_t)blockIdx.x * 2 * C + which * C + ch] = acc;
  }
}

__global__ void __launch_bounds__(kFinThreads)
bn_bwd_finalize_kernel(const float* __restrict__ partials, int P, int C,
                       double count, const float* __restrict__ scale,
                       const float* __restrict__ mean,
                       const float* __restrict__ invstd, float* dgamma,
                       float* dbeta, float* coef) {
  double s, q;
  if (!column_sums_32(partials, P, C, s, q)) return;
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  if (dbeta) dbeta[c] = (float)s;
  if (dgamma) dgamma[c] = (float)q;
  // dy = scale*(g - c1 - xhat*c2) = scale*g - P*y - Q
  const double c1 = s / count, c2 = q / count;
  const double Pc = (double)scale[c] * c2 * (double)invstd[c];
  const double Qc = (double)scale[c] * c1 - Pc * (double)mean[c];
  coef[c] = (float)Pc;
  coef[C + c] = (float)Qc;
}

// Backward through an eval-mode BatchNorm (running statistics are constants): dy = scale * g, so the apply
// coefficients are zero; dgamma = sum g*xhat, dbeta = sum g, and the conv bias (no longer cancelled by a batch
// mean) gets dbias = sum dy = scale * sum g.
__global__ void __launch_bounds__(kFinThreads)
bn_bwd_finalize_frozen_kernel(const float* __restrict__ partials, int P, int C,
                              const float* __restrict__ scale, float* dgamma, float* dbeta, float* dbias,
                              float* coef) {
  double s, q;
  if (!column_sums_32(partials, P, C, s, q)) return;
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  if (dbeta) dbeta[c] = (float)s;
  if (dgamma) dgamma[c] = (float)q;
  if (dbias) dbias[c] = (float)((double)scale[c] * s);
  coef[c] = 0.f;
  coef[C + c] = 0.f;
}

__global__ void bn_relu_bwd_apply_kernel(const __nv_bfloat16* __restrict__ da, long ldda,
                                         const __nv_bfloat16* __restrict__ y, long ldy,
                                         __nv_bfloat16* __restrict__ dy, long lddy,
                                         const float* __restrict__ scale,
                                         const float* __restrict__ shift,
                                         const float* __restrict__ coef, long num_pixels, int CG) {
  const int C = CG * 8;
  const long pgroups = (num_pixels + kPixPerThread - 1) / kPixPerThread;
  const long total = pgroups * CG;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total;
       i += (long)gridDim.x * blockDim.x) {
    const int cg = i % CG;
    const long pg = i / CG;
    float sc[8], sh[8], Pc[8], Qc[8];
    load8f(scale + cg * 8, sc);
    load8f(shift + cg * 8, sh);
    load8f(coef + cg * 8, Pc);
    load8f(coef + C + cg * 8, Qc);
    uint4 gin[kPixPerThread], yin[kPixPerThread];
#pragma unroll
    for (int k = 0; k < kPixPerThread; ++k) {
      const long px = pg + k * pgroups;
      if (px < num_pixels) {
        gin[k] = ld_stream(da + px * ldda + cg * 8);
        yin[k] = ld_stream(y + px * ldy + cg * 8);
      }
    }
#pragma unroll
    for (int k = 0; k < kPixPerThread; ++k) {
      const long px = pg + k * pgroups;
      if (px < num_pixels) {
        float g[8], v[8], o[8];
        unpack8(gin[k], g);
        unpack8(yin[k], v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float gg = fmaf(v[j], sc[j], sh[j]) > 0.f ? g[j] : 0.f;
          o[j] = fmaf(sc[j], gg, -fmaf(Pc[j], v[j], Qc[j]));
        }
        *reinterpret_cast<uint4*>(dy + px * lddy + cg * 8) = pack8(o);
      }
    }
  }
}

// ---------------------------------------------------------------------------
// bilinear x2 upsample (align_corners=True) + zero pad, written into a concat view
// ---------------------------------------------------------------------------
__device__ __forceinline__ void bilinear_src(int dst, float ratio, int in_size, int& i0, int& i1,
                                             float& l0, float& l1) {
  const float src = ratio * (float)dst;
  i0 = (int)src;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = src - (float)i0;
  l0 = 1.f - l1;
}

// One block per group of kUpRows output rows of one image.  The group's output rows read at most
// kUpSrcRows consecutive source rows (ratio < 1/2), so each thread interpolates those source rows
// HORIZONTALLY once (2 loads per source row instead of 4 per output row) and then mixes them
// vertically with block-uniform taps; consecutive threads write consecutive 16-byte channel
// groups, a thread keeps its channel group for the whole block (no per-iteration division).
// Interior groups of the x2 / align_corners pattern (output rows 2j-1 .. 2j+2 read source rows
// (j-1,j) (j,j+1) (j,j+1) (j+1,j+2)) take a fully static path; the taps are still the ones
// bilinear_src computes, the pattern is only CHECKED, so edge groups, padding rows and any
// rounding surprise fall back to the generic walk.  History (ncu, 64ch 256->512, batch 64):
// 4 loads per output vector 0.83 ms, issue-active 78 %; separable 0.69 ms, latency bound;
// loads hoisted 0.61 ms, 121 instructions per stored vector, half of them tap bookkeeping.
constexpr int kUpRows = 4;
constexpr int kUpSrcRows = 4;
constexpr int kUpThreads = 256;

__device__ __forceinline__ void hlerp8(const uint4& a, const uint4& b, float lw0, float lw1,
                                       float (&o)[8]) {
  float v0[8], v1[8];
  unpack8(a, v0);
  unpack8(b, v1);
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j] = lw0 * v0[j] + lw1 * v1[j];
}

// lh0 * top + lh1 * bottom: same expression order as ATen's upsample_bilinear2d
__device__ __forceinline__ uint4 vlerp8(const float (&top)[8], const float (&bot)[8], float lh0,
                                        float lh1) {
  float o[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j] = fmaf(lh1, bot[j], lh0 * top[j]);
  return pack8(o);
}

__global__ void __launch_bounds__(kUpThreads)
upsample2x_pad_fwd_kernel(const __nv_bfloat16* __restrict__ x, long ldx,
                          __nv_bfloat16* __restrict__ out, long ldo, int h, int w, int Ho, int Wo,
                          int CG, int pad_top, int pad_left, float rh, float rw) {
  const int groups = (Ho + kUpRows - 1) / kUpRows;
  const int n = blockIdx.x / groups;
  const int ho0 = (blockIdx.x - n * groups) * kUpRows;
  const __nv_bfloat16* xin = x + (long)n * h * w * ldx;
  // block-uniform vertical taps of output row ho0 + r: source rows base + k0[r] / base + k1[r]
  int k0[kUpRows], k1[kUpRows];
  float lh0[kUpRows], lh1[kUpRows];
  int base = -1;
  unsigned row_in = 0, src_used = 0;
#pragma unroll
  for (int r = 0; r < kUpRows; ++r) {
    const int uh = ho0 + r - pad_top;
    k0[r] = k1[r] = 0;
    lh0[r] = lh1[r] = 0.f;
    if (ho0 + r < Ho && uh >= 0 && uh < 2 * h) {
      int h0, h1;
      bilinear_src(uh, rh, h, h0, h1, lh0[r], lh1[r]);
      if (base < 0) base = h0;
      k0[r] = h0 - base;
      k1[r] = h1 - base;
      row_in |= 1u << r;
      src_used |= (1u << k0[r]) | (1u << k1[r]);
    }
  }
  const bool interior = row_in == 0xFu && k0[0] == 0 && k1[0] == 1 && k0[1] == 1 && k1[1] == 2 &&
                        k0[2] == 1 && k1[2] == 2 && k0[3] == 2 && k1[3] == 3;
  const int cols_per_pass = kUpThreads / CG;
  const int cg = threadIdx.x % CG;
  const int col0 = threadIdx.x / CG;
  if (col0 >= cols_per_pass) return;
  const long ostride = (long)Wo * ldo;
  for (int wo = col0; wo < Wo; wo += cols_per_pass) {
    const int uw = wo - pad_left;
    const bool col_in = uw >= 0 && uw < 2 * w;
    __nv_bfloat16* orow = out + (((long)n * Ho + ho0) * Wo + wo) * ldo + cg * 8;
    if (!col_in || row_in == 0) {
#pragma unroll
      for (int r = 0; r < kUpRows; ++r)
        if (ho0 + r < Ho) *reinterpret_cast<uint4*>(orow + r * ostride) = make_uint4(0, 0, 0, 0);
      continue;
    }
    int w0, w1;
    float lw0, lw1;
    bilinear_src(uw, rw, w, w0, w1, lw0, lw1);
    // all source vectors of the group are requested before the first use (8 independent 16-byte
    // loads in flight per thread; rows beyond the image are clamped onto the last row = an L1 hit)
    uint4 q0[kUpSrcRows], q1[kUpSrcRows];
#pragma unroll
    for (int k = 0; k < kUpSrcRows; ++k) {
      const int hk = min(base + k, h - 1);
      const __nv_bfloat16* rp = xin + (long)hk * w * ldx + cg * 8;
      q0[k] = *reinterpret_cast<const uint4*>(rp + (long)w0 * ldx);
      q1[k] = *reinterpret_cast<const uint4*>(rp + (long)w1 * ldx);
    }
    if (interior) {
      float ha[8], hb[8];
      hlerp8(q0[0], q1[0], lw0, lw1, ha);
      hlerp8(q0[1], q1[1], lw0, lw1, hb);
      *reinterpret_cast<uint4*>(orow) = vlerp8(ha, hb, lh0[0], lh1[0]);
      hlerp8(q0[2], q1[2], lw0, lw1, ha);
      *reinterpret_cast<uint4*>(orow + ostride) = vlerp8(hb, ha, lh0[1], lh1[1]);
      *reinterpret_cast<uint4*>(orow + 2 * ostride) = vlerp8(hb, ha, lh0[2], lh1[2]);
      hlerp8(q0[3], q1[3], lw0, lw1, hb);
      *reinterpret_cast<uint4*>(orow + 3 * ostride) = vlerp8(ha, hb, lh0[3], lh1[3]);
      continue;
    }
    // generic walk over the source rows, keeping the previous horizontally-interpolated row: an
    // output row is emitted when its lower tap has been computed (taps chosen by block-uniform
    // compares, so everything stays in statically indexed registers)
    float prev[8], cur[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) prev[j] = cur[j] = 0.f;
#pragma unroll
    for (int k = 0; k < kUpSrcRows; ++k) {
      if (!(src_used & (1u << k))) continue;       // block-uniform
#pragma unroll
      for (int j = 0; j < 8; ++j) prev[j] = cur[j];
      hlerp8(q0[k], q1[k], lw0, lw1, cur);
#pragma unroll
      for (int r = 0; r < kUpRows; ++r) {
        if (!(row_in & (1u << r)) || k1[r] != k) continue;
        if (k0[r] == k) *reinterpret_cast<uint4*>(orow + r * ostride) = vlerp8(cur, cur, lh0[r], lh1[r]);
        else *reinterpret_cast<uint4*>(orow + r * ostride) = vlerp8(prev, cur, lh0[r], lh1[r]);
      }
    }
    // rows of the group that fall into the zero padding
#pragma unroll
    for (int r = 0; r < kUpRows; ++r)
      if (ho0 + r < Ho && !(row_in & (1u << r)))
        *reinterpret_cast<uint4*>(orow + r * ostride) = make_uint4(0, 0, 0, 0);
  }
}

// weight with which upsampled row/col `o` reads source index `i`
__device__ __forceinline__ float bilinear_weight(int o, float ratio, int in_size, int i) {
  int i0, i1;
  float l0, l1;
  bilinear_src(o, ratio, in_size, i0, i1, l0, l1);
  return (i0 == i ? l0 : 0.f) + (i1 == i ? l1 : 0.f);
}

// candidate upsampled indices whose 2-tap stencil can touch source index i (conservative)
__device__ __forceinline__ void bilinear_candidates(int i, float ratio, int in_size, int& lo, int& hi) {
  lo = 0;
  hi = 2 * in_size - 1;
  if (ratio > 0.f) {
    lo = max(0, (int)floorf((float)(i - 1) / ratio) - 1);
    hi = min(2 * in_size - 1, (int)ceilf((float)(i + 1) / ratio) + 1);
  }
}

constexpr int kUpMaxTaps = 8;
constexpr int kUpBwdHalo = 12;      // extra output columns staged per chunk (conservative stencil reach)
constexpr int kUpBwdMaxCols = 64;   // source columns per block (size of the column tap table)

__host__ __device__ inline int up_bwd_chunk_cols(int C) {   // source columns per block
  const int c = 4096 / C;
  return c < 8 ? 8 : (c > kUpBwdMaxCols ? kUpBwdMaxCols : c);
}

constexpr int kUpBwdRows = 8;       // source rows per block (they share the column tap table)

// Gather-form backward (deterministic, no atomics), separable in two stages.  One block =
// (image, group of kUpBwdRows source rows, chunk of source columns); per source row:
//   1. vertical:   V[ow][c] = sum over the <= 4 contributing output rows of wh * dout[oh][ow][c]
//      for every output column the chunk can touch, fp32, into shared memory;
//   2. horizontal: dx[wi][c] = sum over the <= 4 contributing output columns of ww * V[ow][c].
// The taps (which output rows / columns touch a source row / column, with which weight) are
// found with the forward's own bilinear_src arithmetic, once per block, one candidate per lane,
// into shared-memory tables (one per source row of the group, one per source column of the chunk).
// History (ncu, 64ch 512->256, batch 64): one stage, ~980 instructions per source vector, 2.2 TB/s;
// two stages with per-thread tap search 1.12 ms; tables per (row, chunk) block 0.85 ms, still
// issue bound (772 M warp instructions: 17 % tap search, 16 % 64-bit address multiplies); this
// version amortises the tables over 8 rows and addresses with 32-bit offsets inside one image.
__global__ void __launch_bounds__(kUpThreads)
upsample2x_pad_bwd_kernel(const __nv_bfloat16* __restrict__ dout, long lddo,
                          __nv_bfloat16* __restrict__ dx, long lddx, int h, int w, int Ho, int Wo,
                          int CG, int pad_top, int pad_left, float rh, float rw, int chunk_cols,
                          int chunks, int row_groups) {
  // V as [col][half][cg] float4: the lanes of a warp store / load consecutive 16-byte words
  extern __shared__ float4 s_v[];   // [2 * chunk_cols + kUpBwdHalo][2][CG]
  __shared__ int s_roff[kUpBwdRows][kUpMaxTaps];     // element offset of the output row in its image
  __shared__ float s_wts[kUpBwdRows][kUpMaxTaps];
  __shared__ int s_cnt[kUpBwdRows];
  __shared__ int s_ccnt[kUpBwdMaxCols];
  __shared__ int s_ccol[kUpBwdMaxCols][kUpMaxTaps];     // staged column index (relative to ow_lo)
  __shared__ float s_cwt[kUpBwdMaxCols][kUpMaxTaps];
  const int chunk = blockIdx.x % chunks;
  const int rg = (blockIdx.x / chunks) % row_groups;
  const int n = blockIdx.x / (chunks * row_groups);
  const int hi0 = rg * kUpBwdRows;
  const int nrows = min(kUpBwdRows, h - hi0);
  const int wi0 = chunk * chunk_cols;
  const int wi1 = min(w, wi0 + chunk_cols);
  // output columns (upsampled coordinates) whose stencil can touch [wi0, wi1)
  int ow_lo, ow_hi, t0, t1;
  bilinear_candidates(wi0, rw, w, ow_lo, t0);
  bilinear_candidates(wi1 - 1, rw, w, t1, ow_hi);
  int ncols = ow_hi - ow_lo + 1;
  if (ncols > 2 * chunk_cols + kUpBwdHalo) ncols = 2 * chunk_cols + kUpBwdHalo;   // never (reach <= 3)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // tap search, one candidate per lane (the candidate window is <= 10 wide, see
  // bilinear_candidates); hits are compacted in candidate order with a ballot.
  // rows: warp r handles source row hi0 + r
  if (warp < nrows) {
    const int hi = hi0 + warp;
    int lo, hi_c, cnt = 0;
    bilinear_candidates(hi, rh, h, lo, hi_c);
    for (int b0 = lo; b0 <= hi_c; b0 += 32) {
      const int oh = b0 + lane;
      const float wh = oh <= hi_c ? bilinear_weight(oh, rh, h, hi) : 0.f;
      const int ph = oh + pad_top;
      const bool ok = wh != 0.f && ph >= 0 && ph < Ho;
      const unsigned m = __ballot_sync(0xffffffffu, ok);
      const int pos = cnt + __popc(m & ((1u << lane) - 1u));
      if (ok && pos < kUpMaxTaps) { s_roff[warp][pos] = ph * Wo * (int)lddo; s_wts[warp][pos] = wh; }
      cnt = min(kUpMaxTaps, cnt + __popc(m));
    }
    __syncwarp();
    // pad to a multiple of four taps (weight 0 on a row that is read anyway) so that the gather
    // below always has four independent loads in flight
    if (lane >= cnt && lane < kUpMaxTaps) {
      s_roff[warp][lane] = cnt > 0 ? s_roff[warp][0] : 0;
      s_wts[warp][lane] = 0.f;
    }
    if (lane == 0) s_cnt[warp] = cnt;
  }
  // columns: 8 lanes per source column of the chunk
  {
    const int grp = threadIdx.x >> 3, l8 = lane & 7;
    constexpr int kGroups = kUpThreads / 8;
    for (int wb = 0; wb < wi1 - wi0; wb += kGroups) {        // block-uniform trip count
      const int wl = wb + grp;
      const bool col_ok = wl < wi1 - wi0;
      int c_lo = 0, c_hi = -1, cnt = 0;
      if (col_ok) bilinear_candidates(wi0 + wl, rw, w, c_lo, c_hi);
#pragma unroll
      for (int it = 0; it < 2; ++it) {                        // window <= 16 candidates
        const int ow = c_lo + it * 8 + l8;
        const float ww = (col_ok && ow <= c_hi) ? bilinear_weight(ow, rw, w, wi0 + wl) : 0.f;
        const int col = ow - ow_lo;
        const bool ok = ww != 0.f && col >= 0 && col < ncols;
        const unsigned m = (__ballot_sync(0xffffffffu, ok) >> (lane & 24)) & 0xffu;
        const int pos = cnt + __popc(m & ((1u << l8) - 1u));
        if (ok && pos < kUpMaxTaps) { s_ccol[wl][pos] = col; s_cwt[wl][pos] = ww; }
        cnt = min(kUpMaxTaps, cnt + __popc(m));
      }
      if (col_ok && l8 == 0) s_ccnt[wl] = cnt;
    }
  }
  __syncthreads();
  const int cols_per_pass = kUpThreads / CG;
  const int cg = threadIdx.x % CG;
  const int col0 = threadIdx.x / CG;
  const bool active = col0 < cols_per_pass;
  const __nv_bfloat16* img = dout + (long)n * Ho * Wo * lddo + cg * 8;
  const int ildo = (int)lddo;
  for (int r = 0; r < nrows; ++r) {
    const int cnt_h = s_cnt[r];
    if (active) {
      for (int col = col0; col < ncols; col += cols_per_pass) {
        const int pw = ow_lo + col + pad_left;
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        if (pw >= 0 && pw < Wo) {
          const __nv_bfloat16* src = img + pw * ildo;
          for (int t0 = 0; t0 < cnt_h; t0 += 4) {
            uint4 q[4];
#pragma unroll
            // plain read-only loads: each output row is gathered by two source rows and the second
            // read must hit L2 (ld_stream's L1::no_allocate marks the line evict-first in L2:
            // measured 4.17 GB of DRAM reads for 2.15 GB of gradient)
            for (int t = 0; t < 4; ++t) q[t] = __ldg(reinterpret_cast<const uint4*>(src + s_roff[r][t0 + t]));
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              float g[8];
              unpack8(q[t], g);
              const float wt = s_wts[r][t0 + t];
#pragma unroll
              for (int j = 0; j < 8; ++j) acc[j] = fmaf(wt, g[j], acc[j]);
            }
          }
        }
        s_v[(col * 2 + 0) * CG + cg] = make_float4(acc[0], acc[1], acc[2], acc[3]);
        s_v[(col * 2 + 1) * CG + cg] = make_float4(acc[4], acc[5], acc[6], acc[7]);
      }
    }
    __syncthreads();
    if (active) {
      __nv_bfloat16* drow = dx + ((long)n * h + hi0 + r) * w * lddx + cg * 8;
      for (int wl = col0; wl < wi1 - wi0; wl += cols_per_pass) {
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        const int cnt = s_ccnt[wl];
        for (int t = 0; t < cnt; ++t) {
          const float ww = s_cwt[wl][t];
          const int col = s_ccol[wl][t];
          const float4 a = s_v[(col * 2 + 0) * CG + cg], b = s_v[(col * 2 + 1) * CG + cg];
          acc[0] = fmaf(ww, a.x, acc[0]); acc[1] = fmaf(ww, a.y, acc[1]);
          acc[2] = fmaf(ww, a.z, acc[2]); acc[3] = fmaf(ww, a.w, acc[3]);
          acc[4] = fmaf(ww, b.x, acc[4]); acc[5] = fmaf(ww, b.y, acc[5]);
          acc[6] = fmaf(ww, b.z, acc[6]); acc[7] = fmaf(ww, b.w, acc[7]);
        }
        *reinterpret_cast<uint4*>(drow + (long)(wi0 + wl) * lddx) = pack8(acc);
      }
    }
    __syncthreads();   // V is overwritten by the next row
  }
}

// ---------------------------------------------------------------------------
// Adam
// ---------------------------------------------------------------------------
// step_dev (nullable): device-resident step counter, so the launch is CUDA-graph replayable;
// the LAST block to finish increments it (threadfence + atomic ticket).
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                            float* __restrict__ m, float* __restrict__ v, long n, float lr,
                            float beta1, float beta2, float eps, float bc1, float bc2_sqrt,
                            float grad_scale, int* step_dev, unsigned int* ticket) {
  if (step_dev != nullptr) {
    const float t = (float)(*step_dev + 1);
    bc1 = 1.f - powf(beta1, t);
    bc2_sqrt = sqrtf(1.f - powf(beta2, t));
  }
  const float step_size = lr / bc1;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n;
       i += (long)gridDim.x * blockDim.x) {
    const float gi = g[i] * grad_scale;
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] -= step_size * (mi / denom);
  }
  if (step_dev != nullptr) {
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      if (atomicAdd(ticket, 1u) == gridDim.x - 1) {
        *ticket = 0u;
        *step_dev += 1;
      }
    }
  }
}

}  // namespace fp

using namespace fp;

extern "C" {

int fpb200_abi_version(void) { return 1; }

int fpb200_ingest_nchw_f32_to_nhwc_bf16(const float* const* srcs, const int* src_channels,
                                        int n_src, void* dst, int c_pad, int N, int H, int W,
                                        void* stream) {
  if (n_src < 1 || n_src > 8 || c_pad % 8 != 0 || c_pad > kIngestMaxC) return FPB200_ERR_SHAPE;
  for (int s = 0; s < n_src; ++s)
    if (src_channels[s] < 1 || src_channels[s] > 255) return FPB200_ERR_SHAPE;   // 255 is the zero-pad marker
  IngestArgs a;
  int c = 0;
  for (int s = 0; s < 8; ++s) { a.src[s] = nullptr; a.src_c[s] = 0; }
  for (int s = 0; s < n_src; ++s) {
    a.src[s] = srcs[s];
    a.src_c[s] = src_channels[s];
    for (int k = 0; k < src_channels[s]; ++k) {
      if (c >= c_pad) return FPB200_ERR_SHAPE;
      a.ch_src[c] = (unsigned char)s;
      a.ch_idx[c] = (unsigned char)k;
      ++c;
    }
  }
  for (; c < kIngestMaxC; ++c) { a.ch_src[c] = 0; a.ch_idx[c] = 255; }
  const long total = (long)N * H * W;
  ingest_kernel<<<grid_for(total, 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(
      a, (__nv_bfloat16*)dst, c_pad, N, (long)H * W);
  return check_launch("ingest");
}

int fpb200_ingest_scene_tiles(const float* scene, int C, long H, long W, const int* tiles,
                              int n_tiles, int th, int tw, void* dst, int c_pad, void* stream) {
  if (c_pad % 8 != 0 || c_pad > kIngestMaxC || C > c_pad || n_tiles < 1) return FPB200_ERR_SHAPE;
  const long total = (long)n_tiles * th * tw;
  ingest_scene_tiles_kernel<<<grid_for(total, 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(
      scene, C, H, W, tiles, n_tiles, th, tw, (__nv_bfloat16*)dst, c_pad);
  return check_launch("ingest_scene_tiles");
}

int fpb200_repack_weights_fprop(const float* w_oihw, void* w_packed, int Cout, int Cin,
                                int cin_pad, void* stream) {
  if (cin_pad < Cin) return FPB200_ERR_SHAPE;
  const long total = (long)Cout * 9 * cin_pad;
  repack_fprop_kernel<<<grid_for(total, 256, 148 * 8), 256, 0, (cudaStream_t)stream>>>(
      w_oihw, (__nv_bfloat16*)w_packed, Cout, Cin, cin_pad);
  return check_launch("repack_fprop");
}

int fpb200_repack_weights_dgrad(const float* w_oihw, void* w_packed, int Cout, int Cin,
                                void* stream) {
  const long total = (long)Cout * 9 * Cin;
  repack_dgrad_kernel<<<grid_for(total, 256, 148 * 8), 256, 0, (cudaStream_t)stream>>>(
      w_oihw, (__nv_bfloat16*)w_packed, Cout, Cin);
  return check_launch("repack_dgrad");
}

int fpb200_repack_weights_batch(const void* table, int n_entries, long total_blocks, void* stream) {
  if (n_entries < 1 || total_blocks < 1 || total_blocks > 0x7fffffffL) return FPB200_ERR_SHAPE;
  static_assert(sizeof(RepackEntry) == 40, "RepackEntry is mirrored field by field on the host");
  repack_batch_kernel<<<(int)total_blocks, 256, 0, (cudaStream_t)stream>>>(
      (const RepackEntry*)table, n_entries);
  return check_launch("repack_weights_batch");
}

int fpb200_bn_stats_finalize(const float* partials, int num_partials, int C, double count,
                             const float* gamma, const float* beta, const float* conv_bias,
                             float eps, float momentum, float* running_mean, float* running_var,
                             float* scale, float* shift, float* save_mean, float* save_invstd,
                             void* stream) {
  if (C <= 0 || num_partials <= 0) return FPB200_ERR_SHAPE;
  bn_stats_finalize_kernel<<<(C + 31) / 32, kFinThreads, 0, (cudaStream_t)stream>>>(
      partials, num_partials, C, count, gamma, beta, conv_bias, eps, momentum, running_mean,
      running_var, scale, shift, save_mean, save_invstd);
  return check_launch("bn_stats_finalize");
}

int fpb200_bn_fold_eval(const float* gamma, const float* beta, const float* conv_bias,
                        const float* running_mean, const float* running_var, float eps, int C,
                        float* scale, float* shift, void* stream) {
  bn_fold_eval_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      gamma, beta, conv_bias, running_mean, running_var, eps, C, scale, shift);
  return check_launch("bn_fold_eval");
}

int fpb200_bn_apply_relu(const void* y, long ldy, void* a, long lda, const float* scale,
                         const float* shift, long num_pixels, int C, void* stream) {
  if (C % 8 != 0 || ldy % 8 != 0 || lda % 8 != 0) return FPB200_ERR_SHAPE;
  const long total = ((num_pixels + kPixPerThread - 1) / kPixPerThread) * (C / 8);
  bn_apply_relu_kernel<<<grid_for(total, 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)y, ldy, (__nv_bfloat16*)a, lda, scale, shift, num_pixels, C / 8);
  return check_launch("bn_apply_relu");
}

int fpb200_bn_apply_relu_maxpool2(const void* y, long ldy, void* a, long lda, void* pooled,
                                  long ldp, uint8_t* pool_idx, const float* scale,
                                  const float* shift, int N, int H, int W, int C, void* stream) {
  if (C % 8 != 0 || ldy % 8 != 0 || ldp % 8 != 0 || (a && lda % 8 != 0)) return FPB200_ERR_SHAPE;
  bn_apply_relu_maxpool2_kernel<<<N * ((H + 1) / 2), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)y, ldy, (__nv_bfloat16*)a, lda, (__nv_bfloat16*)pooled, ldp, pool_idx,
      scale, shift, H, W, C / 8);
  return check_launch("bn_apply_relu_maxpool2");
}

int fpb200_maxpool2_bwd(const void* dpooled, long lddp, const uint8_t* pool_idx, const void* dskip,
                        long ldds, void* dx, long lddx, int N, int H, int W, int C, const void* bn_y,
                        long ld_bn_y, const float* bn_scale, const float* bn_shift,
                        const float* bn_mean, const float* bn_invstd, float* bn_partials,
                        void* stream) {
  if (C % 8 != 0 || lddp % 8 != 0 || lddx % 8 != 0) return FPB200_ERR_SHAPE;
  const int rows = N * ((H + 1) / 2);
  if (bn_y == nullptr) {
    maxpool2_bwd_kernel<false><<<rows, 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)dpooled, lddp, pool_idx, (const __nv_bfloat16*)dskip, ldds,
        (__nv_bfloat16*)dx, lddx, N, H, W, C / 8, nullptr, 0, nullptr, nullptr, nullptr, nullptr, nullptr);
    return check_launch("maxpool2_bwd");
  }
  if (256 % (C / 8) != 0 || ld_bn_y % 8 != 0 || !bn_scale || !bn_shift || !bn_mean || !bn_invstd ||
      !bn_partials)
    return FPB200_ERR_SHAPE;
  const int prow = fpb200_bn_bwd_rows() * 2;
  const int grid = rows < prow ? rows : prow;
  if (grid < prow &&
      cudaMemsetAsync(bn_partials + (size_t)grid * 2 * C, 0, (size_t)(prow - grid) * 2 * C * sizeof(float),
                      (cudaStream_t)stream) != cudaSuccess)
    return check_launch("maxpool2_bwd memset");
  maxpool2_bwd_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)dpooled, lddp, pool_idx, (const __nv_bfloat16*)dskip, ldds,
      (__nv_bfloat16*)dx, lddx, N, H, W, C / 8, (const __nv_bfloat16*)bn_y, ld_bn_y, bn_scale, bn_shift,
      bn_mean, bn_invstd, bn_partials);
  return check_launch("maxpool2_bwd_fused");
}

int fpb200_bn_bwd_rows(void) { return 4 * sm_count(); }

int fpb200_bn_relu_bwd_reduce(const void* da, long ldda, const void* y, long ldy,
                              const float* scale, const float* shift, const float* save_mean,
                              const float* save_invstd, float* partials, long num_pixels, int C,
                              void* stream) {
  if (C % 64 != 0 || C > 2048 || (256 % (C / 8)) != 0) return FPB200_ERR_SHAPE;
  bn_relu_bwd_reduce_kernel<<<fpb200_bn_bwd_rows(), kBnBwdThreads, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)da, ldda, (const __nv_bfloat16*)y, ldy, scale, shift, save_mean,
      save_invstd, partials, num_pixels, C);
  return check_launch("bn_relu_bwd_reduce");
}

int fpb200_bn_bwd_finalize(const float* partials, int num_partials, int C, double count,
                           const float* scale, const float* save_mean, const float* save_invstd,
                           float* dgamma, float* dbeta, float* coef, void* stream) {
  bn_bwd_finalize_kernel<<<(C + 31) / 32, kFinThreads, 0, (cudaStream_t)stream>>>(
      partials, num_partials, C, count, scale, save_mean, save_invstd, dgamma, dbeta, coef);
  return check_launch("bn_bwd_finalize");
}

int fpb200_bn_eval_stats(const float* conv_bias, const float* running_mean, const float* running_var,
                         float eps, int C, float* mean_eff, float* invstd, void* stream) {
  if (C < 1) return FPB200_ERR_SHAPE;
  bn_eval_stats_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(conv_bias, running_mean, running_var,
                                                                         eps, C, mean_eff, invstd);
  return check_launch("bn_eval_stats");
}

int fpb200_bn_bwd_finalize_frozen(const float* partials, int num_partials, int C, const float* scale,
                                  float* dgamma, float* dbeta, float* dbias, float* coef, void* stream) {
  if (C < 1 || num_partials < 1) return FPB200_ERR_SHAPE;
  bn_bwd_finalize_frozen_kernel<<<(C + 31) / 32, kFinThreads, 0, (cudaStream_t)stream>>>(partials, num_partials, C, scale,
                                                                                  dgamma, dbeta, dbias, coef);
  return check_launch("bn_bwd_finalize_frozen");
}

int fpb200_bn_relu_bwd_apply(const void* da, long ldda, const void* y, long ldy, void* dy,
                             long lddy, const float* scale, const float* shift, const float* coef,
                             long num_pixels, int C, void* stream) {
  if (C % 8 != 0) return FPB200_ERR_SHAPE;
  const long total = ((num_pixels + kPixPerThread - 1) / kPixPerThread) * (C / 8);
  bn_relu_bwd_apply_kernel<<<grid_for(total, 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)da, ldda, (const __nv_bfloat16*)y, ldy, (__nv_bfloat16*)dy, lddy, scale,
      shift, coef, num_pixels, C / 8);
  return check_launch("bn_relu_bwd_apply");
}

static inline float align_corners_ratio(int in_size, int out_size) {
  return out_size > 1 ? (float)(in_size - 1) / (float)(out_size - 1) : 0.f;
}

int fpb200_upsample2x_pad_concat_fwd(const void* x, long ldx, void* out, long ldo, int N, int h,
                                     int w, int Ho, int Wo, int C, void* stream) {
  if (C % 8 != 0 || C / 8 > kUpThreads || Ho < 2 * h || Wo < 2 * w) return FPB200_ERR_SHAPE;
  const int pad_top = (Ho - 2 * h) / 2, pad_left = (Wo - 2 * w) / 2;
  upsample2x_pad_fwd_kernel<<<N * ((Ho + kUpRows - 1) / kUpRows), kUpThreads, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)x, ldx, (__nv_bfloat16*)out, ldo, h, w, Ho, Wo, C / 8, pad_top,
      pad_left, align_corners_ratio(h, 2 * h), align_corners_ratio(w, 2 * w));
  return check_launch("upsample2x_pad_concat_fwd");
}

int fpb200_upsample2x_pad_concat_bwd(const void* dout, long lddo, void* dx, long lddx, int N,
                                     int h, int w, int Ho, int Wo, int C, void* stream) {
  if (C % 8 != 0 || C / 8 > kUpThreads || Ho < 2 * h || Wo < 2 * w) return FPB200_ERR_SHAPE;
  const int pad_top = (Ho - 2 * h) / 2, pad_left = (Wo - 2 * w) / 2;
  const int chunk_cols = up_bwd_chunk_cols(C);
  const int chunks = (w + chunk_cols - 1) / chunk_cols;
  const int smem = (2 * chunk_cols + kUpBwdHalo) * C * (int)sizeof(float);
  static int smem_set[kMaxDevices] = {0};   // per device, like the attribute itself
  const int dev_ = current_device();
  if (smem > smem_set[dev_]) {
    if (cudaFuncSetAttribute(upsample2x_pad_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             64 * 1024) != cudaSuccess)
      return check_launch("upsample2x_pad_concat_bwd smem attribute");
    smem_set[dev_] = 64 * 1024;
  }
  // 32-bit element offsets inside one image of dout
  if (smem > 64 * 1024 || (long)Ho * Wo * lddo >= (1L << 31)) return FPB200_ERR_SHAPE;
  const int row_groups = (h + kUpBwdRows - 1) / kUpBwdRows;
  upsample2x_pad_bwd_kernel<<<N * row_groups * chunks, kUpThreads, smem, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)dout, lddo, (__nv_bfloat16*)dx, lddx, h, w, Ho, Wo, C / 8, pad_top,
      pad_left, align_corners_ratio(h, 2 * h), align_corners_ratio(w, 2 * w), chunk_cols, chunks,
      row_groups);
  return check_launch("upsample2x_pad_concat_bwd");
}

int fpb200_adam_step(float* p, const float* g, float* m, float* v, long n, float lr, float beta1,
                     float beta2, float eps, int step, float grad_scale, void* stream) {
  if (n <= 0 || step < 1) return FPB200_ERR_SHAPE;
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2 = 1.f - powf(beta2, (float)step);
  adam_kernel<<<grid_for(n, 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(
      p, g, m, v, n, lr, beta1, beta2, eps, bc1, sqrtf(bc2), grad_scale, nullptr, nullptr);
  return check_launch("adam_step");
}

int fpb200_adam_step_graphable(float* p, const float* g, float* m, float* v, long n, float lr,
                               float beta1, float beta2, float eps, int* step_state,
                               float grad_scale, void* stream) {
  if (n <= 0 || step_state == nullptr) return FPB200_ERR_SHAPE;
  adam_kernel<<<grid_for(n, 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(
      p, g, m, v, n, lr, beta1, beta2, eps, 1.f, 1.f, grad_scale, step_state,
      reinterpret_cast<unsigned int*>(step_state + 1));
  return check_launch("adam_step_graphable");
}

}  // extern "C"
