// Bandwidth kernels around the feature-level seams of the network:
//
//   * NCHW fp32 <-> NHWC bf16 feature-map transposes (any channel count that is a multiple of
//     8): the layout/cast step at the UNet.encode / UNet.decode API boundary of the reference
//     (st_water_seg/models/unet.py:113-131, UNetEncoder/UNetDecoder :134-191), whose callers
//     exchange fp32 NCHW feature lists.  Tiled through shared memory so both sides move
//     128-byte rows.
//   * per-channel pixel sums of an NHWC bf16 view: the bias gradient of the late-fusion
//     `concat_convs` (nn.Conv2d(fs*k, fs, 1, 1), st_water_seg/models/lf_model.py:40-45),
//     two-stage and deterministic like every other reduction of the path.
//   * 1x1 weight packing (fp32 [Cout][Cin] -> bf16 GEMM operand, optionally transposed for
//     the data gradient).
#include "host_common.h"
#include "ptx.cuh"

namespace fp {

constexpr int kTrC = 64;   // channels per tile
constexpr int kTrP = 32;   // pixels per tile

// src [N][C][HW] fp32  ->  dst view [N][HW][ld] bf16 (channels [0, C))
__global__ void __launch_bounds__(256)
nchw_f32_to_nhwc_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long ld,
                             int C, long HW, long tiles_p, int tiles_c) {
  __shared__ float tile[kTrC][kTrP + 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.y;
  for (long t = blockIdx.x; t < tiles_p * tiles_c; t += gridDim.x) {
    const int tc = (int)(t % tiles_c);
    const long tp = t / tiles_c;
    const long p0 = tp * kTrP;
    const int c0 = tc * kTrC;
    // load: warp w reads channels w, w+8, ...; lane = pixel (128-byte rows)
#pragma unroll
    for (int j = 0; j < kTrC / 8; ++j) {
      const int c = c0 + warp + 8 * j;
      const long p = p0 + lane;
      tile[warp + 8 * j][lane] = (c < C && p < HW) ? __ldg(src + ((long)n * C + c) * HW + p) : 0.f;
    }
    __syncthreads();
    // store: warp w writes pixels w, w+8, ...; lane = channel pair (128-byte rows)
#pragma unroll
    for (int j = 0; j < kTrP / 8; ++j) {
      const int pl = warp + 8 * j;
      const long p = p0 + pl;
      const int c = c0 + 2 * lane;
      if (p < HW && c < C) {
        const uint32_t v = pack_bf16x2(tile[2 * lane][pl], tile[2 * lane + 1][pl]);
        *reinterpret_cast<uint32_t*>(dst + ((long)n * HW + p) * ld + c) = v;
      }
    }
    __syncthreads();
  }
}

// src view [N][HW][ld] bf16  ->  dst [N][C][HW] fp32
__global__ void __launch_bounds__(256)
nhwc_bf16_to_nchw_f32_kernel(const __nv_bfloat16* __restrict__ src, long ld, float* __restrict__ dst,
                             int C, long HW, long tiles_p, int tiles_c) {
  __shared__ float tile[kTrC][kTrP + 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.y;
  for (long t = blockIdx.x; t < tiles_p * tiles_c; t += gridDim.x) {
    const int tc = (int)(t % tiles_c);
    const long tp = t / tiles_c;
    const long p0 = tp * kTrP;
    const int c0 = tc * kTrC;
#pragma unroll
    for (int j = 0; j < kTrP / 8; ++j) {
      const int pl = warp + 8 * j;
      const long p = p0 + pl;
      const int c = c0 + 2 * lane;
      uint32_t v = 0;
      if (p < HW && c < C) v = __ldg(reinterpret_cast<const uint32_t*>(src + ((long)n * HW + p) * ld + c));
      tile[2 * lane][pl] = bf16_lo(v);
      tile[2 * lane + 1][pl] = bf16_hi(v);
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kTrC / 8; ++j) {
      const int c = c0 + warp + 8 * j;
      const long p = p0 + lane;
      if (c < C && p < HW) dst[((long)n * C + c) * HW + p] = tile[warp + 8 * j][lane];
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------
// per-channel sum over pixels (bias gradient of a pointwise convolution)
// ---------------------------------------------------------------------------
constexpr int kCsThreads = 256;

__device__ __forceinline__ uint4 ld_stream16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

__global__ void __launch_bounds__(kCsThreads)
channel_sum_kernel(const __nv_bfloat16* __restrict__ x, long ld, float* __restrict__ partials,
                   long num_pixels, int C) {
  __shared__ float red[kCsThreads * 8];
  const int CG = C >> 3;              // 8-channel groups; the launcher guarantees CG divides 256
  const int TP = kCsThreads / CG;     // pixel lanes per block
  const int cg = threadIdx.x % CG;
  const int pl = threadIdx.x / CG;
  float s[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = 0.f;
  for (long px = (long)blockIdx.x * TP + pl; px < num_pixels; px += (long)gridDim.x * TP) {
    const uint4 u = ld_stream16(x + px * ld + cg * 8);
    s[0] += bf16_lo(u.x); s[1] += bf16_hi(u.x);
    s[2] += bf16_lo(u.y); s[3] += bf16_hi(u.y);
    s[4] += bf16_lo(u.z); s[5] += bf16_hi(u.z);
    s[6] += bf16_lo(u.w); s[7] += bf16_hi(u.w);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[threadIdx.x * 8 + j] = s[j];
  __syncthreads();
  for (int o = threadIdx.x; o < CG * 8; o += kCsThreads) {
    const int ocg = o >> 3, oj = o & 7;
    float acc = 0.f;
    for (int p = 0; p < TP; ++p) acc += red[(p * CG + ocg) * 8 + oj];
    partials[(size_t)blockIdx.x * C + ocg * 8 + oj] = acc;
  }
}

__global__ void channel_sum_finalize_kernel(const float* __restrict__ partials, int P, int C,
                                            float* __restrict__ out) {
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (c >= C) return;
  double s = 0.0;
  for (int p = lane; p < P; p += 32) s += (double)partials[(size_t)p * C + c];
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[c] = (float)s;
}

// out[r][c] = w[r][c] (transpose == 0, [Cout][Cin]) or w[c][r] (transpose != 0, [Cin][Cout])
__global__ void repack_1x1_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out,
                                  int Cout, int Cin, int transpose) {
  const long total = (long)Cout * Cin;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total;
       i += (long)gridDim.x * blockDim.x) {
    float v;
    if (!transpose) {
      v = w[i];
    } else {
      const int co = i % Cout;
      const int ci = i / Cout;
      v = w[(long)co * Cin + ci];
    }
    out[i] = __float2bfloat16(v);
  }
}

}  // namespace fp

using namespace fp;

extern "C" {

int fpb200_nchw_f32_to_nhwc_bf16(const float* src, void* dst, long ld, int N, int C, int H, int W,
                                 void* stream) {
  if (N <= 0 || C <= 0 || H <= 0 || W <= 0 || C % 8 != 0 || ld % 8 != 0 || ld < C)
    return FPB200_ERR_SHAPE;
  if ((reinterpret_cast<uintptr_t>(dst) & 15) || (reinterpret_cast<uintptr_t>(src) & 3))
    return FPB200_ERR_ALIGN;
  const long HW = (long)H * W;
  const long tiles_p = (HW + kTrP - 1) / kTrP;
  const int tiles_c = (C + kTrC - 1) / kTrC;
  long gx = tiles_p * tiles_c;
  if (gx > 148L * 16) gx = 148L * 16;
  dim3 grid((unsigned)gx, (unsigned)N);
  nchw_f32_to_nhwc_bf16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
      src, reinterpret_cast<__nv_bfloat16*>(dst), ld, C, HW, tiles_p, tiles_c);
  return check_launch("nchw_f32_to_nhwc_bf16");
}

int fpb200_nhwc_bf16_to_nchw_f32(const void* src, long ld, float* dst, int N, int C, int H, int W,
                                 void* stream) {
  if (N <= 0 || C <= 0 || H <= 0 || W <= 0 || C % 8 != 0 || ld % 8 != 0 || ld < C)
    return FPB200_ERR_SHAPE;
  if ((reinterpret_cast<uintptr_t>(src) & 15) || (reinterpret_cast<uintptr_t>(dst) & 3))
    return FPB200_ERR_ALIGN;
  const long HW = (long)H * W;
  const long tiles_p = (HW + kTrP - 1) / kTrP;
  const int tiles_c = (C + kTrC - 1) / kTrC;
  long gx = tiles_p * tiles_c;
  if (gx > 148L * 16) gx = 148L * 16;
  dim3 grid((unsigned)gx, (unsigned)N);
  nhwc_bf16_to_nchw_f32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(src), ld, dst, C, HW, tiles_p, tiles_c);
  return check_launch("nhwc_bf16_to_nchw_f32");
}

int fpb200_channel_sum_rows(void) { return 4 * sm_count(); }

int fpb200_channel_sum_bf16_nhwc(const void* x, long ld, float* partials, float* out,
                                 long num_pixels, int C, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (num_pixels <= 0 || C < 8 || C % 8 != 0 || ld % 8 != 0 || ld < C) return FPB200_ERR_SHAPE;
  // the block maps 256 threads onto C/8 channel groups x pixel lanes
  if ((C >> 3) > kCsThreads || kCsThreads % (C >> 3) != 0) return FPB200_ERR_SHAPE;
  if (reinterpret_cast<uintptr_t>(x) & 15) return FPB200_ERR_ALIGN;
  const int rows = fpb200_channel_sum_rows();
  channel_sum_kernel<<<rows, kCsThreads, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), ld,
                                                      partials, num_pixels, C);
  int rc = check_launch("channel_sum");
  if (rc != FPB200_OK) return rc;
  channel_sum_finalize_kernel<<<(C + 7) / 8, 256, 0, stream>>>(partials, rows, C, out);
  return check_launch("channel_sum_finalize");
}

int fpb200_repack_weights_1x1(const float* w_oi, void* w_packed, int Cout, int Cin, int transpose,
                              void* stream) {
  if (Cout <= 0 || Cin <= 0) return FPB200_ERR_SHAPE;
  const long total = (long)Cout * Cin;
  long g = (total + 255) / 256;
  if (g > 148L * 8) g = 148L * 8;
  repack_1x1_kernel<<<(int)g, 256, 0, (cudaStream_t)stream>>>(
      w_oi, reinterpret_cast<__nv_bfloat16*>(w_packed), Cout, Cin, transpose);
  return check_launch("repack_weights_1x1");
}

}  // extern "C"
