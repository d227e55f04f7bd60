// Per-sample normalise + augment on the device: what the reference does in its
// DataLoader workers, sample by sample in numpy / torchvision
//   BaseDataset.normalize          st_water_seg/datasets/base_dataset.py:77-113
//   BaseDataset.apply_transforms   st_water_seg/datasets/base_dataset.py:532-555
//   (hflip -> vflip -> rotate, nearest / fill 0; st_water_seg/datasets/floodplanet.py:616-640)
// is ONE gather pass over the batch here: every output pixel computes the source pixel it reads
// through the composed transform, normalises it and writes the fp32 NCHW image the reference's
// DataLoader would have produced and / or the NHWC bf16 tensor the first convolution consumes
// (fused ingest), plus the int64 annotation.  HBM bound: C*4 B read + C*4 B (or c_pad*2 B) written
// per pixel, 8 + 8 B for the annotation.
//
// The rotation reproduces torchvision's CPU arithmetic bit for bit (oracle/augment_oracle.py):
//   grid  = fma(Y, t_y, X * t_x) + t_0            (fp32; X, Y = the linspace base grid)
//   index = nearbyint((grid + 1) * (size / 2) - 0.5), zero outside the image
// with explicit round-to-nearest intrinsics so that the compiler cannot contract differently.
#include "host_common.h"
#include "ptx.cuh"

namespace fp {

constexpr int kAugFlagH = 1, kAugFlagV = 2, kAugFlagR = 4;

__device__ __forceinline__ bool aug_source(int flags, const float* __restrict__ t, float X, float Y,
                                           int H, int W, int y, int x, int& sy, int& sx) {
  sy = y;
  sx = x;
  bool ok = true;
  if (flags & kAugFlagR) {
    const float gx = __fadd_rn(__fmaf_rn(Y, t[2], __fmul_rn(X, t[0])), t[4]);
    const float gy = __fadd_rn(__fmaf_rn(Y, t[3], __fmul_rn(X, t[1])), t[5]);
    const float ix = __fsub_rn(__fmul_rn(__fadd_rn(gx, 1.f), 0.5f * (float)W), 0.5f);
    const float iy = __fsub_rn(__fmul_rn(__fadd_rn(gy, 1.f), 0.5f * (float)H), 0.5f);
    const float xn = rintf(ix), yn = rintf(iy);
    ok = xn >= 0.f && xn <= (float)(W - 1) && yn >= 0.f && yn <= (float)(H - 1);
    sx = ok ? (int)xn : 0;
    sy = ok ? (int)yn : 0;
  }
  // the rotation samples the vertically flipped image, which samples the horizontally flipped one
  if (flags & kAugFlagV) sy = H - 1 - sy;
  if (flags & kAugFlagH) sx = W - 1 - sx;
  return ok;
}

// numpy `image -= mean; image /= std` with float64 (global) or float32 (local) statistics: each
// in-place op is evaluated in the wider type and rounded to the fp32 image; for fp32 statistics the
// double evaluation + rounding equals the fp32 operation (53 >= 2*24 + 2 bits).
__device__ __forceinline__ float aug_normalise(float v, const double* __restrict__ mean,
                                               const double* __restrict__ stdv, long nc) {
  if (mean == nullptr) return v;
  const float d = (float)((double)v - mean[nc]);
  return (float)((double)d / stdv[nc]);
}

// One block = one 32 x 32 output tile of one sample, 32 x 8 threads, 4 rows per thread: under a
// rotation the tile reads a rotated square of the source, so the 32-byte sectors fetched for one
// thread's pixel serve its neighbours through L1 (a row-major 1-D mapping walks a slanted line and
// uses 4 bytes of every sector: measured 1.0 TB/s), and every thread has 4 x (C + 1) independent
// gathers in flight.  Stores are 128-byte rows per warp.
constexpr int kAugTile = 32, kAugRows = 8, kAugPerThread = kAugTile / kAugRows;

__global__ void __launch_bounds__(kAugTile * kAugRows)
augment_kernel(const float* __restrict__ img, float* __restrict__ out_f32,
               __nv_bfloat16* __restrict__ out_bf16, int c_pad, const int64_t* __restrict__ tgt,
               int64_t* __restrict__ tgt_out, const float* __restrict__ theta,
               const int* __restrict__ flags, const float* __restrict__ xgrid,
               const float* __restrict__ ygrid, const double* __restrict__ mean,
               const double* __restrict__ stdv, int N, int C, int H, int W, int tiles_x,
               int tiles_y) {
  const long hw = (long)H * W;
  const int tile = blockIdx.x % (tiles_x * tiles_y);
  const int n = blockIdx.x / (tiles_x * tiles_y);
  const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
  const int x = tx * kAugTile + (threadIdx.x & (kAugTile - 1));
  const int y0 = ty * kAugTile + (threadIdx.x >> 5);
  if (x >= W) return;
  const int fl = flags[n];
  const float* th = theta + (long)n * 6;
  const float t[6] = {th[0], th[1], th[2], th[3], th[4], th[5]};
  const float X = xgrid[x];
  long so[kAugPerThread];
  bool ok[kAugPerThread], in[kAugPerThread];
#pragma unroll
  for (int k = 0; k < kAugPerThread; ++k) {
    const int y = y0 + k * kAugRows;
    in[k] = y < H;
    int sy, sx;
    ok[k] = aug_source(fl, t, X, ygrid[in[k] ? y : 0], H, W, y, x, sy, sx) && in[k];
    so[k] = ok[k] ? (long)sy * W + sx : 0;
  }
  if (tgt_out != nullptr) {
    int64_t v[kAugPerThread];
#pragma unroll
    for (int k = 0; k < kAugPerThread; ++k) v[k] = ok[k] ? __ldg(tgt + (long)n * hw + so[k]) : 0;
#pragma unroll
    for (int k = 0; k < kAugPerThread; ++k)
      if (in[k]) tgt_out[(long)n * hw + (long)(y0 + k * kAugRows) * W + x] = v[k];
  }
  if (out_f32 == nullptr && out_bf16 == nullptr) return;
  const int c_end = out_bf16 != nullptr ? c_pad : C;
  for (int g = 0; g < c_end; g += 8) {
    float f[kAugPerThread][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = g + j;
#pragma unroll
      for (int k = 0; k < kAugPerThread; ++k)
        f[k][j] = (c < C && ok[k]) ? __ldg(img + ((long)n * C + c) * hw + so[k]) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = g + j;
      if (c >= C) continue;
#pragma unroll
      for (int k = 0; k < kAugPerThread; ++k) {
        if (ok[k]) f[k][j] = aug_normalise(f[k][j], mean, stdv, (long)n * C + c);
        if (out_f32 != nullptr && in[k])
          out_f32[((long)n * C + c) * hw + (long)(y0 + k * kAugRows) * W + x] = f[k][j];
      }
    }
    if (out_bf16 != nullptr) {
#pragma unroll
      for (int k = 0; k < kAugPerThread; ++k) {
        if (!in[k]) continue;
        uint4 pk;
        pk.x = pack_bf16x2(f[k][0], f[k][1]);
        pk.y = pack_bf16x2(f[k][2], f[k][3]);
        pk.z = pack_bf16x2(f[k][4], f[k][5]);
        pk.w = pack_bf16x2(f[k][6], f[k][7]);
        *reinterpret_cast<uint4*>(out_bf16 + ((long)n * hw + (long)(y0 + k * kAugRows) * W + x) * c_pad + g) = pk;
      }
    }
  }
}

// Per-(sample, channel) mean and population standard deviation of an fp32 NCHW batch (norm_mode
// 'local', base_dataset.py:98-104).  One block per plane, fp64 accumulation of the sum and of the
// squared deviations from a first-pass mean (two passes, like numpy's mean / std).
constexpr int kStatThreads = 512;

__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();   // `red` may still be read from the previous call
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double s = 0.0;
  for (int w = 0; w < kStatThreads / 32; ++w) s += red[w];
  return s;
}

__global__ void __launch_bounds__(kStatThreads)
plane_mean_std_kernel(const float* __restrict__ img, double* __restrict__ mean,
                      double* __restrict__ stdv, long hw) {
  __shared__ double red[kStatThreads / 32];
  const float* plane = img + (long)blockIdx.x * hw;
  double s = 0.0;
  for (long i = threadIdx.x; i < hw; i += kStatThreads) s += (double)plane[i];
  const double m = block_sum(s, red) / (double)hw;
  double q = 0.0;
  for (long i = threadIdx.x; i < hw; i += kStatThreads) {
    const double d = (double)plane[i] - m;
    q = fma(d, d, q);
  }
  const double var = block_sum(q, red) / (double)hw;
  if (threadIdx.x == 0) {
    // the reference's statistics are fp32 (numpy mean/std of an fp32 array): round like it does
    mean[blockIdx.x] = (double)(float)m;
    stdv[blockIdx.x] = (double)(float)sqrt(var);
  }
}

}  // namespace fp

using namespace fp;

extern "C" {

int fpb200_augment_nchw_f32(const float* img, float* out_f32, void* out_nhwc_bf16, int c_pad,
                            const int64_t* tgt, int64_t* tgt_out, const float* theta,
                            const int* flags, const float* xgrid, const float* ygrid,
                            const double* mean, const double* stdv, int N, int C, int H, int W,
                            void* stream) {
  if (N < 1 || C < 1 || H < 1 || W < 1 || (out_f32 == nullptr && out_nhwc_bf16 == nullptr && tgt_out == nullptr))
    return FPB200_ERR_SHAPE;
  if (out_nhwc_bf16 != nullptr && (c_pad % 8 != 0 || c_pad < C)) return FPB200_ERR_SHAPE;
  if ((tgt == nullptr) != (tgt_out == nullptr) || (mean == nullptr) != (stdv == nullptr))
    return FPB200_ERR_SHAPE;
  if ((out_f32 != nullptr || out_nhwc_bf16 != nullptr) && img == nullptr) return FPB200_ERR_SHAPE;
  const int tiles_x = (W + kAugTile - 1) / kAugTile, tiles_y = (H + kAugTile - 1) / kAugTile;
  if ((long)N * tiles_x * tiles_y > 0x7fffffffL) return FPB200_ERR_SHAPE;
  augment_kernel<<<N * tiles_x * tiles_y, kAugTile * kAugRows, 0, (cudaStream_t)stream>>>(
      img, out_f32, (__nv_bfloat16*)out_nhwc_bf16, c_pad, tgt, tgt_out, theta, flags, xgrid, ygrid,
      mean, stdv, N, C, H, W, tiles_x, tiles_y);
  return check_launch("augment_nchw_f32");
}

int fpb200_plane_mean_std_f32(const float* img, double* mean, double* stdv, int planes, long hw,
                              void* stream) {
  if (planes < 1 || hw < 1) return FPB200_ERR_SHAPE;
  plane_mean_std_kernel<<<planes, kStatThreads, 0, (cudaStream_t)stream>>>(img, mean, stdv, hw);
  return check_launch("plane_mean_std_f32");
}

}  // extern "C"
