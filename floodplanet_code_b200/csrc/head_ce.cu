// OutConv 1x1 head (models/unet.py:70-77) and masked cross-entropy + argmax + confusion
// counts (models/water_seg_model.py:40,103-109): pure bandwidth passes over the 64-channel
// full-resolution map and the fp32 NCHW logits, with warp-shuffle reductions.
#include "host_common.h"
#include "ptx.cuh"

namespace fp {

constexpr int kMaxClasses = 8;
constexpr int kHeadC = 64;

__device__ __forceinline__ uint4 ld_nc_v4(const void* p) {
  uint4 v;
  asm("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
      : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

__device__ __forceinline__ void unpack8h(const uint4& u, float (&f)[8]) {
  f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x);
  f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z);
  f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}

// ---------------------------------------------------------------------------
// head forward: 8 lanes x 8 channels cover one pixel and a warp iteration covers 32 pixels
// (8 independent, fully coalesced 16-byte loads per lane in flight).  The 8 x NC partial dot
// products of a lane are summed across its 8-lane group with a TRANSPOSING butterfly (each
// step halves the number of live values), which leaves lane (sub, cg) with the finished
// logits of pixel base + 4*cg + sub: the 32 lanes then store 32 consecutive floats per class.
// (The first version used one thread per pixel: 16-byte loads 128 bytes apart and 3*64
// shared-memory weight reads per pixel held it at 3.2 TB/s.)
// ---------------------------------------------------------------------------
constexpr int kHeadFwdThreads = 256;

template <int NC>
__global__ void __launch_bounds__(kHeadFwdThreads)
head1x1_fwd_kernel(const __nv_bfloat16* __restrict__ x, long ldx, const float* __restrict__ w,
                   const float* __restrict__ b, float* __restrict__ logits, int N, long hw,
                   int ncls, const float* __restrict__ bn_scale, const float* __restrict__ bn_shift) {
  const int lane = threadIdx.x & 31;
  const int sub = lane >> 3;  // which pixel of a quad
  const int cg = lane & 7;    // channel group: channels cg*8 .. cg*8+7
  const bool fused_bn = bn_scale != nullptr;   // x is the raw conv output: a = relu(x*scale+shift)
  float wr[NC][8], sc[8], sh[8], bias[NC];
#pragma unroll
  for (int k = 0; k < NC; ++k) {
    bias[k] = k < ncls ? __ldg(b + k) : 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) wr[k][j] = k < ncls ? __ldg(w + k * kHeadC + cg * 8 + j) : 0.f;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = fused_bn ? __ldg(bn_scale + cg * 8 + j) : 1.f;
    sh[j] = fused_bn ? __ldg(bn_shift + cg * 8 + j) : 0.f;
  }
  const long total = (long)N * hw;
  const long warps_total = (long)gridDim.x * (kHeadFwdThreads / 32);
  for (long base = ((long)blockIdx.x * (kHeadFwdThreads / 32) + (threadIdx.x >> 5)) * 32; base < total;
       base += warps_total * 32) {
    uint4 q[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const long px = base + u * 4 + sub;
      q[u] = px < total ? ld_nc_v4(x + px * ldx + cg * 8) : make_uint4(0, 0, 0, 0);
    }
    float acc[8][NC];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      float f[8];
      unpack8h(q[u], f);
      if (fused_bn) {
        // same arithmetic and bf16 rounding as bn_apply_relu, so fusing changes no bit
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          const uint32_t pk = pack_bf16x2(fmaxf(fmaf(f[j], sc[j], sh[j]), 0.f),
                                          fmaxf(fmaf(f[j + 1], sc[j + 1], sh[j + 1]), 0.f));
          f[j] = bf16_lo(pk);
          f[j + 1] = bf16_hi(pk);
        }
      }
#pragma unroll
      for (int k = 0; k < NC; ++k) {
        float a = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) a = fmaf(f[j], wr[k][j], a);
        acc[u][k] = a;
      }
    }
    // transposing butterfly over the 8 lanes of a pixel: after the step with lane mask m a lane
    // keeps the pixels u whose bit m equals its own bit m
    float r4[4][NC], r2[2][NC], r1[NC];
    const bool b4 = cg & 4, b2 = cg & 2, b1 = cg & 1;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int k = 0; k < NC; ++k) {
        const float mine = b4 ? acc[i + 4][k] : acc[i][k];
        const float theirs = b4 ? acc[i][k] : acc[i + 4][k];
        r4[i][k] = mine + __shfl_xor_sync(0xffffffffu, theirs, 4);
      }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int k = 0; k < NC; ++k) {
        const float mine = b2 ? r4[i + 2][k] : r4[i][k];
        const float theirs = b2 ? r4[i][k] : r4[i + 2][k];
        r2[i][k] = mine + __shfl_xor_sync(0xffffffffu, theirs, 2);
      }
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      const float mine = b1 ? r2[1][k] : r2[0][k];
      const float theirs = b1 ? r2[0][k] : r2[1][k];
      r1[k] = mine + __shfl_xor_sync(0xffffffffu, theirs, 1);
    }
    const long px = base + cg * 4 + sub;   // u == cg
    if (px < total) {
      const long n = px / hw, o = px - n * hw;
#pragma unroll
      for (int k = 0; k < NC; ++k)
        if (k < ncls) logits[((long)n * ncls + k) * hw + o] = r1[k] + bias[k];
    }
  }
}

// ---------------------------------------------------------------------------
// head backward: 8 lanes x 8 channels cover one pixel, kHeadBwdPx pixels per warp iteration
// with all of a lane's 16-byte loads issued before the first use.
//   dx[p, c]  = sum_k dl[k, p] * w[k, c]
//   dW[k, c] += dl[k, p] * x[p, c],   db[k] += dl[k, p]
// ---------------------------------------------------------------------------
constexpr int kHeadBwdThreads = 256;
constexpr int kHeadBwdPx = 16;   // pixels per warp iteration (4 per quad lane; 8 spills at 128 registers)
constexpr int kHeadBwdStages = 2;  // cp.async ring of x vectors per warp: this iteration + the next

template <int NC, bool FUSED>
__global__ void __launch_bounds__(kHeadBwdThreads, 2)
head1x1_bwd_kernel(const float* __restrict__ dlogits, const __nv_bfloat16* __restrict__ x, long ldx,
                   const float* __restrict__ w, __nv_bfloat16* __restrict__ dx, long lddx,
                   float* __restrict__ partials, int N, long hw, int ncls,
                   const float* __restrict__ bn_scale, const float* __restrict__ bn_shift,
                   const float* __restrict__ bn_mean, const float* __restrict__ bn_invstd,
                   float* __restrict__ bn_partials) {
  __shared__ float sw[NC * kHeadC];
  // the cp.async ring of the main loop and the block-reduction scratch of the tail share one allocation
  constexpr int kRingBytes = (kHeadBwdThreads / 32) * kHeadBwdStages * kHeadBwdPx * 8 * 16;
  constexpr int kRedBytes = (kHeadBwdThreads / 32) * NC * (kHeadC + 1) * 4;
  __shared__ __align__(16) uint8_t pool[kRingBytes > kRedBytes ? kRingBytes : kRedBytes];
  float* red = reinterpret_cast<float*>(pool);
  uint4* xs = reinterpret_cast<uint4*>(pool);
  for (int i = threadIdx.x; i < NC * kHeadC; i += blockDim.x) sw[i] = i < ncls * kHeadC ? w[i] : 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane >> 3;  // which pixel of a quad
  const int cg = lane & 7;    // channel group: channels cg*8 .. cg*8+7
  // accumulators and per-pixel values are kept as fp32 PAIRS: the kernel is issue bound (~150 instructions per
  // 8-channel pixel slice, half of them FMAs), and FFMA2 / FADD2 do two of them per issue slot with the same bits
  float2 dw[NC][4];
  float db[NC];
#pragma unroll
  for (int k = 0; k < NC; ++k) {
    db[k] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) dw[k][j] = make_float2(0.f, 0.f);
  }
  // fused mode: x is the raw conv output y of the last DoubleConv layer; the activation is
  // recomputed on load and the layer's BatchNorm-backward sums (sum g, sum g*xhat with
  // g = dx * [a > 0]) are accumulated here, where dx is produced, instead of in a separate pass
  // (the second sum is accumulated as sum g*y and converted to sum g*xhat = invstd*(sum g*y -
  // mean*sum g) once per block, which keeps mean/invstd out of the per-pixel loop)
  constexpr bool fused_bn = FUSED;
  float2 sc[4], sh[4], s1[4], s2[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    sc[j] = fused_bn ? make_float2(__ldg(bn_scale + cg * 8 + 2 * j), __ldg(bn_scale + cg * 8 + 2 * j + 1))
                     : make_float2(1.f, 1.f);
    sh[j] = fused_bn ? make_float2(__ldg(bn_shift + cg * 8 + 2 * j), __ldg(bn_shift + cg * 8 + 2 * j + 1))
                     : make_float2(0.f, 0.f);
    s1[j] = make_float2(0.f, 0.f);
    s2[j] = make_float2(0.f, 0.f);
  }
  const long total = (long)N * hw;
  const long warps_total = (long)gridDim.x * (kHeadBwdThreads / 32);
  constexpr int U = kHeadBwdPx / 4;
  // The x vectors of the NEXT iteration are fetched with cp.async into a per-warp shared-memory ring while this
  // iteration computes: with register loads the warp alternated between a load phase and ~600 issue slots of
  // arithmetic with nothing in flight (3.6 TB/s at 16 warps per SM).  Every lane reads back exactly the 16 bytes it
  // copied itself, so the ring needs no barrier -- cp.async.wait_group is per thread.
  const uint32_t ring = smem_u32(xs) + (uint32_t)(warp * kHeadBwdStages * kHeadBwdPx * 8 + lane) * 16u;   // + (stage*16 + u*4)*128
  const long x_quad = 4 * ldx, dx_quad = 4 * lddx;       // element stride between a lane's consecutive pixels
  auto prefetch = [&](long b, int stage) {
    const __nv_bfloat16* src = x + (b + sub) * ldx + cg * 8;
    if (b + kHeadBwdPx <= total) {
#pragma unroll
      for (int u = 0; u < U; ++u)
        cp_async_16(ring + (uint32_t)(stage * kHeadBwdPx + u * 4) * 128u, src + u * x_quad, 16u);
    } else {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const bool in = b + u * 4 + sub < total;
        cp_async_16(ring + (uint32_t)(stage * kHeadBwdPx + u * 4) * 128u, in ? src + u * x_quad : x, in ? 16u : 0u);
      }
    }
    cp_async_commit();
  };
  // dlogits of the iteration at `b`: NC coalesced loads, lane L < 16 holds pixel b + L, handed to the 8 lanes of each
  // pixel by shuffles later.  (img, off) = (b / hw, b % hw) is carried along incrementally: the per-iteration 64-bit
  // division was a quarter of the kernel's integer instructions.
  auto load_dl = [&](long b, long img, long off, float (&v)[NC]) {
    long o = off + lane, n = img;
    while (o >= hw) { o -= hw; ++n; }
    const bool in = lane < kHeadBwdPx && b + lane < total;
    const float* src = dlogits + (n * ncls) * hw + o;
#pragma unroll
    for (int k = 0; k < NC; ++k) v[k] = (in && k < ncls) ? __ldg(src + k * hw) : 0.f;
  };
  const long step = warps_total * kHeadBwdPx;
  const long step_img = step / hw, step_off = step - step_img * hw;
  const long base0 = ((long)blockIdx.x * (kHeadBwdThreads / 32) + warp) * kHeadBwdPx;
  long img = base0 / hw, off = base0 - img * hw;
  float dlv[NC];
#pragma unroll
  for (int k = 0; k < NC; ++k) dlv[k] = 0.f;
  if (base0 < total) {
    prefetch(base0, 0);
    load_dl(base0, img, off, dlv);
  }
  int stage = 0;
  for (long base = base0; base < total; base += step) {
    // software pipeline: the NEXT iteration's x vectors (cp.async) and dlogits (registers) are requested before this
    // iteration's arithmetic; ncu had 21 % of all samples on the first use of the dlogits loaded in the same iteration
    const long next = base + step;
    float dl_next[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) dl_next[k] = 0.f;
    img += step_img;
    off += step_off;
    if (off >= hw) { off -= hw; ++img; }
    if (next < total) {
      prefetch(next, stage ^ 1);
      load_dl(next, img, off, dl_next);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    const uint32_t cur = ring + (uint32_t)(stage * kHeadBwdPx) * 128u;
    stage ^= 1;
    __nv_bfloat16* dst = dx + (base + sub) * lddx + cg * 8;
    const bool full = base + kHeadBwdPx <= total;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float dl[NC];
#pragma unroll
      for (int k = 0; k < NC; ++k) dl[k] = __shfl_sync(0xffffffffu, dlv[k], u * 4 + sub);
      const uint4 xv = ld_shared_v4(cur + (uint32_t)(u * 4) * 128u);   // read where it is used: 12 fewer live registers
      const uint32_t xw[4] = {xv.x, xv.y, xv.z, xv.w};
      float2 f[4], g[4], yraw[4], pre[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        f[j] = make_float2(bf16_lo(xw[j]), bf16_hi(xw[j]));
        if (fused_bn) {
          yraw[j] = f[j];
          pre[j] = fma2(f[j], sc[j], sh[j]);       // y * scale + shift: its sign is the ReLU mask
          const uint32_t pk = pack_bf16x2(fmaxf(pre[j].x, 0.f), fmaxf(pre[j].y, 0.f));
          f[j] = make_float2(bf16_lo(pk), bf16_hi(pk));
        }
        g[j] = make_float2(0.f, 0.f);
      }
#pragma unroll
      for (int k = 0; k < NC; ++k) {
        const float4 w0 = *reinterpret_cast<const float4*>(sw + k * kHeadC + cg * 8);
        const float4 w1 = *reinterpret_cast<const float4*>(sw + k * kHeadC + cg * 8 + 4);
        const float2 wk[4] = {make_float2(w0.x, w0.y), make_float2(w0.z, w0.w), make_float2(w1.x, w1.y),
                              make_float2(w1.z, w1.w)};
        const float2 dl2 = make_float2(dl[k], dl[k]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          g[j] = fma2(dl2, wk[j], g[j]);
          dw[k][j] = fma2(dl2, f[j], dw[k][j]);
        }
        db[k] += dl[k];
      }
      uint4 o4;
      o4.x = pack_bf16x2(g[0].x, g[0].y);
      o4.y = pack_bf16x2(g[1].x, g[1].y);
      o4.z = pack_bf16x2(g[2].x, g[2].y);
      o4.w = pack_bf16x2(g[3].x, g[3].y);
      if (full || base + u * 4 + sub < total) *reinterpret_cast<uint4*>(dst + u * dx_quad) = o4;
      if (fused_bn) {
        // the sums use the bf16-rounded dx that is stored (what the apply pass will read);
        // out-of-range pixels have dl = 0, hence contribute exactly 0
        const uint32_t ow[4] = {o4.x, o4.y, o4.z, o4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 gg = make_float2(pre[j].x > 0.f ? bf16_lo(ow[j]) : 0.f, pre[j].y > 0.f ? bf16_hi(ow[j]) : 0.f);
          s1[j] = add2(s1[j], gg);
          s2[j] = fma2(gg, yraw[j], s2[j]);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < NC; ++k) dlv[k] = dl_next[k];
  }
  // reduce the 4 pixel sub-lanes of the warp
  __syncthreads();   // every warp is done with the ring: `red` aliases it
  auto quad_sum = [](float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    return v;
  };
#pragma unroll
  for (int k = 0; k < NC; ++k) {
#pragma unroll
    for (int j = 0; j < 4; ++j) dw[k][j] = make_float2(quad_sum(dw[k][j].x), quad_sum(dw[k][j].y));
    db[k] = quad_sum(db[k]);
  }
  const int stride = kHeadC + 1;
  if (sub == 0) {
#pragma unroll
    for (int k = 0; k < NC; ++k) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        red[(warp * NC + k) * stride + cg * 8 + 2 * j] = dw[k][j].x;
        red[(warp * NC + k) * stride + cg * 8 + 2 * j + 1] = dw[k][j].y;
      }
      if (cg == 0) red[(warp * NC + k) * stride + kHeadC] = db[k];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < ncls * stride; i += blockDim.x) {
    const int k = i / stride, c = i - k * stride;
    float acc = 0.f;
    for (int wv = 0; wv < kHeadBwdThreads / 32; ++wv) acc += red[(wv * NC + k) * stride + c];
    partials[(size_t)blockIdx.x * ncls * stride + i] = acc;
  }
  if (fused_bn) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      s1[j] = make_float2(quad_sum(s1[j].x), quad_sum(s1[j].y));
      s2[j] = make_float2(quad_sum(s2[j].x), quad_sum(s2[j].y));
    }
    __syncthreads();   // `red` is reused: [warp][2][64]
    if (sub == 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        red[(warp * 2 + 0) * kHeadC + cg * 8 + 2 * j] = s1[j].x;
        red[(warp * 2 + 0) * kHeadC + cg * 8 + 2 * j + 1] = s1[j].y;
        red[(warp * 2 + 1) * kHeadC + cg * 8 + 2 * j] = s2[j].x;
        red[(warp * 2 + 1) * kHeadC + cg * 8 + 2 * j + 1] = s2[j].y;
      }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < kHeadC; c += blockDim.x) {
      float a1 = 0.f, a2 = 0.f;
      for (int wv = 0; wv < kHeadBwdThreads / 32; ++wv) {
        a1 += red[(wv * 2 + 0) * kHeadC + c];
        a2 += red[(wv * 2 + 1) * kHeadC + c];
      }
      bn_partials[(size_t)blockIdx.x * 2 * kHeadC + c] = a1;                                   // sum g
      bn_partials[(size_t)blockIdx.x * 2 * kHeadC + kHeadC + c] =
          __ldg(bn_invstd + c) * (a2 - __ldg(bn_mean + c) * a1);                               // sum g*xhat
    }
  }
}

__global__ void head_bwd_finalize_kernel(const float* __restrict__ partials, int P, int ncls,
                                         float* dw, float* db) {
  const int stride = kHeadC + 1;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ncls * stride) return;
  double acc = 0.0;
  for (int p = 0; p < P; ++p) acc += (double)partials[(size_t)p * ncls * stride + i];
  const int k = i / stride, c = i - k * stride;
  if (c < kHeadC) dw[k * kHeadC + c] = (float)acc;
  else db[k] = (float)acc;
}

// ---------------------------------------------------------------------------
// masked softmax cross-entropy + argmax + confusion counts
// ---------------------------------------------------------------------------
constexpr int kCeThreads = 256;

// V pixels per thread (V = 4: 16-byte loads of each class plane, 2 x 16 bytes of int64 target and
// prediction; V = 1 when hw is not a multiple of 4 or a pointer is not 16-byte aligned).  The confusion
// counts are aggregated per warp with match_any (one shared-memory atomic per distinct (target,
// prediction) pair and warp instead of one per pixel: with two live target classes the per-pixel
// atomics were 32-way conflicts).  The loop trip count is block-uniform so that every lane takes part
// in the warp votes.
template <int V, int NC>
__global__ void __launch_bounds__(kCeThreads)
softmax_ce_argmax_fwd_kernel(const float* __restrict__ logits, const int64_t* __restrict__ target,
                             long ignore_index, int64_t* __restrict__ pred,
                             unsigned long long* __restrict__ confusion,
                             double* __restrict__ partials, int N, int ncls, long hw) {
  __shared__ unsigned int s_conf[kMaxClasses * kMaxClasses];
  __shared__ double s_red[3][kCeThreads / 32];
  if (threadIdx.x < kMaxClasses * kMaxClasses) s_conf[threadIdx.x] = 0;
  __syncthreads();
  double loss_sum = 0.0;
  double cnt = 0.0, bad = 0.0;
  const long groups = ((long)N * hw) / V;          // V divides hw
  const long stride = (long)gridDim.x * blockDim.x;
  const long iters = (groups + stride - 1) / stride;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (long it = 0; it < iters; ++it) {
    const long gi = it * stride + blockIdx.x * (long)blockDim.x + threadIdx.x;
    const bool in = gi < groups;
    const long px = gi * V;
    const long n = in ? px / hw : 0, o = in ? px - n * hw : 0;
    float l[NC][V];
    int64_t t[V];
    if (in) {
#pragma unroll
      for (int k = 0; k < NC; ++k) {
        if (k < ncls) {
          const float* src = logits + ((long)n * ncls + k) * hw + o;
          if (V == 4) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(src));
            l[k][0] = v.x; l[k][1 % V] = v.y; l[k][2 % V] = v.z; l[k][3 % V] = v.w;
          } else {
            l[k][0] = __ldg(src);
          }
        }
      }
      if (V == 4) {
        const longlong2 a = __ldg(reinterpret_cast<const longlong2*>(target + px));
        const longlong2 b2 = __ldg(reinterpret_cast<const longlong2*>(target + px) + 1);
        t[0] = a.x; t[1 % V] = a.y; t[2 % V] = b2.x; t[3 % V] = b2.y;
      } else {
        t[0] = target[px];
      }
    }
    int64_t pv[V];
#pragma unroll
    for (int j = 0; j < V; ++j) {
      int key = -1;
      if (in) {
        float best = 0.f;
        int bi = 0;
#pragma unroll
        for (int k = 0; k < NC; ++k) {
          if (k < ncls) {
            // first maximum wins; NaN counts as the maximum (torch.argmax semantics)
            if (k == 0 || (l[k][j] > best) || (l[k][j] != l[k][j] && best == best)) { best = l[k][j]; bi = k; }
          }
        }
        pv[j] = bi;
        if (t[j] != ignore_index) {
          if (t[j] < 0 || t[j] >= ncls) {
            bad += 1.0;
          } else {
            float mx = l[0][j];
#pragma unroll
            for (int k = 1; k < NC; ++k)
              if (k < ncls) mx = fmaxf(mx, l[k][j]);
            float se = 0.f, lt = 0.f;
#pragma unroll
            for (int k = 0; k < NC; ++k) {
              if (k < ncls) {
                se += expf(l[k][j] - mx);
                if (k == (int)t[j]) lt = l[k][j];
              }
            }
            loss_sum += (double)((mx + logf(se)) - lt);
            cnt += 1.0;
            key = (int)t[j] * kMaxClasses + bi;
          }
        }
      }
      if (confusion != nullptr) {
        const unsigned same = __match_any_sync(0xffffffffu, key);
        if (key >= 0 && lane == __ffs(same) - 1) atomicAdd(&s_conf[key], (unsigned)__popc(same));
      }
    }
    if (in && pred != nullptr) {
      if (V == 4) {
        longlong2* dst = reinterpret_cast<longlong2*>(pred + px);
        dst[0] = make_longlong2(pv[0], pv[1 % V]);
        dst[1] = make_longlong2(pv[2 % V], pv[3 % V]);
      } else {
        pred[px] = pv[0];
      }
    }
  }
  // block reduce
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    loss_sum += __shfl_xor_sync(0xffffffffu, loss_sum, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    bad += __shfl_xor_sync(0xffffffffu, bad, o);
  }
  if (lane == 0) { s_red[0][warp] = loss_sum; s_red[1][warp] = cnt; s_red[2][warp] = bad; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b = 0, c = 0;
    for (int wv = 0; wv < kCeThreads / 32; ++wv) { a += s_red[0][wv]; b += s_red[1][wv]; c += s_red[2][wv]; }
    partials[(size_t)blockIdx.x * 4 + 0] = a;
    partials[(size_t)blockIdx.x * 4 + 1] = b;
    partials[(size_t)blockIdx.x * 4 + 2] = c;
  }
  if (confusion != nullptr && threadIdx.x < kMaxClasses * kMaxClasses) {
    const int t = threadIdx.x / kMaxClasses, pcls = threadIdx.x % kMaxClasses;
    const unsigned int v = s_conf[threadIdx.x];
    if (v != 0 && t < ncls && pcls < ncls)
      atomicAdd(&confusion[t * ncls + pcls], (unsigned long long)v);
  }
}

__global__ void ce_finalize_kernel(const double* __restrict__ partials, int P, double* result) {
  // single warp
  double a = 0, b = 0, c = 0;
  for (int p = threadIdx.x; p < P; p += 32) {
    a += partials[(size_t)p * 4 + 0];
    b += partials[(size_t)p * 4 + 1];
    c += partials[(size_t)p * 4 + 2];
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
    c += __shfl_xor_sync(0xffffffffu, c, o);
  }
  if (threadIdx.x == 0) {
    result[0] = a;
    result[1] = b;
    result[2] = c;
    result[3] = a / b;  // 0/0 = NaN, like torch's mean over an empty set
  }
}

template <int V, int NC>
__global__ void __launch_bounds__(kCeThreads)
softmax_ce_bwd_kernel(const float* __restrict__ logits, const int64_t* __restrict__ target,
                      long ignore_index, const double* __restrict__ result,
                      const float* __restrict__ grad_out, float* __restrict__ dlogits, int N,
                      int ncls, long hw) {
  const double count = result[1];
  const float go = grad_out != nullptr ? __ldg(grad_out) : 1.f;
  const float coef = count > 0.0 ? go / (float)count : 0.f;
  const long groups = ((long)N * hw) / V;
  for (long gi = blockIdx.x * (long)blockDim.x + threadIdx.x; gi < groups;
       gi += (long)gridDim.x * blockDim.x) {
    const long px = gi * V;
    const long n = px / hw, o = px - n * hw;
    int64_t t[V];
    if (V == 4) {
      const longlong2 a = __ldg(reinterpret_cast<const longlong2*>(target + px));
      const longlong2 b2 = __ldg(reinterpret_cast<const longlong2*>(target + px) + 1);
      t[0] = a.x; t[1 % V] = a.y; t[2 % V] = b2.x; t[3 % V] = b2.y;
    } else {
      t[0] = target[px];
    }
    bool active[V], any = false;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      active[j] = (t[j] != ignore_index) && t[j] >= 0 && t[j] < ncls && coef != 0.f;
      any |= active[j];
    }
    float l[NC][V];
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      if (k < ncls) {
        const float* src = logits + ((long)n * ncls + k) * hw + o;
        if (V == 4) {
          // logits of fully ignored groups are not read at all (58 % of the default labels)
          const float4 v = any ? __ldg(reinterpret_cast<const float4*>(src)) : make_float4(0.f, 0.f, 0.f, 0.f);
          l[k][0] = v.x; l[k][1 % V] = v.y; l[k][2 % V] = v.z; l[k][3 % V] = v.w;
        } else {
          l[k][0] = any ? __ldg(src) : 0.f;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < V; ++j) {
      float mx = -INFINITY;
#pragma unroll
      for (int k = 0; k < NC; ++k)
        if (k < ncls) { if (!active[j]) l[k][j] = 0.f; mx = fmaxf(mx, l[k][j]); }
      float se = 0.f;
#pragma unroll
      for (int k = 0; k < NC; ++k)
        if (k < ncls) { l[k][j] = expf(l[k][j] - mx); se += l[k][j]; }
      const float inv = 1.f / se;
#pragma unroll
      for (int k = 0; k < NC; ++k)
        if (k < ncls) l[k][j] = active[j] ? (l[k][j] * inv - (k == (int)t[j] ? 1.f : 0.f)) * coef : 0.f;
    }
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      if (k < ncls) {
        float* dst = dlogits + ((long)n * ncls + k) * hw + o;
        if (V == 4) *reinterpret_cast<float4*>(dst) = make_float4(l[k][0], l[k][1 % V], l[k][2 % V], l[k][3 % V]);
        else dst[0] = l[k][0];
      }
    }
  }
}

// ---------------------------------------------------------------------------
// sliding-window inference post-processing (infer.py:122-184, utils_image.py:410-494):
// softmax over classes, scatter-add of tile probabilities + weights into the scene canvas,
// then canvas/(w+1e-5) -> argmax -> clip(0,1)*255 as uint8.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
softmax_stitch_add_kernel(const float* __restrict__ logits, float* __restrict__ canvas,
                          float* __restrict__ weight, const int* __restrict__ tiles, int n, int ncls,
                          int th, int tw, long H, long W) {
  const long hw = (long)th * tw;
  const long total = (long)n * hw;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total;
       i += (long)gridDim.x * blockDim.x) {
    const int t = (int)(i / hw);
    const long o = i - (long)t * hw;
    const int y = (int)(o / tw), x = (int)(o - (long)y * tw);
    const int h0 = tiles[t * 4 + 0], w0 = tiles[t * 4 + 1], hh = tiles[t * 4 + 2], ww = tiles[t * 4 + 3];
    const long gy = h0 + y, gx = w0 + x;
    if (y >= hh || x >= ww || gy >= H || gx >= W) continue;
    float l[kMaxClasses];
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < kMaxClasses; ++k)
      if (k < ncls) { l[k] = __ldg(logits + ((long)t * ncls + k) * hw + o); mx = fmaxf(mx, l[k]); }
    float se = 0.f;
#pragma unroll
    for (int k = 0; k < kMaxClasses; ++k)
      if (k < ncls) { l[k] = expf(l[k] - mx); se += l[k]; }
    const float inv = 1.f / se;
    float* dst = canvas + (gy * W + gx) * ncls;
#pragma unroll
    for (int k = 0; k < kMaxClasses; ++k)
      if (k < ncls) atomicAdd(dst + k, l[k] * inv);
    atomicAdd(weight + gy * W + gx, 1.f);
  }
}

__global__ void __launch_bounds__(256)
canvas_to_mask_kernel(const float* __restrict__ canvas, const float* __restrict__ weight,
                      uint8_t* __restrict__ mask, long npix, int ncls) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < npix;
       i += (long)gridDim.x * blockDim.x) {
    const float inv = 1.f / (weight[i] + 1e-5f);
    float best = 0.f;
    int bi = 0;
#pragma unroll
    for (int k = 0; k < kMaxClasses; ++k) {
      if (k < ncls) {
        float v = canvas[i * ncls + k] * inv;
        if (v != v) v = 0.f;  // np.nan_to_num
        if (k == 0 || v > best) { best = v; bi = k; }
      }
    }
    mask[i] = bi >= 1 ? 255 : 0;  // np.clip(argmax, 0, 1) * 255
  }
}

}  // namespace fp

using namespace fp;

extern "C" {

int fpb200_head1x1_fwd(const void* x, long ldx, const float* w, const float* b, float* logits,
                       int N, int H, int W, int C, int n_classes, const float* bn_scale,
                       const float* bn_shift, void* stream) {
  if (C != kHeadC || n_classes < 1 || n_classes > kMaxClasses || ldx % 8 != 0)
    return FPB200_ERR_SHAPE;
  const long total = (long)N * H * W;
  long g = (total + kHeadFwdThreads - 1) / kHeadFwdThreads;   // 32 pixels per warp iteration
  if (g > 8L * sm_count()) g = 8L * sm_count();
#define FP_HEAD_FWD(nc)                                                                         \
  head1x1_fwd_kernel<nc><<<(int)g, kHeadFwdThreads, 0, (cudaStream_t)stream>>>(                 \
      (const __nv_bfloat16*)x, ldx, w, b, logits, N, (long)H * W, n_classes, bn_scale, bn_shift)
  if (n_classes <= 2) FP_HEAD_FWD(2);
  else if (n_classes == 3) FP_HEAD_FWD(3);
  else if (n_classes == 4) FP_HEAD_FWD(4);
  else FP_HEAD_FWD(8);
#undef FP_HEAD_FWD
  return check_launch("head1x1_fwd");
}

int fpb200_head_bwd_rows(void) { return 8 * sm_count(); }

int fpb200_head1x1_bwd(const float* dlogits, const void* x, long ldx, const float* w, void* dx,
                       long lddx, float* dw, float* db, float* partials, int N, int H, int W,
                       int C, int n_classes, const float* bn_scale, const float* bn_shift,
                       const float* bn_mean, const float* bn_invstd, float* bn_partials,
                       void* stream) {
  if (C != kHeadC || n_classes < 1 || n_classes > kMaxClasses || ldx % 8 != 0 || lddx % 8 != 0)
    return FPB200_ERR_SHAPE;
  if (bn_scale != nullptr && (!bn_shift || !bn_mean || !bn_invstd || !bn_partials)) return FPB200_ERR_SHAPE;
  const int rows = fpb200_head_bwd_rows();
#define FP_HEAD_BWD(nc, fused)                                                                   \
  head1x1_bwd_kernel<nc, fused><<<rows, kHeadBwdThreads, 0, (cudaStream_t)stream>>>(             \
      dlogits, (const __nv_bfloat16*)x, ldx, w, (__nv_bfloat16*)dx, lddx, partials, N, (long)H * W, \
      n_classes, bn_scale, bn_shift, bn_mean, bn_invstd, bn_partials)
  if (n_classes <= 3) {   // the reference's n_classes = 3: no padded class in registers
    if (bn_scale != nullptr) FP_HEAD_BWD(3, true); else FP_HEAD_BWD(3, false);
  } else if (n_classes == 4) {
    if (bn_scale != nullptr) FP_HEAD_BWD(4, true); else FP_HEAD_BWD(4, false);
  } else {
    if (bn_scale != nullptr) FP_HEAD_BWD(8, true); else FP_HEAD_BWD(8, false);
  }
#undef FP_HEAD_BWD
  int rc = check_launch("head1x1_bwd");
  if (rc != FPB200_OK) return rc;
  const int n = n_classes * (kHeadC + 1);
  head_bwd_finalize_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(partials, rows,
                                                                              n_classes, dw, db);
  return check_launch("head1x1_bwd_finalize");
}

int fpb200_ce_rows(void) { return 8 * sm_count(); }

int fpb200_softmax_ce_argmax_fwd(const float* logits, const int64_t* target, long ignore_index,
                                 double* result, int64_t* pred, int64_t* confusion,
                                 double* partials, int N, int n_classes, long hw, void* stream) {
  if (n_classes < 1 || n_classes > kMaxClasses || N < 1 || hw < 1) return FPB200_ERR_SHAPE;
  const long total = (long)N * hw;
  const bool vec = hw % 4 == 0 && ((reinterpret_cast<uintptr_t>(logits) | reinterpret_cast<uintptr_t>(target) |
                                    reinterpret_cast<uintptr_t>(pred)) & 15) == 0;
  long g = (total / (vec ? 4 : 1) + kCeThreads - 1) / kCeThreads;
  const int rows = fpb200_ce_rows();
  if (g > rows) g = rows;
#define FP_CE_FWD(v, nc)                                                                          \
  softmax_ce_argmax_fwd_kernel<v, nc><<<(int)g, kCeThreads, 0, (cudaStream_t)stream>>>(             \
      logits, target, ignore_index, pred, (unsigned long long*)confusion, partials, N, n_classes, hw)
  if (vec && n_classes <= 4) FP_CE_FWD(4, 4);
  else if (vec) FP_CE_FWD(4, kMaxClasses);
  else FP_CE_FWD(1, kMaxClasses);
#undef FP_CE_FWD
  int rc = check_launch("softmax_ce_argmax_fwd");
  if (rc != FPB200_OK) return rc;
  ce_finalize_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(partials, (int)g, result);
  return check_launch("softmax_ce_finalize");
}

int fpb200_softmax_ce_bwd(const float* logits, const int64_t* target, long ignore_index,
                          const double* result, const float* grad_out, float* dlogits, int N,
                          int n_classes, long hw, void* stream) {
  if (n_classes < 1 || n_classes > kMaxClasses) return FPB200_ERR_SHAPE;
  const long total = (long)N * hw;
  const bool vec = hw % 4 == 0 && ((reinterpret_cast<uintptr_t>(logits) | reinterpret_cast<uintptr_t>(target) |
                                    reinterpret_cast<uintptr_t>(dlogits)) & 15) == 0;
  long g = (total / (vec ? 4 : 1) + kCeThreads - 1) / kCeThreads;
  if (g > 148L * 16) g = 148L * 16;
#define FP_CE_BWD(v, nc)                                                                          \
  softmax_ce_bwd_kernel<v, nc><<<(int)g, kCeThreads, 0, (cudaStream_t)stream>>>(                    \
      logits, target, ignore_index, result, grad_out, dlogits, N, n_classes, hw)
  if (vec && n_classes <= 4) FP_CE_BWD(4, 4);
  else if (vec) FP_CE_BWD(4, kMaxClasses);
  else FP_CE_BWD(1, kMaxClasses);
#undef FP_CE_BWD
  return check_launch("softmax_ce_bwd");
}

int fpb200_softmax_stitch_add(const float* logits, float* canvas, float* weight, const int* tiles,
                              int n_tiles, int n_classes, int th, int tw, long H, long W,
                              void* stream) {
  if (n_classes < 1 || n_classes > kMaxClasses || n_tiles < 1) return FPB200_ERR_SHAPE;
  const long total = (long)n_tiles * th * tw;
  long g = (total + 255) / 256;
  if (g > 148L * 16) g = 148L * 16;
  softmax_stitch_add_kernel<<<(int)g, 256, 0, (cudaStream_t)stream>>>(logits, canvas, weight, tiles,
                                                                      n_tiles, n_classes, th, tw, H, W);
  return check_launch("softmax_stitch_add");
}

int fpb200_canvas_to_mask_u8(const float* canvas, const float* weight, uint8_t* mask, long npix,
                             int n_classes, void* stream) {
  if (n_classes < 1 || n_classes > kMaxClasses) return FPB200_ERR_SHAPE;
  long g = (npix + 255) / 256;
  if (g > 148L * 16) g = 148L * 16;
  canvas_to_mask_kernel<<<(int)g, 256, 0, (cudaStream_t)stream>>>(canvas, weight, mask, npix, n_classes);
  return check_launch("canvas_to_mask_u8");
}

}  // extern "C"
