// Microbenchmark of the tcgen05.mma OPERAND PATH on B200 (evidence for DESIGN.md section 3.1 / VERDICT r1 #3):
// how many cycles does one M128 x N x K16 bf16 MMA take when it is issued back to back from fixed shared
// memory operands -- no TMA, no epilogue, nothing else running on the SM -- as a function of N, of the
// A-operand source (shared-memory descriptor "SS" vs tensor memory "TS"), of the A start address (1024-byte
// aligned vs the 128-byte-aligned, 18-row-strided tap-shifted descriptors of the halo kernel), and what
// staging A into tensor memory with tcgen05.cp costs.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I floodplanet_code_b200/csrc \
//        scripts/umma_microbench.cu -o gpurun_out/umma_microbench && gpurun_out/umma_microbench
//
// Tensor-pipe floor per MMA: 128 * N / 256 cycles (N = 64: 32, N = 128: 64, N = 256: 128).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "ptx.cuh"

using namespace fp;

enum Mode { SS_ALIGNED = 0, SS_TAPSHIFT = 1, TS = 2, CP_ONLY = 3, CP_PLUS_TS = 4, SS_SAME_A = 5 };

__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void utccp_128x256b(uint32_t tmem_dst, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(tmem_dst), "l"(sdesc) : "memory");
}

// smem: A region = an 18 x 18 pixel box of 128-byte rows (as in the halo kernel) = 41472 B (+ slack), B region = 256 rows
constexpr int kABytes = 48 * 1024;
constexpr int kBBytes = 256 * 128;

template <int N, int MODE>
__global__ void __launch_bounds__(128, 1) umma_bench(int iters, long long* out_cycles, int* out_count) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = base, b_base = base + kABytes;
  const uint32_t bar = base + kABytes + kBBytes, slot = bar + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // deterministic, finite operand contents (zeros): timing does not depend on values
  for (uint32_t o = threadIdx.x * 16; o < kABytes + kBBytes; o += blockDim.x * 16)
    st_shared_v4(base + o, make_uint4(0, 0, 0, 0));
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem_raw + (slot - smem_u32(smem_raw)));
  constexpr uint32_t kIdesc = make_idesc_bf16(128, N, 0, 0);
  constexpr uint32_t kRB = 128;
  long long t0 = 0, t1 = 0;
  int count = 0;
  if (warp == 0) {
    const bool leader = elect_one();
    const uint32_t b_lo = smem_desc_lo(b_base, 16);
    constexpr uint32_t kBHi = smem_desc_hi(8 * kRB, 128);
    constexpr uint32_t kAHiAligned = smem_desc_hi(8 * kRB, 128);        // dense 128-row tile, 8-row groups 1024 B apart
    constexpr uint32_t kAHiBox = smem_desc_hi(18 * kRB, 128);           // halo box: 8-row groups 18 rows apart
    const uint32_t a_lo = smem_desc_lo(a_base, 16);
    const uint32_t tmem_a = tmem + 256;                                 // A staging columns (TS modes): 4 slices x 8 columns x 2 buffers
    __syncwarp();
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int r = tap / 3, s = tap - 3 * r;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t acc = (it | tap | k) != 0 ? 1u : 0u;
            const uint32_t d = tmem + (t * N) % 256;          // N = 256: both tiles share one accumulator
            if (MODE == SS_ALIGNED) {
              if (leader) umma_bf16(d, smem_desc_join(a_lo + ((t * 16384 + k * 32) >> 4), kAHiAligned),
                                    smem_desc_join(b_lo + ((k * 32) >> 4), kBHi), kIdesc, acc);
            } else if (MODE == SS_TAPSHIFT) {
              const uint32_t a_off = (uint32_t(r * 18 + t * 8 + s) * kRB + k * 32) >> 4;
              if (leader) umma_bf16(d, smem_desc_join(a_lo + a_off, kAHiBox),
                                    smem_desc_join(b_lo + ((k * 32) >> 4), kBHi), kIdesc, acc);
            } else if (MODE == SS_SAME_A) {   // every MMA re-reads the SAME A slice: does the hardware keep A?
              if (leader) umma_bf16(d, smem_desc_join(a_lo, kAHiAligned),
                                    smem_desc_join(b_lo + ((k * 32) >> 4), kBHi), kIdesc, acc);
            } else if (MODE == TS) {
              if (leader) umma_bf16_ts(d, tmem_a + k * 8, smem_desc_join(b_lo + ((k * 32) >> 4), kBHi), kIdesc, acc);
            } else if (MODE == CP_ONLY) {
              const uint32_t a_off = (uint32_t(r * 18 + t * 8 + s) * kRB + k * 32) >> 4;
              if (leader) utccp_128x256b(tmem_a + ((tap + t) & 1) * 32 + k * 8, smem_desc_join(a_lo + a_off, kAHiBox));
            } else {                           // CP_PLUS_TS: stage the tap-shifted A slice, then MMA from tensor memory
              const uint32_t a_off = (uint32_t(r * 18 + t * 8 + s) * kRB + k * 32) >> 4;
              const uint32_t ta = tmem_a + ((tap + t) & 1) * 32 + k * 8;
              if (leader) {
                utccp_128x256b(ta, smem_desc_join(a_lo + a_off, kAHiBox));
                umma_bf16_ts(d, ta, smem_desc_join(b_lo + ((k * 32) >> 4), kBHi), kIdesc, acc);
              }
            }
            ++count;
          }
        }
      }
    }
    if (leader) umma_commit(bar);
    __syncwarp();
    mbar_wait(bar, 0);
    t1 = clock64();
    tc_fence_after();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
    if (lane == 0 && blockIdx.x == 0) { *out_cycles = t1 - t0; *out_count = count; }
  }
}

template <int N, int MODE>
static void run(const char* name, int iters, int grid) {
  long long* d_cyc; int* d_cnt;
  cudaMalloc(&d_cyc, 8); cudaMalloc(&d_cnt, 4);
  const int smem = kABytes + kBBytes + 2048;
  cudaFuncSetAttribute(umma_bench<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  umma_bench<N, MODE><<<grid, 128, smem>>>(8, d_cyc, d_cnt);          // warm-up
  umma_bench<N, MODE><<<grid, 128, smem>>>(iters, d_cyc, d_cnt);
  cudaError_t e = cudaDeviceSynchronize();
  long long cyc = 0; int cnt = 0;
  cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(&cnt, d_cnt, 4, cudaMemcpyDeviceToHost);
  const double per = cnt ? (double)cyc / cnt : 0.0;
  const double floor_c = 128.0 * N / 256.0;
  printf("%-34s N=%3d grid=%3d  %8.1f cycles/op   floor %5.0f   tensor-pipe bound %5.1f %%   %s\n", name, N, grid, per,
         floor_c, MODE == CP_ONLY ? 0.0 : 100.0 * floor_c / per, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d_cyc); cudaFree(d_cnt);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  printf("%s, %d SMs, SM clock %d MHz; one op = one M128 x N x K16 bf16 MMA (or one 128x256b tcgen05.cp = 4 KB); "
         "clock64 around 72 ops x iters issued by one thread, completion by tcgen05.commit\n", p.name,
         p.multiProcessorCount, p.clockRate / 1000);
  const int it = 200;
  for (int grid : {1, 148}) {
    run<64, SS_ALIGNED>("SS, A 1024B-aligned tile", it, grid);
    run<64, SS_TAPSHIFT>("SS, A tap-shifted halo-box rows", it, grid);
    run<64, SS_SAME_A>("SS, same A slice every time", it, grid);
    run<64, TS>("TS, A resident in tensor memory", it, grid);
    run<64, CP_ONLY>("tcgen05.cp 128x256b only", it, grid);
    run<64, CP_PLUS_TS>("tcgen05.cp + TS MMA per slice", it, grid);
    run<128, SS_ALIGNED>("SS, A 1024B-aligned tile", it, grid);
    run<128, SS_TAPSHIFT>("SS, A tap-shifted halo-box rows", it, grid);
    run<128, TS>("TS, A resident in tensor memory", it, grid);
    run<128, CP_PLUS_TS>("tcgen05.cp + TS MMA per slice", it, grid);
    run<256, SS_ALIGNED>("SS, A 1024B-aligned tile", it, grid);
    run<256, TS>("TS, A resident in tensor memory", it, grid);
  }
  return 0;
}
