// Host-side helpers shared by the C-ABI launchers: status codes, launch checks and
// TMA tensor-map encoding through the driver entry point (resolved at run time so
// the library has no link-time dependency on libcuda and still loads on a box
// without a GPU, where only symbol presence is checked).
#pragma once
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/floodplanet_b200.h"

namespace fp {

// Per-device caches: one process may drive several devices (a model on cuda:1 while cuda:0 is
// current elsewhere), so nothing device-dependent is kept in a single process-wide static.
constexpr int kMaxDevices = 64;

inline int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return 0;
  return dev;
}

inline int sm_count() {
  static int n[kMaxDevices] = {0};
  const int dev = current_device();
  if (n[dev] == 0) {
    if (cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n[dev] = 148;
  }
  return n[dev];
}

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    fprintf(stderr, "[floodplanet_b200] %s: %s\n", what, cudaGetErrorString(e));
    return FPB200_ERR_LAUNCH;
  }
  return FPB200_OK;
}

inline PFN_cuTensorMapEncodeTiled_v12000 tmap_encoder() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

inline CUtensorMapSwizzle swizzle_for_bytes(int inner_bytes) {
  return inner_bytes >= 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                            : (inner_bytes >= 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                 : CU_TENSOR_MAP_SWIZZLE_32B);
}

// bf16 NHWC activation view [N][H][W][C] with a pixel pitch of `ld` elements
// (the view may be a channel slice of a wider concat buffer).  Box = (box_c, box_w, box_h, 1).
// Out-of-bounds box elements (the conv halo) are zero-filled by the TMA unit.
inline int make_tmap_act(CUtensorMap* m, const void* base, int N, int H, int W, int C, long ld,
                         int box_c, int box_w, int box_h) {
  auto enc = tmap_encoder();
  if (!enc) return FPB200_ERR_DRIVER;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * 2 * W,
                           (cuuint64_t)ld * 2 * W * H};
  cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for_bytes(box_c * 2),
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "[floodplanet_b200] cuTensorMapEncodeTiled(act) failed: %d (C=%d W=%d H=%d N=%d ld=%ld box=%d,%d,%d)\n",
            (int)r, C, W, H, N, ld, box_c, box_w, box_h);
    return FPB200_ERR_TENSORMAP;
  }
  return FPB200_OK;
}

// bf16 row-major matrix [rows][cols] (cols contiguous).  Box = (box_cols, box_rows).
inline int make_tmap_mat(CUtensorMap* m, const void* base, long rows, long cols, int box_cols,
                         int box_rows) {
  auto enc = tmap_encoder();
  if (!enc) return FPB200_ERR_DRIVER;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for_bytes(box_cols * 2),
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "[floodplanet_b200] cuTensorMapEncodeTiled(mat) failed: %d (rows=%ld cols=%ld box=%d,%d)\n",
            (int)r, rows, cols, box_cols, box_rows);
    return FPB200_ERR_TENSORMAP;
  }
  return FPB200_OK;
}

}  // namespace fp
