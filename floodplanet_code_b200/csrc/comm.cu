// Data-parallel exchange step over an ncclComm_t (C ABI: fpb200_nccl_* / fpb200_allreduce_f32).
//
// The reference is single-GPU (st_water_seg/fit.py:86-88); this is the glue the drop-in needs
// for BASELINE.json configs[2]: one gradient all-reduce per training step, launched bucket by
// bucket from the backward pass on a communication stream (parallel.py).  NCCL is bound at run
// time with dlopen/dlsym so that the library neither links libnccl nor cares which copy
// (torch's bundled 2.28 or the system 2.27) the process already holds; only the TYPES come from
// <nccl.h> at build time.  ncclConfig_t carries its own size/version, so a struct of the
// header's version is accepted by a newer runtime.
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include "host_common.h"

namespace fp {
namespace {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRankConfig)(ncclComm_t*, int, ncclUniqueId, int, ncclConfig_t*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

const NcclApi& nccl() {
  static NcclApi api = [] {
    NcclApi a;
    // the copy already mapped into the process wins (torch loads its bundled libnccl.so.2)
    a.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!a.handle) a.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!a.handle) a.handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!a.handle) return a;
#define FP_SYM(field, name) a.field = reinterpret_cast<decltype(a.field)>(dlsym(a.handle, name))
    FP_SYM(GetVersion, "ncclGetVersion");
    FP_SYM(GetUniqueId, "ncclGetUniqueId");
    FP_SYM(CommInitRankConfig, "ncclCommInitRankConfig");
    FP_SYM(CommInitRank, "ncclCommInitRank");
    FP_SYM(CommDestroy, "ncclCommDestroy");
    FP_SYM(AllReduce, "ncclAllReduce");
    FP_SYM(GetErrorString, "ncclGetErrorString");
#undef FP_SYM
    a.ok = a.GetVersion && a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllReduce;
    return a;
  }();
  return api;
}

int check_nccl(ncclResult_t r, const char* what) {
  if (r == ncclSuccess) return FPB200_OK;
  const NcclApi& a = nccl();
  fprintf(stderr, "[floodplanet_b200] %s: NCCL error %d (%s)\n", what, (int)r,
          a.GetErrorString ? a.GetErrorString(r) : "?");
  return FPB200_ERR_NCCL;
}

}  // namespace
}  // namespace fp

using namespace fp;

extern "C" {

int fpb200_nccl_version(void) {
  const NcclApi& a = nccl();
  if (!a.ok) return 0;
  int v = 0;
  return a.GetVersion(&v) == ncclSuccess ? v : 0;
}

int fpb200_nccl_unique_id(void* id128) {
  const NcclApi& a = nccl();
  if (!a.ok || id128 == nullptr) return FPB200_ERR_NCCL;
  static_assert(sizeof(ncclUniqueId) == FPB200_NCCL_UNIQUE_ID_BYTES, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  int rc = check_nccl(a.GetUniqueId(&id), "ncclGetUniqueId");
  if (rc == FPB200_OK) memcpy(id128, &id, sizeof(id));
  return rc;
}

int fpb200_nccl_comm_create(void** comm, int world, int rank, const void* id128, int max_ctas) {
  const NcclApi& a = nccl();
  if (!a.ok || comm == nullptr || id128 == nullptr) return FPB200_ERR_NCCL;
  if (world < 1 || rank < 0 || rank >= world) return FPB200_ERR_SHAPE;
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t c = nullptr;
  ncclResult_t r;
  if (a.CommInitRankConfig != nullptr) {
    ncclConfig_t cfg = NCCL_CONFIG_INITIALIZER;
    if (max_ctas > 0) {
      cfg.maxCTAs = max_ctas;
      cfg.minCTAs = max_ctas < 4 ? max_ctas : 4;
    }
    r = a.CommInitRankConfig(&c, world, id, rank, &cfg);
  } else {
    r = a.CommInitRank(&c, world, id, rank);
  }
  int rc = check_nccl(r, "ncclCommInitRank");
  if (rc == FPB200_OK) *comm = c;
  return rc;
}

int fpb200_nccl_comm_destroy(void* comm) {
  const NcclApi& a = nccl();
  if (!a.ok || comm == nullptr) return FPB200_ERR_NCCL;
  return check_nccl(a.CommDestroy(static_cast<ncclComm_t>(comm)), "ncclCommDestroy");
}

int fpb200_allreduce_f32(void* comm, float* buf, long count, int op_avg, void* stream) {
  const NcclApi& a = nccl();
  if (!a.ok || comm == nullptr || buf == nullptr) return FPB200_ERR_NCCL;
  if (count < 0) return FPB200_ERR_SHAPE;
  if (count == 0) return FPB200_OK;
  return check_nccl(a.AllReduce(buf, buf, (size_t)count, ncclFloat32, op_avg ? ncclAvg : ncclSum,
                                static_cast<ncclComm_t>(comm), static_cast<cudaStream_t>(stream)),
                    "ncclAllReduce");
}

}  // extern "C"
