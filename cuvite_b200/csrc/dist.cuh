// dist.cuh — the multi-GPU phase driver behind the C ABI: one engine per GPU /
// process, the reference's contiguous vertex ranges, NCCL on the engine stream.
//
// Replaces, for N > 1 ranks, louvain.cpp:591-753 with its MPI ghost exchange
// (exchangeVertexReqs 3118-3264, fillRemoteCommunities 2588-2959,
// updateRemoteCommunities 2983-3116, the Allreduce of distComputeModularity
// 2466/2522 and of distCalcConstantForSecondTerm 2175):
//   per sweep  local kernels -> ncclAllGather(new labels, 4 B / vertex) ->
//              every rank derives the community deltas of ALL ranks' moves from
//              the label diff on its own GPU -> ncclAllReduce of the two sums.
// NCCL is resolved at run time (dlopen), so the library has no link-time
// dependency on it and shares whatever libnccl the host process already loaded
// (torch's bundled copy under Python, the system one under a C++ main).
#pragma once

#include <dlfcn.h>

#include <vector>

#include "common.cuh"

namespace cvb {

// the few NCCL declarations this file needs (ABI of nccl.h 2.x)
typedef struct ncclComm *ncclComm_t;
typedef struct {
  char internal[128];
} ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclInt8 = 0, ncclChar = 0, ncclUint8 = 1, ncclInt32 = 2, ncclInt = 2, ncclUint32 = 3, ncclInt64 = 4,
       ncclUint64 = 5, ncclFloat16 = 6, ncclFloat32 = 7, ncclFloat64 = 8 };
enum { ncclSum = 0, ncclProd = 1, ncclMax = 2, ncclMin = 3 };

struct NcclApi {
  void *lib = nullptr;
  int (*GetUniqueId)(ncclUniqueId *) = nullptr;
  int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Broadcast)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
};

inline NcclApi *nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api.lib ? &api : nullptr;
  tried = true;
  const char *names[] = {getenv("CVB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  for (const char *nm : names) {
    if (!nm) continue;
    api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (api.lib) break;
  }
  if (!api.lib) return nullptr;
  bool ok = true;
  auto sym = [&](const char *s) {
    void *p = dlsym(api.lib, s);
    if (!p) ok = false;
    return p;
  };
  api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
  api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
  api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
  api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
  api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
  api.Broadcast = reinterpret_cast<decltype(api.Broadcast)>(sym("ncclBroadcast"));
  api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
  api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
  api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
  api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
  api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
  if (!ok) {
    dlclose(api.lib);
    api.lib = nullptr;
    return nullptr;
  }
  return &api;
}

#define CVB_NCCL(call)                                                                              \
  do {                                                                                              \
    int cvb_nrc_ = (call);                                                                          \
    if (cvb_nrc_ != ncclSuccess) {                                                                  \
      ::cvb::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,                         \
                       ::cvb::nccl_api() ? ::cvb::nccl_api()->GetErrorString(cvb_nrc_) : "nccl?"); \
      return CVB_ERR_CUDA;                                                                          \
    }                                                                                               \
  } while (0)

struct DistState {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
  std::vector<int64_t> parts;  // [world+1] vertex ranges of the current graph
  bool even = true;
  double total_weight = 0.0;
  // CVB_DIST_TIMING=1: CUDA-event stamps around the stages of a sweep
  bool timing = false;
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  double stage_ms[5] = {0, 0, 0, 0, 0};
  int64_t stage_n = 0;
};

// u32 weights -> double (a rank that picked the exact-integer mode joins ranks that could not)
__global__ void k_weights_u32_to_f64(const uint32_t *__restrict__ src, double *__restrict__ dst, int64_t n) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    dst[i] = static_cast<double>(src[i]);
}

}  // namespace cvb
