// common.cuh — shared types, error plumbing and device helpers.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <string>

#include "../../include/cuvite_b200.h"

namespace cvb {

void set_error(const char *fmt, ...);

#define CVB_CUDA(call)                                                                      \
  do {                                                                                      \
    cudaError_t cvb_err_ = (call);                                                          \
    if (cvb_err_ != cudaSuccess) {                                                          \
      ::cvb::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,                 \
                       cudaGetErrorString(cvb_err_));                                       \
      return CVB_ERR_CUDA;                                                                  \
    }                                                                                       \
  } while (0)

#define CVB_CHECK(cond, status, ...)                                                        \
  do {                                                                                      \
    if (!(cond)) {                                                                          \
      ::cvb::set_error(__VA_ARGS__);                                                        \
      return (status);                                                                      \
    }                                                                                       \
  } while (0)

#define CVB_TRY(expr)                                                                       \
  do {                                                                                      \
    int cvb_rc_ = (expr);                                                                   \
    if (cvb_rc_ != CVB_OK) return cvb_rc_;                                                  \
  } while (0)

// Device id / offset types.  Vertex and community ids are int32 on the device
// (every BASELINE config has nv_global < 2^31); CSR offsets stay int64 so one
// GPU can hold more than 2^31 entries.
using vid_t = int32_t;
using eid_t = int64_t;

constexpr vid_t kEmptyKey = -1;

// Degree classes.  Boundaries are inclusive upper degrees.
enum : int {
  kClsZero = 0,   // deg == 0: never moves, no work
  kClsG4 = 1,     // 1..4     4 lanes / vertex, single batch, no table
  kClsG8 = 2,     // 5..8     8 lanes
  kClsG16 = 3,    // 9..16    16 lanes
  kClsG32 = 4,    // 17..32   32 lanes
  kClsW256 = 5,   // 33..128  warp / vertex, 256-slot warp-private table
  kClsW1K = 6,    // 129..512 warp / vertex, 1024-slot warp-private table
  kClsB4K = 7,    // 513..2048  block / vertex, 4096-slot shared table
  kClsB16K = 8,   // 2049..8192 block / vertex, 16384-slot shared table
  kClsGiant = 9,  // > 8192   block / vertex, table in global memory
  kNumClasses = 10
};

__host__ __device__ inline int degree_class(int64_t deg) {
  if (deg == 0) return kClsZero;
  if (deg <= 4) return kClsG4;
  if (deg <= 8) return kClsG8;
  if (deg <= 16) return kClsG16;
  if (deg <= 32) return kClsG32;
  if (deg <= 128) return kClsW256;
  if (deg <= 512) return kClsW1K;
  if (deg <= 2048) return kClsB4K;
  if (deg <= 8192) return kClsB16K;
  return kClsGiant;
}

}  // namespace cvb
