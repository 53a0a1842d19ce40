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

#ifdef __CUDACC__
// Exact uint32 -> double without the conversion unit: 2^52 + x is representable with x in the low
// mantissa word, and subtracting 2^52 is exact.  Measured on B200 (profiles/r02_lm_v2_raw.csv):
// I2F.F64.U32 runs on the XU pipe, which the two conversions per CSR entry of the sweep kept 57 %
// busy — as limiting as the L1TEX pipe; this form is one DADD on the (idle) FP64 pipe.
__device__ __forceinline__ double u32_to_f64(uint32_t x) {
  return __hiloint2double(0x43300000, static_cast<int>(x)) - 4503599627370496.0;
}
#endif

// Degree classes.  Boundaries are inclusive upper degrees.
enum : int {
  kClsZero = 0,  // deg == 0: never moves, no work
  kClsG4 = 1,    // 1..4      one thread / vertex (u32) or 4 lanes / vertex (f64)
  kClsG8 = 2,    // 5..8      8 lanes / vertex, single batch, no table
  kClsG16 = 3,   // 9..16     16 lanes
  kClsG32 = 4,   // 17..32    32 lanes
  kClsH128 = 5,  // 33..64    hash classes: table capacity = 2 x upper degree
  kClsH256 = 6,  // 65..128
  kClsH512 = 7,  // 129..256
  kClsH1K = 8,   // 257..512
  kClsH2K = 9,   // 513..1024
  kClsH4K = 10,  // 1025..2048
  kClsH8K = 11,  // 2049..4096
  kClsH16K = 12, // 4097..8192
  kClsC2 = 13,   // 8193..16384   thread-block cluster of 2 CTAs, table in distributed shared memory
  kClsC4 = 14,   // 16385..32768  cluster of 4
  kClsC8 = 15,   // 32769..65536  cluster of 8
  kClsGiant = 16,  // > 65536     tables in global memory
  kNumClasses = 17
};

__host__ __device__ inline int degree_class(int64_t deg) {
  if (deg == 0) return kClsZero;
  if (deg <= 4) return kClsG4;
  if (deg <= 8) return kClsG8;
  if (deg <= 16) return kClsG16;
  if (deg <= 32) return kClsG32;
  if (deg <= 64) return kClsH128;
  if (deg <= 128) return kClsH256;
  if (deg <= 256) return kClsH512;
  if (deg <= 512) return kClsH1K;
  if (deg <= 1024) return kClsH2K;
  if (deg <= 2048) return kClsH4K;
  if (deg <= 4096) return kClsH8K;
  if (deg <= 8192) return kClsH16K;
  if (deg <= 16384) return kClsC2;
  if (deg <= 32768) return kClsC4;
  if (deg <= 65536) return kClsC8;
  return kClsGiant;
}

}  // namespace cvb
