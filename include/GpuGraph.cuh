// GpuGraph.cuh — drop-in replacement for pnnl/cuVite's GpuGraph.cuh.
//
// Put this directory BEFORE the reference tree on the include path
// (-I<cuvite_b200>/include -I<cuVite>): louvain.hpp:23 then picks this header,
// and the reference's main.cpp / louvain.cpp compile unchanged against it.
// Same namespace, class name, constructor and the three members main.cpp uses
// (GpuGraph.cuh:38-45 ctor, devMemAllocate(), pinnedAndDevMemAllocate(),
// getRefVecSize(); called main.cpp:209-213).  Everything else in the reference
// class was private plumbing of its own kernels (27 device arrays, ~50
// getters) and is gone: the device state lives behind the C ABI of
// cuvite_b200.h.  Host-only C++; no CUDA syntax, so g++/mpicxx can include it.
#ifndef __GPUGRAPH_CUH__
#define __GPUGRAPH_CUH__

#include <cstdio>
#include <cstdlib>

#include "edge.hpp"  // the reference's GraphElem / GraphWeight typedefs

#include "cuvite_b200.h"

namespace CuVite {

class GpuGraph {
 public:
  // nv/ne/ng = local vertices, local CSR entries, ghost count of PHASE 0
  // (main.cpp:203-209).  Later, smaller phases reuse the same object; buffers
  // are sized per uploaded graph, so no MODCURRCOMM_SIZE_FACTOR-style capacity
  // constant (GpuGraph.cuh:40-45) and no 20,971,520-vertex limit
  // (louvain_cuda.cu:2606) exist here.
  GpuGraph(GraphElem nv, GraphElem ne, GraphElem ng) : nv_(nv), ne_(ne), ng_(ng), engine_(nullptr) {}
  ~GpuGraph() {
    if (engine_) cvb_destroy(engine_);
  }
  GpuGraph(const GpuGraph &) = delete;
  GpuGraph &operator=(const GpuGraph &) = delete;

  // GpuGraph.cu:9-71: allocate the device side.  Fails like the reference's
  // CUDA_SAFE (GpuGraph.cuh:23): message + exit(1).
  void devMemAllocate() {
    if (!engine_) engine_ = cvb_create(nullptr);
    if (!engine_) {
      std::printf("CUDA CALL FAILED: %s\n", cvb_last_error());
      std::exit(1);
    }
  }
  // GpuGraph.cu:80-101: pinned staging is owned by the engine and grows on demand.
  void pinnedAndDevMemAllocate() { devMemAllocate(); }

  inline GraphElem getRefVecSize() { return nv_; }  // GpuGraph.cuh:67-72

  cvb_engine *engine() {
    devMemAllocate();
    return engine_;
  }

 private:
  GraphElem nv_, ne_, ng_;
  cvb_engine *engine_;
};

}  // namespace CuVite

#endif  // __GPUGRAPH_CUH__
