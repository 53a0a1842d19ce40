// adapter_resident.cpp — the OUTER drop-in boundary (SURVEY.md 8b "outer-boundary option",
// INTEGRATION.md section 5): the GPU overload of distLouvainMethod
//     louvain.hpp:93-97 (declaration), louvain.cpp:591-753 (the reference's definition),
//     called from main.cpp:344-345
// implemented on the resident engine: the graph of the phase is uploaded ONCE, all sweeps of
// the phase run on the device (cvb_phase_run) and only cvect comes back — instead of the
// reference's per-iteration gpu_for_louvain_iteration call with ~56 bytes/vertex over PCIe
// per sweep.  Build:
//   main.cpp ... -DRUNONGPU=1 -I<repo>/include   (uses the GPU overload, constructs GpuGraph)
//   louvain.cpp  WITHOUT -DRUNONGPU              (so that it does not define the overload)
//   + adapter.cpp (set_gpuDevices) + this file + -lcuvite_b200
// One rank per process without MPI peers (nprocs == 1): the multi-GPU resident path is
// cuvite_b200/dist.py over NCCL; an MPI multi-rank run keeps the per-iteration plugin
// (adapter.cpp).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "louvain.hpp"  // reference header (brings GpuGraph.cuh from include/)

#include "cuvite_b200.h"

namespace {
[[noreturn]] void die_resident(const char *what) {
  std::printf("CUDA CALL FAILED in %s: %s\n", what, cvb_last_error());  // GpuGraph.cuh:23 convention
  std::exit(1);
}
}  // namespace

GraphWeight distLouvainMethod(const int me, const int nprocs, const DistGraph &dg, size_t &ssz, size_t &rsz,
                              std::vector<GraphElem> &ssizes, std::vector<GraphElem> &rsizes,
                              std::vector<GraphElem> &svdata, std::vector<GraphElem> &rvdata, CommunityVector &cvect,
                              const GraphWeight lower, const GraphWeight thresh, int &iters, GpuGraph &gpu_graph) {
  if (nprocs != 1) {
    std::printf("cuvite_b200 resident distLouvainMethod: %d MPI ranks — link adapter.cpp's per-iteration plugin "
                "for MPI runs, or use cuvite_b200/dist.py (NCCL) for the multi-GPU resident path\n", nprocs);
    std::exit(1);
  }
  static_assert(sizeof(Edge) == 16 && sizeof(GraphElem) == 8 && sizeof(GraphWeight) == 8,
                "cuvite_b200 implements the reference's 64-bit build (no -DUSE_32_BIT_GRAPH)");
  const Graph &g = dg.getLocalGraph();
  const GraphElem nv = g.getNumVertices();
  // what exchangeVertexReqs (louvain.cpp:3118-3264) leaves on one rank: no partners, no ghosts;
  // distbuildNextLevelGraph consumes these lists (main.cpp:413-414)
  ssz = 0;
  rsz = 0;
  ssizes.assign(nprocs, 0);
  rsizes.assign(nprocs, 0);
  svdata.clear();
  rvdata.clear();
  cvect.resize(nv);
  iters = 0;
  if (nv <= 0) return lower;
  cvb_engine *e = gpu_graph.engine();
  const bool timing = std::getenv("CVB_ADAPTER_TIMING") != nullptr;
  auto now = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t0 = now();
  if (cvb_graph_upload_host(e, nv, dg.getBase(me), dg.getTotalNumVertices(), g.edgeListIndexes.data(),
                            g.edgeList.data(), nullptr, nullptr) != CVB_OK)
    die_resident("distLouvainMethod/upload");
  const double t1 = now();
  double q = lower;
  if (cvb_phase_run(e, lower, thresh, 0, &q, &iters) != CVB_OK) die_resident("distLouvainMethod/phase");
  const double t2 = now();
  if (cvb_get_labels(e, CVB_LABELS_PAST, cvect.data()) != CVB_OK) die_resident("distLouvainMethod/labels");
  if (timing)
    std::printf("cuvite_b200 resident phase: nv=%lld entries=%lld upload %.4f s, phase %.4f s (%d iterations), labels %.4f s\n",
                (long long)nv, (long long)g.getNumEdges(), t1 - t0, t2 - t1, iters, now() - t2);
  return q;
}
