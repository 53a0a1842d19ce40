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
// nprocs == 1: one upload + cvb_phase_run.  nprocs > 1 (one MPI rank per GPU): the ranks create an
// NCCL communicator once (the 128-byte id travels over the host program's own MPI_Bcast) and the
// phase runs through cvb_dist_phase_run — label all-gather + local community rebuild + all-reduce of
// the two sums per sweep, all on the device (include/cuvite_b200.h, "N > 1 ranks").  The request
// lists distbuildNextLevelGraph consumes afterwards (what exchangeVertexReqs, louvain.cpp:3118-3264,
// leaves behind; it is `static` in louvain.hpp and cannot be called from here) are restated below.
// This image has no MPI: the nprocs > 1 branch is compiled against the single-rank shim only and has
// not been executed with real MPI ranks; the same cvb_dist_* calls are exercised on 2..8 GPUs through
// cuvite_b200/dist.py (tests/test_gpu_dist.py, bench.py --gpus N).
#include <algorithm>
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

// What exchangeVertexReqs (louvain.cpp:3118-3264) leaves in its output arguments: after its final swap,
// rvdata/rsizes = the ghost vertices THIS rank needs, grouped by owner; svdata/ssizes = the vertices of
// this rank the OTHER ranks need.  (The reference collects each group in a std::unordered_set; the
// order inside a group is unspecified there and ascending here.)
void resident_vertex_reqs(const DistGraph &dg, size_t &ssz, size_t &rsz, std::vector<GraphElem> &ssizes,
                          std::vector<GraphElem> &rsizes, std::vector<GraphElem> &svdata, std::vector<GraphElem> &rvdata,
                          const int me, const int nprocs) {
  const Graph &g = dg.getLocalGraph();
  std::vector<std::vector<GraphElem>> need(nprocs);
  for (const Edge &edge : g.edgeList) {
    const int owner = dg.getOwner(edge.tail);
    if (owner != me) need[owner].push_back(edge.tail);
  }
  std::vector<GraphElem> want_sizes(nprocs), asked_sizes(nprocs);
  std::vector<int> scnts(nprocs), rcnts(nprocs), sdispls(nprocs), rdispls(nprocs);
  std::vector<GraphElem> want;
  for (int p = 0; p < nprocs; ++p) {
    std::sort(need[p].begin(), need[p].end());
    need[p].erase(std::unique(need[p].begin(), need[p].end()), need[p].end());
    want_sizes[p] = static_cast<GraphElem>(need[p].size());
    sdispls[p] = static_cast<int>(want.size());
    scnts[p] = static_cast<int>(need[p].size());
    want.insert(want.end(), need[p].begin(), need[p].end());
  }
  MPI_Alltoall(want_sizes.data(), 1, MPI_GRAPH_TYPE, asked_sizes.data(), 1, MPI_GRAPH_TYPE, MPI_COMM_WORLD);
  size_t asked_total = 0;
  for (int p = 0; p < nprocs; ++p) {
    rdispls[p] = static_cast<int>(asked_total);
    rcnts[p] = static_cast<int>(asked_sizes[p]);
    asked_total += static_cast<size_t>(asked_sizes[p]);
  }
  std::vector<GraphElem> asked(asked_total);
  MPI_Alltoallv(want.data(), scnts.data(), sdispls.data(), MPI_GRAPH_TYPE, asked.data(), rcnts.data(), rdispls.data(),
                MPI_GRAPH_TYPE, MPI_COMM_WORLD);
  rvdata.swap(want);      // ghosts I need
  rsizes.swap(want_sizes);
  rsz = rvdata.size();
  svdata.swap(asked);     // my vertices the others need
  ssizes.swap(asked_sizes);
  ssz = svdata.size();
}
}  // namespace

GraphWeight distLouvainMethod(const int me, const int nprocs, const DistGraph &dg, size_t &ssz, size_t &rsz,
                              std::vector<GraphElem> &ssizes, std::vector<GraphElem> &rsizes,
                              std::vector<GraphElem> &svdata, std::vector<GraphElem> &rvdata, CommunityVector &cvect,
                              const GraphWeight lower, const GraphWeight thresh, int &iters, GpuGraph &gpu_graph) {
  static_assert(sizeof(Edge) == 16 && sizeof(GraphElem) == 8 && sizeof(GraphWeight) == 8,
                "cuvite_b200 implements the reference's 64-bit build (no -DUSE_32_BIT_GRAPH)");
  const Graph &g = dg.getLocalGraph();
  const GraphElem nv = g.getNumVertices();
  if (nprocs == 1) {
    // what exchangeVertexReqs (louvain.cpp:3118-3264) leaves on one rank: no partners, no ghosts;
    // distbuildNextLevelGraph consumes these lists (main.cpp:413-414)
    ssz = 0;
    rsz = 0;
    ssizes.assign(nprocs, 0);
    rsizes.assign(nprocs, 0);
    svdata.clear();
    rvdata.clear();
  } else {
    resident_vertex_reqs(dg, ssz, rsz, ssizes, rsizes, svdata, rvdata, me, nprocs);
  }
  cvect.resize(nv);
  iters = 0;
  if (nv <= 0 && nprocs == 1) return lower;
  cvb_engine *e = gpu_graph.engine();
  const bool timing = std::getenv("CVB_ADAPTER_TIMING") != nullptr;
  auto now = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t0 = now();
  if (cvb_graph_upload_host(e, nv, dg.getBase(me), dg.getTotalNumVertices(), g.edgeListIndexes.data(),
                            g.edgeList.data(), nullptr, nullptr) != CVB_OK)
    die_resident("distLouvainMethod/upload");
  const double t1 = now();
  double q = lower;
  if (nprocs == 1) {
    if (cvb_phase_run(e, lower, thresh, 0, &q, &iters) != CVB_OK) die_resident("distLouvainMethod/phase");
  } else {
    static bool comm_ready = false;  // one communicator per process, created at the first phase
    if (!comm_ready) {
      char id[128] = {0};
      if (me == 0 && cvb_dist_unique_id(id, sizeof id) != CVB_OK) die_resident("distLouvainMethod/nccl id");
      MPI_Bcast(id, sizeof id, MPI_BYTE, 0, MPI_COMM_WORLD);
      if (cvb_dist_init(e, id, sizeof id, me, nprocs) != CVB_OK) die_resident("distLouvainMethod/nccl init");
      comm_ready = true;
    }
    std::vector<int64_t> parts(nprocs + 1);
    for (int p = 0; p < nprocs; ++p) parts[p] = dg.getBase(p);
    parts[nprocs] = dg.getTotalNumVertices();
    if (cvb_dist_set_parts(e, parts.data()) != CVB_OK || cvb_dist_agree(e) != CVB_OK ||
        cvb_dist_phase_run(e, lower, thresh, 0, &q, &iters) != CVB_OK)
      die_resident("distLouvainMethod/dist phase");
  }
  const double t2 = now();
  if (cvb_get_labels(e, CVB_LABELS_PAST, cvect.data()) != CVB_OK) die_resident("distLouvainMethod/labels");
  if (timing)
    std::printf("cuvite_b200 resident phase: nv=%lld entries=%lld upload %.4f s, phase %.4f s (%d iterations), labels %.4f s\n",
                (long long)nv, (long long)g.getNumEdges(), t1 - t0, t2 - t1, iters, now() - t2);
  return q;
}
