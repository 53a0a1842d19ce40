// adapter.cpp — the reference's own plugin symbols, implemented over the C ABI.
//
// Compiled WITH the reference headers on the include path (it needs
// louvain.hpp's container typedefs), in place of the reference's
// louvain_cuda.cu / modularity.cu / GpuGraph.cu; see INTEGRATION.md.  Defines
// exactly the three functions the host objects import:
//   gpu_for_louvain_iteration   louvain.hpp:232-245, louvain_cuda.cu:1671-1686
//   gpu_reduce_mod              louvain.hpp:247-248, modularity.cu:102-141
//   set_gpuDevices              louvain.hpp:249,     louvain_cuda.cu:1634-1669
// The STL containers are flattened into the plain arrays cvb_louvain_iteration
// takes; remote maps keep std::map (ascending id) order like the reference's own
// flatteners (louvain_cuda_cpp_interface.hpp:4-48).
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "louvain.hpp"  // reference header (brings GpuGraph.cuh from include/)

#include "cuvite_b200.h"

namespace {
[[noreturn]] void die(const char *what) {
  // the reference: printf("CUDA CALL FAILED AT %d") + exit(1), GpuGraph.cuh:23
  std::printf("CUDA CALL FAILED in %s: %s\n", what, cvb_last_error());
  std::exit(1);
}
}  // namespace

void set_gpuDevices(int *me) {
  // The reference derives the node-local rank with MPI_Comm_split_type
  // (louvain_cuda.cu:1641-1658); launchers export it, fall back to the world rank.
  int local = *me;
  for (const char *var : {"LOCAL_RANK", "OMPI_COMM_WORLD_LOCAL_RANK", "MV2_COMM_WORLD_LOCAL_RANK", "SLURM_LOCALID"})
    if (const char *s = std::getenv(var)) {
      local = std::atoi(s);
      break;
    }
  if (cvb_set_device(local) != CVB_OK) die("set_gpuDevices");
}

int gpu_for_louvain_iteration(const GraphElem nv, const DistGraph &dg, CommunityVector &currComm,
                              CommunityVector &targetComm, GraphWeightVector &vDegree, CommVector &localCinfo,
                              CommVector &localCupdate, VertexCommMap &remoteComm, const CommMap &remoteCinfo,
                              CommMap &remoteCupdate, const double constantForSecondTerm,
                              GraphWeightVector &clusterWeight, int me, int numIters, GpuGraph &gpu_graph) {
  if (nv <= 0) return 1;  // louvain_cuda.cu:1685
  const Graph &g = dg.getLocalGraph();
  const GraphElem base = dg.getBase(me);
  const GraphElem nvg = dg.getTotalNumVertices();
  static_assert(sizeof(Edge) == 16 && sizeof(GraphElem) == 8 && sizeof(GraphWeight) == 8,
                "cuvite_b200 implements the reference's 64-bit build (no -DUSE_32_BIT_GRAPH)");
  // localCinfo / localCupdate: AoS Comm{size, degree} -> two arrays
  std::vector<int64_t> csize(nv), usize(nv);
  std::vector<double> cdeg(nv), udeg(nv);
  for (GraphElem i = 0; i < nv; i++) {
    csize[i] = localCinfo[i].size;
    cdeg[i] = localCinfo[i].degree;
    usize[i] = localCupdate[i].size;
    udeg[i] = localCupdate[i].degree;
  }
  std::vector<int64_t> gv, gc, rid, rsz, rusz;
  std::vector<double> rdg, rudg;
  gv.reserve(remoteComm.size());
  gc.reserve(remoteComm.size());
  for (const auto &kv : remoteComm) {
    gv.push_back(kv.first);
    gc.push_back(kv.second);
  }
  for (const auto &kv : remoteCinfo) {  // std::map: ascending community id
    rid.push_back(kv.first);
    rsz.push_back(kv.second.size);
    rdg.push_back(kv.second.degree);
  }
  rusz.assign(rid.size(), 0);
  rudg.assign(rid.size(), 0.0);
  const int rc = cvb_louvain_iteration(
      gpu_graph.engine(), nv, base, nvg, g.edgeListIndexes.data(), g.edgeList.data(), currComm.data(), vDegree.data(),
      csize.data(), cdeg.data(), static_cast<int64_t>(gv.size()), gv.data(), gc.data(),
      static_cast<int64_t>(rid.size()), rid.data(), rsz.data(), rdg.data(), constantForSecondTerm, numIters,
      targetComm.data(), clusterWeight.data(), usize.data(), udeg.data(), rusz.data(), rudg.data());
  if (rc <= 0) die("gpu_for_louvain_iteration");
  for (GraphElem i = 0; i < nv; i++) {
    localCupdate[i].size = usize[i];
    localCupdate[i].degree = udeg[i];
  }
  size_t k = 0;
  for (auto &kv : remoteCupdate) {  // entries pre-created by fillRemoteCommunities (louvain.cpp:2953)
    // remoteCupdate has the same key set as remoteCinfo; both are std::map, same order
    while (k < rid.size() && rid[k] < kv.first) ++k;
    if (k < rid.size() && rid[k] == kv.first) {
      kv.second.size += rusz[k];
      kv.second.degree += rudg[k];
    }
  }
  return 1;
}

void gpu_reduce_mod(CommVector &localCinfo, const GraphWeightVector &clusterWeight, GpuGraph &gpu_graph,
                    GraphWeight le_la_xx[], const GraphElem nv) {
  std::vector<double> cdeg(nv > 0 ? nv : 0);
  for (GraphElem i = 0; i < nv; i++) cdeg[i] = localCinfo[i].degree;
  if (cvb_reduce_mod(gpu_graph.engine(), nv, cdeg.data(), clusterWeight.data(), le_la_xx) != CVB_OK)
    die("gpu_reduce_mod");
}
