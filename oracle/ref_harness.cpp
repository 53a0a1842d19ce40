// oracle/ref_harness.cpp — TEST INFRASTRUCTURE ONLY.
//
// C entry points around the *real* reference CPU implementation (pnnl/cuVite
// louvain.cpp / rebuild.cpp / distgraph.cpp), compiled by oracle/build_ref.sh
// from the sources where they lie under /root/reference into
// oracle/_ref/libcuvite_ref.so.  Nothing from the reference is copied into this
// repository: this file only *includes* the reference translation unit at
// build time (to reach its file-static helpers) and calls its functions.
//
// Used by tests/ (to pin oracle/louvain_oracle.c and to generate
// tests/golden/*), and by bench.py's `--impl reference` / cpu_baseline legs.
// The product (cuvite_b200/, include/) never links or loads this.
//
// Build flags (see build_ref.sh): -DDONT_CREATE_DIAG_FILES -ffp-contract=off,
// 64-bit GraphElem / double GraphWeight (no -DUSE_32_BIT_GRAPH), no -DNDEBUG.

#include "louvain.cpp"  // reference TU (patched copy: default(none)->default(shared))
#include "rebuild.hpp"

#include <chrono>

std::ofstream ofs;  // main.cpp:29 defines this in the reference program
std::ofstream ofcks;

namespace {

struct RefGraph {
  DistGraph *dg = nullptr;
  size_t ssz = 0, rsz = 0;
  std::vector<GraphElem> ssizes, rsizes, svdata, rvdata;
};

bool g_types_ready = false;
void ensure_types() {
  if (!g_types_ready) {
    createCommunityMPIType();  // louvain.cpp:2961
    createEdgeMPIType();       // rebuild.cpp:5
    g_types_ready = true;
  }
}

double now_s() {
  using namespace std::chrono;
  return duration<double>(steady_clock::now().time_since_epoch()).count();
}

}  // namespace

extern "C" {

// Build a one-rank DistGraph from CSR arrays (Vite layout: graph.hpp:35-36).
void *ref_graph_create(int64_t nv, int64_t ne, const int64_t *offsets,
                       const int64_t *tails, const double *weights) {
  ensure_types();
  RefGraph *h = new RefGraph;
  h->dg = new DistGraph(nv, ne);
  PartRanges parts(2);
  parts[0] = 0;
  parts[1] = nv;
  h->dg->createLocalGraph(nv, ne, &parts);
  Graph &g = h->dg->getLocalGraph();
  for (int64_t i = 0; i <= nv; i++) g.edgeListIndexes[i] = offsets[i];
  for (int64_t e = 0; e < ne; e++) {
    g.edgeList[e].tail = tails[e];
    g.edgeList[e].weight = weights[e];
  }
  return h;
}

// Load through the reference's own reader (distgraph.cpp:69-203).
void *ref_graph_load(const char *path) {
  ensure_types();
  RefGraph *h = new RefGraph;
  std::string fn(path);
  loadDistGraphMPIIO(0, 1, 1, h->dg, fn);
  return h;
}

// The reference's edge-balanced vertex ranges (-b): balanceEdges, distgraph.cpp:22-66.
// mbins_out[nprocs + 1].
void ref_balance_edges(int nprocs, const char *path, int64_t *mbins_out) {
  std::string fn(path);
  std::vector<GraphElem> mbins(nprocs + 1, 0);
  balanceEdges(nprocs, fn, mbins);
  for (int i = 0; i <= nprocs; i++) mbins_out[i] = mbins[i];
}

void ref_graph_destroy(void *hp) {
  RefGraph *h = static_cast<RefGraph *>(hp);
  delete h->dg;
  delete h;
}

int64_t ref_graph_nv(void *hp) { return static_cast<RefGraph *>(hp)->dg->getLocalGraph().getNumVertices(); }
int64_t ref_graph_ne(void *hp) { return static_cast<RefGraph *>(hp)->dg->getLocalGraph().getNumEdges(); }

void ref_graph_get(void *hp, int64_t *offsets, int64_t *tails, double *weights) {
  const Graph &g = static_cast<RefGraph *>(hp)->dg->getLocalGraph();
  const int64_t nv = g.getNumVertices(), ne = g.getNumEdges();
  for (int64_t i = 0; i <= nv; i++) offsets[i] = g.edgeListIndexes[i];
  for (int64_t e = 0; e < ne; e++) {
    tails[e] = g.edgeList[e].tail;
    weights[e] = g.edgeList[e].weight;
  }
}

// One phase through the reference's own plain-CPU distLouvainMethod
// (louvain.cpp:425-588).  Returns the accepted modularity; fills cvect[nv],
// *iters and *seconds (wall time of the call).
double ref_louvain_phase(void *hp, double lower, double thresh, int64_t *cvect_out,
                         int *iters_out, double *seconds_out) {
  RefGraph *h = static_cast<RefGraph *>(hp);
  CommunityVector cvect;
  int iters = 0;
  const double t0 = now_s();
  const GraphWeight q = distLouvainMethod(0, 1, *h->dg, h->ssz, h->rsz, h->ssizes, h->rsizes,
                                          h->svdata, h->rvdata, cvect, lower, thresh, iters);
  const double t1 = now_s();
  if (cvect_out) std::copy(cvect.begin(), cvect.end(), cvect_out);
  if (iters_out) *iters_out = iters;
  if (seconds_out) *seconds_out = t1 - t0;
  return q;
}

// The same phase loop, but stepped from here with the reference's own
// file-static per-iteration functions so every iteration's outputs can be
// dumped: targetComm, clusterWeight, the modularity the reference would print,
// and Cinfo after the update.  Mirrors louvain.cpp:449-587 statement by
// statement (one rank => fill/updateRemoteCommunities are no-ops but are
// still called).  Arrays are [max_iters][nv] row-major; any may be NULL.
// Returns the number of iterations executed (numIters at break).
int ref_phase_trace(void *hp, double lower, double thresh, int max_iters,
                    int64_t *target_out, double *cw_out, double *q_out,
                    int64_t *csize_out, double *cdeg_out, int64_t *cvect_out,
                    double *qret_out) {
  RefGraph *h = static_cast<RefGraph *>(hp);
  const DistGraph &dg = *h->dg;
  const int me = 0, nprocs = 1;
  CommunityVector pastComm, currComm, targetComm;
  GraphWeightVector vDegree, clusterWeight;
  CommVector localCinfo, localCupdate;
  VertexCommMap remoteComm;
  CommMap remoteCinfo, remoteCupdate;
  const Graph &g = dg.getLocalGraph();
  const GraphElem nv = g.getNumVertices();
  GraphWeight constantForSecondTerm;
  GraphWeight prevMod = lower, currMod = -1.0;
  int numIters = 0;

  distInitLouvain(dg, pastComm, currComm, vDegree, clusterWeight, localCinfo, localCupdate,
                  constantForSecondTerm, me);
  targetComm.resize(nv);
  exchangeVertexReqs(dg, h->ssz, h->rsz, h->ssizes, h->rsizes, h->svdata, h->rvdata, me, nprocs);

  while (true) {
    numIters++;
    fillRemoteCommunities(dg, me, nprocs, h->ssz, h->rsz, h->ssizes, h->rsizes, h->svdata,
                          h->rvdata, currComm, localCinfo, remoteCinfo, remoteComm, remoteCupdate);
#pragma omp parallel
    {
      distCleanCWandCU(nv, clusterWeight, localCupdate);
#pragma omp for schedule(guided)
      for (GraphElem i = 0; i < nv; i++)
        distExecuteLouvainIteration(i, dg, currComm, targetComm, vDegree, localCinfo, localCupdate,
                                    remoteComm, remoteCinfo, remoteCupdate, constantForSecondTerm,
                                    clusterWeight, me);
    }
#pragma omp parallel
    { distUpdateLocalCinfo(localCinfo, localCupdate); }
    updateRemoteCommunities(dg, localCinfo, remoteCupdate, me, nprocs);
    currMod = distComputeModularity(g, localCinfo, clusterWeight, constantForSecondTerm, me);

    if (numIters <= max_iters) {
      const size_t off = (size_t)(numIters - 1) * (size_t)nv;
      for (GraphElem i = 0; i < nv; i++) {
        if (target_out) target_out[off + i] = targetComm[i];
        if (cw_out) cw_out[off + i] = clusterWeight[i];
        if (csize_out) csize_out[off + i] = localCinfo[i].size;
        if (cdeg_out) cdeg_out[off + i] = localCinfo[i].degree;
      }
      if (q_out) q_out[numIters - 1] = currMod;
    }

    if ((currMod - prevMod) < thresh) break;
    prevMod = currMod;
    if (prevMod < lower) prevMod = lower;
    for (GraphElem i = 0; i < nv; i++) {
      GraphElem tmp = pastComm[i];
      pastComm[i] = currComm[i];
      currComm[i] = targetComm[i];
      targetComm[i] = tmp;
    }
  }
  if (cvect_out) std::copy(pastComm.begin(), pastComm.end(), cvect_out);
  if (qret_out) *qret_out = prevMod;
  return numIters;
}

// Coarsen in place with the reference's own rebuild (rebuild.cpp:430-454).
// Must follow a ref_louvain_phase/ref_phase_trace call on the same handle
// (it consumes the ghost lists exchangeVertexReqs left behind, main.cpp:413).
void ref_rebuild(void *hp, const int64_t *cvect, double *seconds_out) {
  RefGraph *h = static_cast<RefGraph *>(hp);
  const int64_t nv = h->dg->getLocalGraph().getNumVertices();
  CommunityVector cv(cvect, cvect + nv);
  const double t0 = now_s();
  distbuildNextLevelGraph(1, 0, h->dg, h->ssz, h->rsz, h->ssizes, h->rsizes, h->svdata,
                          h->rvdata, cv);
  if (seconds_out) *seconds_out = now_s() - t0;
}

// The default-flag phase driver of main.cpp:218-495 (no coloring, no early
// termination, no threshold cycling): threshold 1e-6 (main.cpp:239), lower =
// currMod = -1 (main.cpp:177,461), rebuild while the phase gained more than
// the threshold (main.cpp:361,413), stop at 200 phases / 10000 iterations
// (main.cpp:486).  level_q / level_iters (cumulative, main.cpp:450) /
// level_nv / level_secs hold up to max_levels accepted levels; comm_out[nv0]
// gets the *true* composition of the per-level assignments (each level's cvect
// pushed through rebuild's first-appearance numbering, rebuild.cpp:72-85).
// Returns the number of accepted levels.  The graph handle is consumed
// (left holding the last coarsened graph).
int ref_louvain_run(void *hp, int max_levels, double *level_q, int *level_iters,
                    int64_t *level_nv, double *level_secs, double *level_rebuild_secs,
                    int64_t *comm_out) {
  RefGraph *h = static_cast<RefGraph *>(hp);
  const int64_t nv0 = h->dg->getLocalGraph().getNumVertices();
  std::vector<int64_t> compose(nv0);
  for (int64_t i = 0; i < nv0; i++) compose[i] = i;
  GraphWeight currMod = -1.0, prevMod = -1.0;
  int phase = 0, iters = 0, tot_iters = 0;
  const GraphWeight threshold = 1.0E-6;
  CommunityVector cvect;
  while (1) {
    const int64_t nv = h->dg->getLocalGraph().getNumVertices();
    const double t1 = now_s();
    currMod = distLouvainMethod(0, 1, *h->dg, h->ssz, h->rsz, h->ssizes, h->rsizes, h->svdata,
                                h->rvdata, cvect, currMod, threshold, iters);
    const double t0 = now_s();
    tot_iters += iters;
    if ((currMod - prevMod) > threshold) {
      // first-appearance renumbering, identical to distReNumber on one rank
      std::unordered_map<int64_t, int64_t> lookUp;
      int64_t lnc = 0;
      for (int64_t i = 0; i < nv; i++)
        if (lookUp.find(cvect[i]) == lookUp.end()) lookUp.insert(std::make_pair(cvect[i], lnc++));
      for (int64_t p = 0; p < nv0; p++) compose[p] = lookUp[cvect[compose[p]]];
      const double t3 = now_s();
      distbuildNextLevelGraph(1, 0, h->dg, h->ssz, h->rsz, h->ssizes, h->rsizes, h->svdata,
                              h->rvdata, cvect);
      if (phase < max_levels && level_rebuild_secs) level_rebuild_secs[phase] = now_s() - t3;
    } else {
      break;
    }
    if (phase < max_levels) {
      if (level_q) level_q[phase] = currMod;
      if (level_iters) level_iters[phase] = tot_iters;
      if (level_nv) level_nv[phase] = nv;
      if (level_secs) level_secs[phase] = t0 - t1;
    }
    prevMod = currMod;
    currMod = -1.0;
    phase++;
    if (phase == 200 || tot_iters > 10000) break;
  }
  if (comm_out) std::copy(compose.begin(), compose.end(), comm_out);
  return phase;
}

int ref_num_threads() { return omp_get_max_threads(); }
void ref_set_num_threads(int n) { omp_set_num_threads(n); }

}  // extern "C"
