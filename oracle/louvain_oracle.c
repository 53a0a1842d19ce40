/* oracle/louvain_oracle.c — TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Plain-C restatement of the reference's CPU Louvain hot path (pnnl/cuVite,
 * default variant), one function per reference function, each citing the
 * reference lines it follows.  Pinned against the real reference compiled by
 * oracle/build_ref.sh (oracle/_ref/libcuvite_ref.so) in tests/test_oracle_pin.py
 * and against tests/golden/ on machines without /root/reference.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * load this.  64-bit ids, double weights (the reference's non-USE_32_BIT_GRAPH
 * build, edge.hpp:15-19).  Compiled with -ffp-contract=off so that the gain
 * expression is evaluated exactly as g++ evaluates louvain.cpp:2228 on x86-64.
 *
 * View: one address space holding rows [base, bound) of the CSR with *global*
 * tails, and global (replicated) label / community-info arrays.  With
 * base = 0, bound = nv this is the reference's one-rank execution; with a
 * sub-range it is what rank p of P computes, because the sweep is synchronous
 * (every vertex reads iteration-k labels and Cinfo, louvain.cpp:2246-2382).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef int64_t gelem;
typedef double gweight;

/* ---- per-thread open-addressing map standing in for ClusterLocalMap + counter
 * (louvain.cpp:2262-2263, 2384-2431).  Insertion order is kept so that
 * counter[] is summed in edge order exactly like the reference. ---- */
typedef struct {
  gelem *keys;      /* table: community id or -1 */
  int64_t *slot;    /* table: index into comm[]/w[] */
  gelem *comm;      /* insertion-ordered community ids  (clmap keys)   */
  gweight *w;       /* insertion-ordered weights        (counter[])    */
  int64_t cap, n, listcap;
} cmap_t;

static void cmap_reserve(cmap_t *m, int64_t deg) {
  int64_t need = 4;
  while (need < 2 * (deg + 1)) need <<= 1;
  if (need > m->cap) {
    free(m->keys); free(m->slot);
    m->keys = (gelem *)malloc(sizeof(gelem) * (size_t)need);
    m->slot = (int64_t *)malloc(sizeof(int64_t) * (size_t)need);
    m->cap = need;
    for (int64_t i = 0; i < need; i++) m->keys[i] = -1;
  }
  if (deg + 1 > m->listcap) {
    free(m->comm); free(m->w);
    m->listcap = deg + 1;
    m->comm = (gelem *)malloc(sizeof(gelem) * (size_t)m->listcap);
    m->w = (gweight *)malloc(sizeof(gweight) * (size_t)m->listcap);
  }
  m->n = 0;
}
static inline uint64_t cmap_hash(gelem k) {
  uint64_t z = (uint64_t)k * 0x9E3779B97F4A7C15ull;
  return z ^ (z >> 29);
}
static inline int64_t cmap_find_or_insert(cmap_t *m, gelem k, int *isnew) {
  uint64_t mask = (uint64_t)m->cap - 1, h = cmap_hash(k) & mask;
  while (m->keys[h] != -1 && m->keys[h] != k) h = (h + 1) & mask;
  if (m->keys[h] == k) { *isnew = 0; return m->slot[h]; }
  m->keys[h] = k; m->slot[h] = m->n; m->comm[m->n] = k; m->w[m->n] = 0.0;
  *isnew = 1;
  return m->n++;
}
static void cmap_clear(cmap_t *m) { /* undo only what was touched */
  uint64_t mask = (uint64_t)m->cap - 1;
  for (int64_t i = 0; i < m->n; i++) {
    uint64_t h = cmap_hash(m->comm[i]) & mask;
    while (m->keys[h] != m->comm[i]) h = (h + 1) & mask;
    m->keys[h] = -1;
  }
  /* linear probing leaves no tombstones because everything is removed */
  m->n = 0;
}
static void cmap_free(cmap_t *m) { free(m->keys); free(m->slot); free(m->comm); free(m->w); }

/* distSumVertexDegree, louvain.cpp:2126-2151: vDegree[i] = sum of stored edge
 * weights in edge order; Cinfo[i] = {size 1, degree vDegree[i]}. */
void orc_sum_vertex_degree(gelem nv, const gelem *offsets, const gweight *weights,
                           gweight *vdeg) {
#pragma omp parallel for schedule(guided)
  for (gelem i = 0; i < nv; i++) {
    gweight tw = 0.0;
    for (gelem k = offsets[i]; k < offsets[i + 1]; k++) tw += weights[k];
    vdeg[i] = tw;
  }
}

/* distCalcConstantForSecondTerm, louvain.cpp:2153-2183 (one rank: plain sum). */
gweight orc_total_weight(gelem nv, const gweight *vdeg) {
  gweight s = 0.0;
  for (gelem i = 0; i < nv; i++) s += vdeg[i];
  return s;
}

/* distGetMaxIndex, louvain.cpp:2185-2244.  Order-independent restatement of
 * the reference's loop over clmap: same comparisons, iterating in insertion
 * order instead of unordered_map order (the result does not depend on it:
 * a strictly larger gain always wins and ties go to the smaller id). */
static gelem orc_get_max_index(const cmap_t *m, gweight selfLoop, const gweight *cdeg,
                               const gelem *csize, gweight vDegree, gelem currSize,
                               gweight currDegree, gelem currComm, gweight constant) {
  gelem maxIndex = currComm;
  gweight curGain = 0.0, maxGain = 0.0;
  gweight eix = m->w[0] - selfLoop;          /* :2199 counter[0] is the own community */
  gweight ax = currDegree - vDegree;         /* :2201 */
  gelem maxSize = currSize;
  for (int64_t s = 0; s < m->n; s++) {
    gelem y = m->comm[s];
    if (y == currComm) continue;             /* :2212 */
    gweight ay = cdeg[y];                    /* :2215-2224 local or remote lookup */
    gelem size = csize[y];
    gweight eiy = m->w[s];
    curGain = 2.0 * (eiy - eix) - 2.0 * vDegree * (ay - ax) * constant; /* :2228 */
    if ((curGain > maxGain) ||
        ((curGain == maxGain) && (curGain != 0.0) && (y < maxIndex))) { /* :2230-2231 */
      maxGain = curGain; maxIndex = y; maxSize = size;
    }
  }
  if ((maxSize == 1) && (currSize == 1) && (maxIndex > currComm)) maxIndex = currComm; /* :2240 */
  return maxIndex;
}

/* One synchronous sweep over local rows [0, nvl) == global vertices
 * [base, base+nvl): distCleanCWandCU (louvain.cpp:2554-2567) +
 * distExecuteLouvainIteration (2246-2382) with distBuildLocalMapCounter
 * (2384-2431).  curr[] / cdeg[] / csize[] are global arrays (nvg entries);
 * dsize[] / ddeg[] (global, zeroed here) receive the Cupdate deltas of
 * 2307-2376; target[] and cw[] (clusterWeight) are local (nvl). */
void orc_sweep(gelem nvl, gelem base, gelem nvg, const gelem *offsets, const gelem *tails,
               const gweight *weights, const gelem *curr, const gweight *vdeg,
               const gweight *cdeg, const gelem *csize, gweight constant,
               gelem *target, gweight *cw, gweight *ddeg, gelem *dsize) {
  memset(ddeg, 0, sizeof(gweight) * (size_t)nvg);
  memset(dsize, 0, sizeof(gelem) * (size_t)nvg);
#pragma omp parallel
  {
    cmap_t m; memset(&m, 0, sizeof m);
#pragma omp for schedule(guided)
    for (gelem i = 0; i < nvl; i++) {
      const gelem e0 = offsets[i], e1 = offsets[i + 1];
      const gelem cc = curr[i + base];
      gelem localTarget;
      cw[i] = 0.0;                                   /* distCleanCWandCU */
      if (e0 != e1) {                                /* :2286 */
        int isnew;
        gweight selfLoop = 0.0;
        cmap_reserve(&m, e1 - e0);
        cmap_find_or_insert(&m, cc, &isnew);         /* :2289-2290 slot 0 = cc, weight 0 */
        for (gelem j = e0; j < e1; j++) {            /* distBuildLocalMapCounter :2396-2428 */
          const gelem tail = tails[j];
          const gweight w = weights[j];
          if (tail == i + base) selfLoop += w;       /* :2403-2404 */
          const gelem tcomm = curr[tail];            /* :2408-2417 */
          int64_t s = cmap_find_or_insert(&m, tcomm, &isnew);
          m.w[s] += w;                               /* :2421-2427 (new entry starts at 0 + w) */
        }
        cw[i] += m.w[0];                             /* :2294 */
        localTarget = orc_get_max_index(&m, selfLoop, cdeg, csize, vdeg[i], csize[cc],
                                        cdeg[cc], cc, constant);
        cmap_clear(&m);
      } else {
        localTarget = cc;                            /* :2299-2300 */
      }
      if (localTarget != cc) {                       /* :2307-2376, all four cases */
#pragma omp atomic update
        ddeg[localTarget] += vdeg[i];
#pragma omp atomic update
        dsize[localTarget] += 1;
#pragma omp atomic update
        ddeg[cc] -= vdeg[i];
#pragma omp atomic update
        dsize[cc] -= 1;
      }
      target[i] = localTarget;                       /* :2381 */
    }
    cmap_free(&m);
  }
}

/* distUpdateLocalCinfo, louvain.cpp:2539-2552 (+ updateRemoteCommunities'
 * apply loop 3101-3111 when the deltas came from other partitions). */
void orc_update_cinfo(gelem nvg, gweight *cdeg, gelem *csize, const gweight *ddeg,
                      const gelem *dsize) {
  for (gelem c = 0; c < nvg; c++) { csize[c] += dsize[c]; cdeg[c] += ddeg[c]; }
}

/* distComputeModularity, louvain.cpp:2433-2481: the two sums (sequential
 * order) and currMod = e*c - a2*c*c (2472-2473). */
gweight orc_modularity(gelem nv, const gweight *cw, const gweight *cdeg, gweight constant,
                       gweight *sums) {
  gweight le_xx = 0.0, la2_x = 0.0;
  for (gelem i = 0; i < nv; i++) { le_xx += cw[i]; la2_x += cdeg[i] * cdeg[i]; }
  if (sums) { sums[0] = le_xx; sums[1] = la2_x; }
  return (le_xx * constant) - (la2_x * constant * constant);
}

/* One phase: distLouvainMethod plain variant, louvain.cpp:425-588.
 * Optional per-iteration dumps ([max_trace][nv] row-major, any may be NULL).
 * Returns the accepted modularity (prevMod, :587); *iters = numIters (:577);
 * cvect = pastComm (:576). */
gweight orc_phase(gelem nv, const gelem *offsets, const gelem *tails, const gweight *weights,
                  gweight lower, gweight thresh, gelem *cvect, int *iters, int max_trace,
                  gelem *target_out, gweight *cw_out, gweight *q_out, gelem *csize_out,
                  gweight *cdeg_out) {
  gelem *past = (gelem *)malloc(sizeof(gelem) * (size_t)nv);
  gelem *curr = (gelem *)malloc(sizeof(gelem) * (size_t)nv);
  gelem *target = (gelem *)malloc(sizeof(gelem) * (size_t)nv);
  gelem *csize = (gelem *)malloc(sizeof(gelem) * (size_t)nv);
  gelem *dsize = (gelem *)malloc(sizeof(gelem) * (size_t)nv);
  gweight *vdeg = (gweight *)malloc(sizeof(gweight) * (size_t)nv);
  gweight *cw = (gweight *)malloc(sizeof(gweight) * (size_t)nv);
  gweight *cdeg = (gweight *)malloc(sizeof(gweight) * (size_t)nv);
  gweight *ddeg = (gweight *)malloc(sizeof(gweight) * (size_t)nv);
  /* distInitLouvain :2103-2124 */
  orc_sum_vertex_degree(nv, offsets, weights, vdeg);
  for (gelem i = 0; i < nv; i++) { cdeg[i] = vdeg[i]; csize[i] = 1; past[i] = i; curr[i] = i; }
  const gweight constant = 1.0 / orc_total_weight(nv, vdeg);
  gweight prevMod = lower, currMod = -1.0;
  int numIters = 0;
  while (1) {
    numIters++;
    orc_sweep(nv, 0, nv, offsets, tails, weights, curr, vdeg, cdeg, csize, constant, target, cw,
              ddeg, dsize);
    orc_update_cinfo(nv, cdeg, csize, ddeg, dsize);                 /* :519-522 */
    currMod = orc_modularity(nv, cw, cdeg, constant, 0);            /* :534 */
    if (numIters <= max_trace) {
      const size_t off = (size_t)(numIters - 1) * (size_t)nv;
      if (target_out) memcpy(target_out + off, target, sizeof(gelem) * (size_t)nv);
      if (cw_out) memcpy(cw_out + off, cw, sizeof(gweight) * (size_t)nv);
      if (csize_out) memcpy(csize_out + off, csize, sizeof(gelem) * (size_t)nv);
      if (cdeg_out) memcpy(cdeg_out + off, cdeg, sizeof(gweight) * (size_t)nv);
      if (q_out) q_out[numIters - 1] = currMod;
    }
    if ((currMod - prevMod) < thresh) break;                        /* :541 */
    prevMod = currMod;                                              /* :548 */
    if (prevMod < lower) prevMod = lower;                           /* :550-551 */
    gelem *tmp = past; past = curr; curr = target; target = tmp;    /* :562-567 */
  }
  memcpy(cvect, past, sizeof(gelem) * (size_t)nv);                  /* :576 */
  if (iters) *iters = numIters;
  free(past); free(curr); free(target); free(csize); free(dsize);
  free(vdeg); free(cw); free(cdeg); free(ddeg);
  return prevMod;
}

/* ---- coarsening, one-rank path of distbuildNextLevelGraph (rebuild.cpp:430-454) ---- */

/* distReNumber, rebuild.cpp:72-85 on one rank: communities are numbered in
 * order of first appearance of cvect[i], i ascending.  newid[] is indexed by
 * old community id (size nv, -1 where dead).  Returns the number of
 * communities (= vertices of the next level). */
gelem orc_renumber(gelem nv, const gelem *cvect, gelem *newid) {
  gelem lnc = 0;
  for (gelem i = 0; i < nv; i++) newid[i] = -1;
  for (gelem i = 0; i < nv; i++)
    if (newid[cvect[i]] < 0) newid[cvect[i]] = lnc++;
  return lnc;
}

typedef struct { gelem s, t; gweight w; int64_t seq; } orc_edge_t;
static int orc_edge_cmp(const void *a, const void *b) {
  const orc_edge_t *x = (const orc_edge_t *)a, *y = (const orc_edge_t *)b;
  if (x->s != y->s) return x->s < y->s ? -1 : 1;
  if (x->t != y->t) return x->t < y->t ? -1 : 1;
  return x->seq < y->seq ? -1 : (x->seq > y->seq);
}

/* fill_newEdgesMap (rebuild.cpp:244-279) + send_newEdges' CSR emit
 * (rebuild.cpp:405-427): edge (i -> tail, w) becomes (newid[cvect[i]] ->
 * newid[cvect[tail]], w); parallel edges are summed in the order the
 * reference meets them (i ascending, j ascending — kept by the seq key so
 * non-integer weights sum identically); each adjacency is emitted sorted by
 * tail (std::map).  Caller passes out arrays of capacity nnv+1 and ne.
 * Returns the number of CSR entries of the new graph. */
gelem orc_rebuild(gelem nv, const gelem *offsets, const gelem *tails, const gweight *weights,
                  const gelem *cvect, const gelem *newid, gelem nnv, gelem *noffsets,
                  gelem *ntails, gweight *nweights) {
  const gelem ne = offsets[nv];
  orc_edge_t *ed = (orc_edge_t *)malloc(sizeof(orc_edge_t) * (size_t)(ne > 0 ? ne : 1));
  for (gelem i = 0; i < nv; i++)
    for (gelem j = offsets[i]; j < offsets[i + 1]; j++) {
      ed[j].s = newid[cvect[i]];
      ed[j].t = newid[cvect[tails[j]]];
      ed[j].w = weights[j];
      ed[j].seq = j;
    }
  qsort(ed, (size_t)ne, sizeof(orc_edge_t), orc_edge_cmp);
  gelem out = 0;
  for (gelem v = 0; v <= nnv; v++) noffsets[v] = 0;
  for (gelem k = 0; k < ne;) {
    gelem k2 = k;
    gweight w = 0.0;
    while (k2 < ne && ed[k2].s == ed[k].s && ed[k2].t == ed[k].t) { w += ed[k2].w; k2++; }
    ntails[out] = ed[k].t; nweights[out] = w; noffsets[ed[k].s + 1]++; out++;
    k = k2;
  }
  for (gelem v = 0; v < nnv; v++) noffsets[v + 1] += noffsets[v];
  free(ed);
  return out;
}

#ifdef _OPENMP
int orc_num_threads(void) { return omp_get_max_threads(); }
void orc_set_num_threads(int n) { omp_set_num_threads(n); }
#else
int orc_num_threads(void) { return 1; }
void orc_set_num_threads(int n) { (void)n; }
#endif
