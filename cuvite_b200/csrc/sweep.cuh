// sweep.cuh — the local-move sweep: neighbour scan, per-community weight
// accumulation, modularity-gain argmax, community-delta commit.
//
// Semantics: distExecuteLouvainIteration / distBuildLocalMapCounter /
// distGetMaxIndex of the reference (louvain.cpp:2246-2382, 2384-2431,
// 2185-2244).  The sweep is synchronous: every vertex reads iteration-k labels
// and community info and writes target[] / deltas, so vertices are independent
// work items and the result does not depend on scheduling.
//
// Work decomposition by degree class (common.cuh):
//   * deg <= 32      G lanes per vertex (G = 4/8/16/32), ONE batch: duplicates
//                    are merged with match.any, no table at all;
//   * 33..512        one warp per vertex, batches of 32 entries merged into a
//                    warp-private open-addressing table in shared memory with a
//                    plain-store claim protocol (no atomics: lanes of one warp
//                    agree through __syncwarp, duplicates inside a batch are
//                    merged by match.any first);
//   * 513..8192      one block per vertex, block-shared table in shared memory
//                    (CAS claim + atomic add, after the same warp-level merge);
//   * > 8192         same, table in a global-memory scratch slab.
// Weights are accumulated in edge order inside a batch, so in CVB_WEIGHTS_F64
// mode the group/warp paths reproduce the reference's per-community sums
// bit for bit; CVB_WEIGHTS_U32 mode is exact by construction.
#pragma once

#include <climits>

#include "common.cuh"

namespace cvb {

struct SweepParams {
  const eid_t *__restrict__ offsets;  // [nv_local+1], rebased to 0
  const vid_t *__restrict__ tails;    // [ne] global vertex ids
  const void *__restrict__ weights;   // [ne] uint32 or double
  const vid_t *__restrict__ labels;   // [nv_global] iteration-k labels
  const double *__restrict__ vdeg;    // [nv_local]
  const double *__restrict__ cdeg;    // [nv_global] community degree
  const int32_t *__restrict__ csize;  // [nv_global] community size
  double constant;                    // 1 / (2m)
  vid_t base;                         // first global id of the local rows
  vid_t *__restrict__ target;         // [nv_local]
  double *__restrict__ cw;            // [nv_local] clusterWeight = counter[0]
  double *__restrict__ ddeg;          // [nv_global] += deltas
  int32_t *__restrict__ dsize;        // [nv_global] += deltas
};

// curGain = 2.0*(eiy - eix) - 2.0*vDegree*(ay - ax)*constant, louvain.cpp:2228,
// evaluated left to right with every rounding the host compiler performs and
// no FMA contraction.
__device__ __forceinline__ double modularity_gain(double eiy, double eix, double vdeg, double ay,
                                                  double ax, double constant) {
  const double t1 = __dmul_rn(2.0, __dsub_rn(eiy, eix));
  const double t2 = __dmul_rn(__dmul_rn(__dmul_rn(2.0, vdeg), __dsub_rn(ay, ax)), constant);
  return __dsub_rn(t1, t2);
}

// (gain, id) ordering of louvain.cpp:2230-2231: larger gain wins, equal gain
// goes to the smaller community id.  Candidates are only ever accepted with
// gain > 0 (maxGain starts at 0 and a tie at 0 is excluded by `curGain != 0`).
__device__ __forceinline__ bool better(double g2, vid_t id2, double g, vid_t id) {
  return (g2 > g) || (g2 == g && id2 < id);
}

__device__ __forceinline__ uint32_t hash_slot(vid_t key, int log2cap) {
  return (static_cast<uint32_t>(key) * 2654435761u) >> (32 - log2cap);
}

// Singleton rule (louvain.cpp:2240-2241), outputs (2294, 2381) and the Cupdate
// deltas (2307-2376; all four local/remote cases collapse to the same four
// adds on the replicated delta arrays).
__device__ __forceinline__ void commit_vertex(const SweepParams &p, vid_t v, vid_t cc, double best_gain,
                                              vid_t best_id, double vd, double counter0) {
  vid_t best = (best_gain > 0.0) ? best_id : cc;
  if (best != cc && best > cc) {
    if (p.csize[best] == 1 && p.csize[cc] == 1) best = cc;
  }
  p.target[v] = best;
  p.cw[v] = counter0;
  if (best != cc && p.ddeg != nullptr) {  // nullptr: labels-only exchange derives the deltas from labels
    atomicAdd(&p.ddeg[best], vd);
    atomicAdd(&p.dsize[best], 1);
    atomicAdd(&p.ddeg[cc], -vd);
    atomicAdd(&p.dsize[cc], -1);
  }
}

// Sum of `w` over the lanes named in `m` (the caller's match group), added in
// ascending lane order (= edge order) on top of `init`.  All lanes of `gmask`
// must call it; the trip count is the largest group in the calling set.
template <typename Acc>
__device__ __forceinline__ Acc group_sum_ordered(unsigned gmask, unsigned m, Acc w, Acc init, unsigned lane) {
  Acc s = init;
  unsigned mm = m;
  while (__any_sync(gmask, mm != 0u)) {
    const int src = mm ? (__ffs(mm) - 1) : static_cast<int>(lane);
    const Acc x = __shfl_sync(gmask, w, src);
    if (mm) {
      s += x;
      mm &= mm - 1u;
    }
  }
  return s;
}

// ---------------------------------------------------------------------------
// deg <= G: G lanes per vertex, one batch, no table.
// ---------------------------------------------------------------------------
template <int G, typename W>
__global__ void __launch_bounds__(256) k_sweep_group(SweepParams p, const vid_t *__restrict__ list, int n) {
  const unsigned lane = threadIdx.x & 31u;
  const int lig = static_cast<int>(lane) & (G - 1);
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~static_cast<unsigned>(G - 1)));
  const int64_t gid = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) / G;
  if (gid >= n) return;  // whole group leaves together
  const vid_t v = list[gid];
  const eid_t e0 = p.offsets[v];
  const int deg = static_cast<int>(p.offsets[v + 1] - e0);
  const vid_t self = v + p.base;
  const vid_t cc = p.labels[self];
  const bool valid = lig < deg;
  vid_t tail = -1;
  W w = 0;
  vid_t comm = -2 - lig;  // distinct negative keys keep idle lanes in singleton groups
  if (valid) {
    tail = p.tails[e0 + lig];
    w = static_cast<const W *>(p.weights)[e0 + lig];
    comm = p.labels[tail];
  }
  const unsigned m = __match_any_sync(gmask, comm);
  const bool leader = valid && (static_cast<unsigned>(__ffs(m) - 1) == lane);
  const W s = group_sum_ordered<W>(gmask, m, w, static_cast<W>(0), lane);

  // counter[0]: weight into the own community (self-loops included)
  const unsigned ccb = __ballot_sync(gmask, leader && comm == cc);
  const W c0 = __shfl_sync(gmask, s, ccb ? (__ffs(ccb) - 1) : static_cast<int>(lane));
  const double counter0 = ccb ? static_cast<double>(c0) : 0.0;
  // selfLoop (louvain.cpp:2403-2404), summed in edge order
  unsigned slb = __ballot_sync(gmask, valid && tail == self);
  double selfloop = 0.0;
  while (slb) {
    selfloop += static_cast<double>(__shfl_sync(gmask, w, __ffs(slb) - 1));
    slb &= slb - 1u;
  }
  const double vd = p.vdeg[v];
  const double eix = counter0 - selfloop;
  const double ax = p.cdeg[cc] - vd;
  double g = -1.0;
  vid_t id = INT_MAX;
  if (leader && comm != cc) {
    g = modularity_gain(static_cast<double>(s), eix, vd, p.cdeg[comm], ax, p.constant);
    id = comm;
  }
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) {
    const double g2 = __shfl_xor_sync(gmask, g, o);
    const vid_t id2 = __shfl_xor_sync(gmask, id, o);
    if (better(g2, id2, g, id)) {
      g = g2;
      id = id2;
    }
  }
  if (lig == 0) commit_vertex(p, v, cc, g, id, vd, counter0);
}

// ---------------------------------------------------------------------------
// 33..512: one warp per vertex, warp-private table, no atomics.
// ---------------------------------------------------------------------------
template <typename W, int CAP>
__global__ void __launch_bounds__(128) k_sweep_warp(SweepParams p, const vid_t *__restrict__ list, int n) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int kWarps = 4;
  const unsigned lane = threadIdx.x & 31u;
  const int wib = threadIdx.x >> 5;
  W *wts = reinterpret_cast<W *>(smem_raw) + static_cast<size_t>(wib) * CAP;
  vid_t *keys = reinterpret_cast<vid_t *>(smem_raw + sizeof(W) * CAP * kWarps) + static_cast<size_t>(wib) * CAP;
  const int64_t gid = static_cast<int64_t>(blockIdx.x) * kWarps + wib;
  if (gid >= n) return;
  const vid_t v = list[gid];
  const eid_t e0 = p.offsets[v], e1 = p.offsets[v + 1];
  const int deg = static_cast<int>(e1 - e0);
  const vid_t self = v + p.base;
  const vid_t cc = p.labels[self];
  // table sized to the vertex: next pow2 >= 2*deg, at most CAP
  int log2cap = 6;
  while ((1 << log2cap) < 2 * deg && (1 << log2cap) < CAP) ++log2cap;
  const int cap = 1 << log2cap;
  const unsigned mask = static_cast<unsigned>(cap - 1);
  for (int s = lane; s < cap; s += 32) keys[s] = kEmptyKey;
  __syncwarp();

  W selfacc = 0;
  for (eid_t eb = e0; eb < e1; eb += 32) {
    const eid_t e = eb + lane;
    const bool valid = e < e1;
    vid_t tail = -1;
    W w = 0;
    vid_t comm = -2 - static_cast<vid_t>(lane);
    if (valid) {
      tail = p.tails[e];
      w = static_cast<const W *>(p.weights)[e];
      comm = p.labels[tail];
    }
    const unsigned m = __match_any_sync(0xffffffffu, comm);
    const bool leader = valid && (static_cast<unsigned>(__ffs(m) - 1) == lane);
    // leaders (distinct keys) find or claim their slot with plain stores
    bool pending = leader;
    unsigned slot = hash_slot(comm, log2cap);
    W cur = 0;
    while (__any_sync(0xffffffffu, pending)) {
      bool wrote = false;
      if (pending && keys[slot] == kEmptyKey) {
        keys[slot] = comm;
        wrote = true;
      }
      __syncwarp();
      if (pending) {
        const vid_t k = keys[slot];
        if (k == comm) {
          cur = wrote ? static_cast<W>(0) : wts[slot];
          pending = false;
        } else {
          slot = (slot + 1u) & mask;
        }
      }
      __syncwarp();
    }
    const W s = group_sum_ordered<W>(0xffffffffu, m, w, leader ? cur : static_cast<W>(0), lane);
    if (leader) wts[slot] = s;
    if (valid && tail == self) selfacc += w;
    __syncwarp();
  }

  // counter[0]: probe for the own community (same address in every lane)
  double counter0 = 0.0;
  {
    unsigned slot = hash_slot(cc, log2cap);
    while (true) {
      const vid_t k = keys[slot];
      if (k == cc) {
        counter0 = static_cast<double>(wts[slot]);
        break;
      }
      if (k == kEmptyKey) break;
      slot = (slot + 1u) & mask;
    }
  }
  double selfloop = static_cast<double>(selfacc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) selfloop += __shfl_xor_sync(0xffffffffu, selfloop, o);
  const double vd = p.vdeg[v];
  const double eix = counter0 - selfloop;
  const double ax = p.cdeg[cc] - vd;
  double g = -1.0;
  vid_t id = INT_MAX;
  for (int s = lane; s < cap; s += 32) {
    const vid_t k = keys[s];
    if (k != kEmptyKey && k != cc) {
      const double g2 = modularity_gain(static_cast<double>(wts[s]), eix, vd, p.cdeg[k], ax, p.constant);
      if (better(g2, k, g, id)) {
        g = g2;
        id = k;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double g2 = __shfl_xor_sync(0xffffffffu, g, o);
    const vid_t id2 = __shfl_xor_sync(0xffffffffu, id, o);
    if (better(g2, id2, g, id)) {
      g = g2;
      id = id2;
    }
  }
  if (lane == 0) commit_vertex(p, v, cc, g, id, vd, counter0);
}

// ---------------------------------------------------------------------------
// > 512: one block per vertex, block-shared table (shared or global memory).
// ---------------------------------------------------------------------------
template <typename W>
__device__ __forceinline__ void table_add(W *addr, W v) {
  atomicAdd(addr, v);
}

// Table reads after the atomic phase.  A table in global memory was updated by
// L2 atomics, so its plain reads must not be served from a stale L1 line.
template <bool GLOBAL_TABLE, typename T>
__device__ __forceinline__ T table_ld(const T *addr) {
  if (GLOBAL_TABLE) return __ldcg(addr);
  return *addr;
}

template <typename W, int CAP, int THREADS, bool GLOBAL_TABLE>
__global__ void __launch_bounds__(THREADS) k_sweep_block(SweepParams p, const vid_t *__restrict__ list, int n,
                                                        unsigned char *__restrict__ scratch,
                                                        size_t scratch_stride, int scratch_log2cap) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double red_g[32];
  __shared__ vid_t red_id[32];
  __shared__ double red_self[32];
  constexpr int kWarps = THREADS / 32;
  const unsigned lane = threadIdx.x & 31u;
  const int wib = threadIdx.x >> 5;
  W *wts;
  vid_t *keys;
  if (GLOBAL_TABLE) {
    unsigned char *mine = scratch + scratch_stride * blockIdx.x;
    wts = reinterpret_cast<W *>(mine);
    keys = reinterpret_cast<vid_t *>(mine + (sizeof(W) << scratch_log2cap));
  } else {
    wts = reinterpret_cast<W *>(smem_raw);
    keys = reinterpret_cast<vid_t *>(smem_raw + sizeof(W) * CAP);
  }
  for (int64_t item = blockIdx.x; item < n; item += gridDim.x) {
    const vid_t v = list[item];
    const eid_t e0 = p.offsets[v], e1 = p.offsets[v + 1];
    const int64_t deg = e1 - e0;
    const vid_t self = v + p.base;
    const vid_t cc = p.labels[self];
    int log2cap = 10;
    const int maxlog = GLOBAL_TABLE ? scratch_log2cap : 31 - __clz(CAP);
    while ((int64_t(1) << log2cap) < 2 * deg && log2cap < maxlog) ++log2cap;
    const int cap = 1 << log2cap;
    const unsigned mask = static_cast<unsigned>(cap - 1);
    for (int s = threadIdx.x; s < cap; s += THREADS) {
      keys[s] = kEmptyKey;
      wts[s] = 0;
    }
    __syncthreads();

    W selfacc = 0;
    for (eid_t eb = e0 + static_cast<eid_t>(wib) * 32; eb < e1; eb += THREADS) {
      const eid_t e = eb + lane;
      const bool valid = e < e1;
      vid_t tail = -1;
      W w = 0;
      vid_t comm = -2 - static_cast<vid_t>(lane);
      if (valid) {
        tail = p.tails[e];
        w = static_cast<const W *>(p.weights)[e];
        comm = p.labels[tail];
      }
      const unsigned m = __match_any_sync(0xffffffffu, comm);
      const bool leader = valid && (static_cast<unsigned>(__ffs(m) - 1) == lane);
      const W s = group_sum_ordered<W>(0xffffffffu, m, w, static_cast<W>(0), lane);
      if (leader) {
        unsigned slot = hash_slot(comm, log2cap);
        while (true) {
          const vid_t k = *reinterpret_cast<volatile vid_t *>(&keys[slot]);
          if (k == comm) break;
          if (k == kEmptyKey) {
            const vid_t old = atomicCAS(&keys[slot], kEmptyKey, comm);
            if (old == kEmptyKey || old == comm) break;
          }
          slot = (slot + 1u) & mask;
        }
        table_add<W>(&wts[slot], s);
      }
      if (valid && tail == self) selfacc += w;
    }
    __syncthreads();

    double counter0 = 0.0;
    {
      unsigned slot = hash_slot(cc, log2cap);
      while (true) {
        const vid_t k = table_ld<GLOBAL_TABLE>(&keys[slot]);
        if (k == cc) {
          counter0 = static_cast<double>(table_ld<GLOBAL_TABLE>(&wts[slot]));
          break;
        }
        if (k == kEmptyKey) break;
        slot = (slot + 1u) & mask;
      }
    }
    double selfloop = static_cast<double>(selfacc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) selfloop += __shfl_xor_sync(0xffffffffu, selfloop, o);
    if (lane == 0) red_self[wib] = selfloop;
    __syncthreads();
    selfloop = 0.0;
    for (int i = 0; i < kWarps; ++i) selfloop += red_self[i];

    const double vd = p.vdeg[v];
    const double eix = counter0 - selfloop;
    const double ax = p.cdeg[cc] - vd;
    double g = -1.0;
    vid_t id = INT_MAX;
    for (int s = threadIdx.x; s < cap; s += THREADS) {
      const vid_t k = table_ld<GLOBAL_TABLE>(&keys[s]);
      if (k != kEmptyKey && k != cc) {
        const double g2 = modularity_gain(static_cast<double>(table_ld<GLOBAL_TABLE>(&wts[s])), eix, vd,
                                          p.cdeg[k], ax, p.constant);
        if (better(g2, k, g, id)) {
          g = g2;
          id = k;
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double g2 = __shfl_xor_sync(0xffffffffu, g, o);
      const vid_t id2 = __shfl_xor_sync(0xffffffffu, id, o);
      if (better(g2, id2, g, id)) {
        g = g2;
        id = id2;
      }
    }
    if (lane == 0) {
      red_g[wib] = g;
      red_id[wib] = id;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int i = 1; i < kWarps; ++i)
        if (better(red_g[i], red_id[i], g, id)) {
          g = red_g[i];
          id = red_id[i];
        }
      commit_vertex(p, v, cc, g, id, vd, counter0);
    }
    __syncthreads();  // table and red_* are reused by the next item
  }
}

}  // namespace cvb
