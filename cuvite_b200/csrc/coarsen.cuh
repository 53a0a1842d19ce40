// coarsen.cuh — next-level graph construction on the device: the one-rank path
// of distbuildNextLevelGraph (rebuild.cpp:430-454).
//   distReNumber      (rebuild.cpp:72-85)   communities are numbered in order of
//                     first appearance of cvect[i], i ascending;
//   fill_newEdgesMap  (rebuild.cpp:244-279) entry (i -> t, w) becomes
//                     (newid[cvect[i]] -> newid[cvect[t]], w), parallel entries
//                     summed in the order met (i ascending, j ascending);
//   send_newEdges     (rebuild.cpp:405-427) CSR emitted with each adjacency
//                     sorted by tail (std::map order).
// Sorting / scanning uses CUB (library code, not on the sweep hot path); a
// stable radix sort keeps the reference's summation order, and in
// CVB_WEIGHTS_F64 mode each merged entry is summed sequentially in that order,
// so the new weights are bit-identical to the reference's.
#pragma once

#include <cub/cub.cuh>

#include "common.cuh"

namespace cvb {

__global__ void k_fill_i32(int32_t *p, int64_t n, int32_t v) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    p[i] = v;
}

// first[c] = smallest vertex index whose community is c
__global__ void k_first_appearance(const vid_t *__restrict__ cvect, int64_t nv, int32_t *__restrict__ first) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nv;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    atomicMin(&first[cvect[i]], static_cast<int32_t>(i));
}

__global__ void k_flag_first(const vid_t *__restrict__ cvect, const int32_t *__restrict__ first, int64_t nv,
                             int32_t *__restrict__ flag) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nv;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    flag[i] = (first[cvect[i]] == static_cast<int32_t>(i)) ? 1 : 0;
}

// vnew[i] = new id of vertex i's community = rank[first[cvect[i]]]
__global__ void k_vertex_newid(const vid_t *__restrict__ cvect, const int32_t *__restrict__ first,
                               const int32_t *__restrict__ rank, int64_t nv, int32_t *__restrict__ vnew) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nv;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    vnew[i] = rank[first[cvect[i]]];
}

// compose[p] = vnew[compose[p]]  (main.cpp:401-404 with rebuild's numbering)
__global__ void k_compose(int32_t *__restrict__ compose, const int32_t *__restrict__ vnew, int64_t n) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    compose[i] = vnew[compose[i]];
}

__global__ void k_iota_i32(int32_t *p, int64_t n) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    p[i] = static_cast<int32_t>(i);
}

// One warp per row: key[e] = (vnew[row] << 32) | vnew[tail[e]]
// (vnew is indexed by GLOBAL vertex id; row r of this engine is vertex base + r)
__global__ void k_edge_keys(const eid_t *__restrict__ offsets, const vid_t *__restrict__ tails,
                            const int32_t *__restrict__ vnew, int64_t base, int64_t nv, uint64_t *__restrict__ keys) {
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const unsigned lane = threadIdx.x & 31u;
  if (row >= nv) return;
  const uint64_t hi = static_cast<uint64_t>(static_cast<uint32_t>(vnew[base + row])) << 32;
  for (eid_t e = offsets[row] + lane; e < offsets[row + 1]; e += 32)
    keys[e] = hi | static_cast<uint32_t>(vnew[tails[e]]);
}

// head[e] = 1 where a new (src, dst) pair starts in the sorted key array
__global__ void k_heads(const uint64_t *__restrict__ keys, int64_t ne, int32_t *__restrict__ head) {
  for (int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < ne;
       e += static_cast<int64_t>(gridDim.x) * blockDim.x)
    head[e] = (e == 0 || keys[e] != keys[e - 1]) ? 1 : 0;
}

// seg[e] = inclusive-scan(head)[e] - 1 is the merged entry of sorted entry e.
// Emits the distinct keys and the start of each merged entry.
__global__ void k_emit_pairs(const uint64_t *__restrict__ keys, const int32_t *__restrict__ head,
                             const int32_t *__restrict__ seg_incl, int64_t ne, uint64_t *__restrict__ ukeys,
                             int64_t *__restrict__ seg_start) {
  for (int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < ne;
       e += static_cast<int64_t>(gridDim.x) * blockDim.x)
    if (head[e]) {
      const int32_t s = seg_incl[e] - 1;
      ukeys[s] = keys[e];
      seg_start[s] = e;
    }
}

// Distinct ascending pairs -> CSR pieces: tails[s] = dst, row_count[src - base]++.
// A source outside [base, base + nv_local) raises *bad (the caller routed a pair wrongly).
// One thread per pair (the grid covers m exactly once).  The pairs ascend by source, so the pairs of a row are
// contiguous: a warp counts its runs with one ballot and only the first lane of a run issues the atomic — one per
// row and warp instead of one per pair (hub rows of a coarsened graph hold millions of pairs).
__global__ void k_pairs_to_rows(const uint64_t *__restrict__ ukeys, int64_t m, int64_t base, int64_t nv_local,
                                vid_t *__restrict__ ntails, unsigned long long *__restrict__ row_count,
                                int *__restrict__ bad) {
  const int64_t s = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const unsigned lane = threadIdx.x & 31u;
  int64_t row = -2;  // beyond the last pair: a run of its own, never counted
  if (s < m) {
    const uint64_t k = ukeys[s];
    row = static_cast<int64_t>(k >> 32) - base;
    ntails[s] = static_cast<vid_t>(k & 0xFFFFFFFFull);
    if (row < 0 || row >= nv_local) {
      *bad = 1;
      row = -1;
    }
  }
  const int64_t prev = __shfl_up_sync(0xffffffffu, row, 1);
  const bool head = lane == 0 || prev != row;
  const unsigned heads = __ballot_sync(0xffffffffu, head);
  if (head && row >= 0) {
    const unsigned above = heads & ~((2u << lane) - 1u);  // heads after this lane (lane 31: none)
    const int next = above ? __ffs(above) - 1 : 32;
    atomicAdd(&row_count[row], static_cast<unsigned long long>(next - static_cast<int>(lane)));
  }
}

// Weight of merged entry s = sum of the sorted weights in [seg_start[s], seg_start[s+1])
// added sequentially (the reference's order).  One thread per merged entry for
// short runs; long runs (heavy self-loops of big communities) use a warp when
// the order is irrelevant (uint32) — see k_sum_segments_u32.
__global__ void k_sum_segments_f64(const double *__restrict__ w_sorted, const int64_t *__restrict__ seg_start,
                                   int64_t nseg, int64_t ne, double *__restrict__ out) {
  for (int64_t s = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; s < nseg;
       s += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t a = seg_start[s], b = (s + 1 < nseg) ? seg_start[s + 1] : ne;
    double acc = 0.0;
    for (int64_t e = a; e < b; ++e) acc += w_sorted[e];
    out[s] = acc;
  }
}

// uint32 weights (exact in any order): one thread per SORTED ENTRY.  The entries of a merged pair are contiguous,
// so a warp first adds the runs it holds with a segmented shuffle scan (seg ids ascend) and only the last lane
// of every run issues an atomic — one per merged pair in the common case, one per warp and run for the long
// runs of heavy community pairs.  `out` must be zero.  (Round 2 launch list, R-MAT scale 22: the former
// warp-per-merged-pair kernel took 17.7 ms per level — 100 M warps for runs of length 1-2 — against 1.4 ms for
// a radix-sort pass over the same entries.)
__global__ void k_sum_segments_u32(const uint32_t *__restrict__ w_sorted, const int32_t *__restrict__ seg_incl,
                                   int64_t ne, uint32_t *__restrict__ out) {
  const int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;  // grid covers ne exactly once
  const unsigned lane = threadIdx.x & 31u;
  const bool valid = e < ne;
  const int32_t sg = valid ? seg_incl[e] : -1;
  uint32_t w = valid ? w_sorted[e] : 0u;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t wu = __shfl_up_sync(0xffffffffu, w, d);
    const int32_t su = __shfl_up_sync(0xffffffffu, sg, d);
    if (lane >= static_cast<unsigned>(d) && su == sg) w += wu;
  }
  const int32_t sn = __shfl_down_sync(0xffffffffu, sg, 1);
  if (valid && (lane == 31u || sn != sg)) atomicAdd(&out[sg - 1], w);
}

}  // namespace cvb
