// setup.cuh — per-phase setup kernels (layout conversion, vertex degrees, degree
// classes), the community-info apply + modularity reduction, and the small
// scatter/gather helpers of the per-iteration boundary call.
#pragma once

#include "common.cuh"

namespace cvb {

// Interface layout of one CSR entry: the reference's Edge{GraphElem tail;
// GraphWeight weight} in its 64-bit build (edge.hpp:22-30).
struct EdgeAoS {
  int64_t tail;
  double weight;
};

// ---- weight statistics: can the graph use the exact uint32 mode? ----------
// flags[0] |= 1 when some weight is negative, non-integer or >= 2^32;
// sum += all weights (double; only compared against 2^32, so the atomic
// order does not matter).
template <bool AOS>
__global__ void k_weight_stats(const void *__restrict__ src, int64_t ne, int *flags, double *sum) {
  double local = 0.0;
  int bad = 0;
  for (int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < ne;
       e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const double w = AOS ? static_cast<const EdgeAoS *>(src)[e].weight : static_cast<const double *>(src)[e];
    if (!(w >= 0.0) || w >= 4294967296.0 || w != floor(w)) bad = 1;
    local += w;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  bad = __any_sync(0xffffffffu, bad);
  if ((threadIdx.x & 31) == 0) {
    if (bad) atomicOr(flags, 1);
    atomicAdd(sum, local);
  }
}

// ---- boundary layout -> device layout --------------------------------------
template <bool AOS, typename W>
__global__ void k_convert_edges(const void *__restrict__ aos, const int64_t *__restrict__ tails64,
                                const double *__restrict__ w64, int64_t ne, vid_t *__restrict__ tails,
                                W *__restrict__ weights) {
  for (int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < ne;
       e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    int64_t t;
    double w;
    if (AOS) {
      const EdgeAoS x = static_cast<const EdgeAoS *>(aos)[e];
      t = x.tail;
      w = x.weight;
    } else {
      t = tails64[e];
      w = w64[e];
    }
    tails[e] = static_cast<vid_t>(t);
    weights[e] = static_cast<W>(w);
  }
}

__global__ void k_flag_to_f64(const int *__restrict__ flag, double *__restrict__ out) { out[0] = flag[0] ? 1.0 : 0.0; }

__global__ void k_fill_f64(double *__restrict__ dst, int64_t n, double v) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    dst[i] = v;
}

__global__ void k_fill_u32(uint32_t *__restrict__ dst, int64_t n, uint32_t v) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    dst[i] = v;
}

template <typename T64, typename T32>
__global__ void k_narrow(const T64 *__restrict__ src, T32 *__restrict__ dst, int64_t n, int64_t add) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    dst[i] = static_cast<T32>(src[i] + add);
}

template <typename T32, typename T64>
__global__ void k_widen(const T32 *__restrict__ src, T64 *__restrict__ dst, int64_t n) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    dst[i] = static_cast<T64>(src[i]);
}

// ---- distSumVertexDegree (louvain.cpp:2126-2151) ----------------------------
// uint32 mode: one warp per vertex, any order is exact.  double mode: one
// thread per vertex adds in edge order, exactly like the reference's loop.
template <typename W>
__global__ void k_vertex_degree(const eid_t *__restrict__ offsets, const W *__restrict__ weights, int64_t nv,
                                double *__restrict__ vdeg);

template <>
__global__ void k_vertex_degree<uint32_t>(const eid_t *__restrict__ offsets, const uint32_t *__restrict__ weights,
                                          int64_t nv, double *__restrict__ vdeg) {
  const int64_t v = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const unsigned lane = threadIdx.x & 31u;
  if (v >= nv) return;
  const eid_t e0 = offsets[v], e1 = offsets[v + 1];
  unsigned long long local = 0;
  for (eid_t e = e0 + lane; e < e1; e += 32) local += weights[e];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if (lane == 0) vdeg[v] = static_cast<double>(local);
}

template <>
__global__ void k_vertex_degree<double>(const eid_t *__restrict__ offsets, const double *__restrict__ weights,
                                        int64_t nv, double *__restrict__ vdeg) {
  const int64_t v = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (v >= nv) return;
  double tw = 0.0;
  for (eid_t e = offsets[v]; e < offsets[v + 1]; ++e) tw += weights[e];
  vdeg[v] = tw;
}

// Deterministic sum of a double array (fixed grid, block partials, last block
// adds them in block order).
__global__ void __launch_bounds__(512) k_sum_f64(const double *__restrict__ x, int64_t n,
                                                 double *__restrict__ partials, unsigned int *__restrict__ ticket,
                                                 double *__restrict__ out) {
  __shared__ double sh[16];
  __shared__ bool last;
  double s = 0.0;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    s += x[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double b = 0.0;
    for (int w = 0; w < 16; ++w) b += sh[w];
    partials[blockIdx.x] = b;
    __threadfence();
    last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    double t = 0.0;
    for (unsigned b = 0; b < gridDim.x; ++b) t += __ldcg(&partials[b]);
    out[0] = t;
    *ticket = 0;
  }
}

// ---- degree classes: stable partition of the local vertex ids --------------
constexpr int kClassBlock = 256;

__global__ void __launch_bounds__(kClassBlock) k_class_count(const eid_t *__restrict__ offsets, int64_t nv,
                                                            int32_t *__restrict__ block_counts,
                                                            unsigned long long *__restrict__ class_entries,
                                                            unsigned long long *__restrict__ max_degree) {
  __shared__ int cnt[kNumClasses];
  __shared__ unsigned long long ent[kNumClasses];
  __shared__ unsigned long long maxd;
  if (threadIdx.x < kNumClasses) {
    cnt[threadIdx.x] = 0;
    ent[threadIdx.x] = 0;
  }
  if (threadIdx.x == 0) maxd = 0;
  __syncthreads();
  const int64_t v = static_cast<int64_t>(blockIdx.x) * kClassBlock + threadIdx.x;
  if (v < nv) {
    const int64_t deg = offsets[v + 1] - offsets[v];
    const int c = degree_class(deg);
    atomicAdd(&cnt[c], 1);
    atomicAdd(&ent[c], static_cast<unsigned long long>(deg));
    atomicMax(&maxd, static_cast<unsigned long long>(deg));
  }
  __syncthreads();
  if (threadIdx.x < kNumClasses) {
    block_counts[static_cast<int64_t>(blockIdx.x) * kNumClasses + threadIdx.x] = cnt[threadIdx.x];
    if (ent[threadIdx.x]) atomicAdd(&class_entries[threadIdx.x], ent[threadIdx.x]);
  }
  if (threadIdx.x == 0 && maxd) atomicMax(max_degree, maxd);
}

// Single block: exclusive scan of block_counts in class-major order.
// class_start[c] = first list position of class c; block_counts[b][c] becomes
// the list position where block b's class-c vertices start.
__global__ void __launch_bounds__(1024) k_class_scan(int32_t *__restrict__ block_counts, int64_t nblocks,
                                                    int64_t *__restrict__ class_start) {
  __shared__ int64_t warp_tot[32];
  __shared__ int64_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int c = 0; c < kNumClasses; ++c) {
    if (threadIdx.x == 0) class_start[c] = carry;
    for (int64_t b0 = 0; b0 < nblocks; b0 += 1024) {
      const int64_t b = b0 + threadIdx.x;
      const int64_t x = (b < nblocks) ? block_counts[b * kNumClasses + c] : 0;
      int64_t incl = x;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int64_t y = __shfl_up_sync(0xffffffffu, incl, o);
        if ((threadIdx.x & 31) >= o) incl += y;
      }
      if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = incl;
      __syncthreads();
      int64_t wbase = 0;
      for (int w = 0; w < (threadIdx.x >> 5); ++w) wbase += warp_tot[w];
      const int64_t excl = carry + wbase + incl - x;
      if (b < nblocks) block_counts[b * kNumClasses + c] = static_cast<int32_t>(excl);
      __syncthreads();
      if (threadIdx.x == 1023) carry = excl + x;
      __syncthreads();
    }
  }
  if (threadIdx.x == 0) class_start[kNumClasses] = carry;
}

__global__ void __launch_bounds__(kClassBlock) k_class_scatter(const eid_t *__restrict__ offsets, int64_t nv,
                                                              const int32_t *__restrict__ block_offsets,
                                                              vid_t *__restrict__ list) {
  __shared__ int warp_cnt[kClassBlock / 32][kNumClasses];
  const int64_t v = static_cast<int64_t>(blockIdx.x) * kClassBlock + threadIdx.x;
  const unsigned lane = threadIdx.x & 31u;
  const int wib = threadIdx.x >> 5;
  const int c = (v < nv) ? degree_class(offsets[v + 1] - offsets[v]) : -1;
  int rank_in_warp = 0;
  for (int k = 0; k < kNumClasses; ++k) {
    const unsigned b = __ballot_sync(0xffffffffu, c == k);
    if (c == k) rank_in_warp = __popc(b & ((1u << lane) - 1u));
    if (lane == 0) warp_cnt[wib][k] = __popc(b);
  }
  __syncthreads();
  if (c >= 0) {
    int before = 0;
    for (int w = 0; w < wib; ++w) before += warp_cnt[w][c];
    list[block_offsets[static_cast<int64_t>(blockIdx.x) * kNumClasses + c] + before + rank_in_warp] =
        static_cast<vid_t>(v);
  }
}

// ---- distInitLouvain / distInitComm (louvain.cpp:2103-2124, 2569-2586) ------
__global__ void k_phase_init_local(int64_t nv_local, vid_t base, vid_t *__restrict__ l0, vid_t *__restrict__ l1,
                                   vid_t *__restrict__ l2, double *__restrict__ cw) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nv_local;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const vid_t id = static_cast<vid_t>(i) + base;
    l0[i] = id;
    l1[i] = id;
    l2[i] = id;
    cw[i] = 0.0;
  }
}

// Community arrays are replicated over all nv_global ids.  `vdeg_global` is
// the vertex degree of every global vertex (one GPU: the local array).
__global__ void k_phase_init_global(int64_t nv_global, const double *__restrict__ vdeg_global,
                                    vid_t *__restrict__ labels_global, double *__restrict__ cdeg,
                                    int32_t *__restrict__ csize, double *__restrict__ ddeg,
                                    int32_t *__restrict__ dsize) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nv_global;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    if (labels_global) labels_global[i] = static_cast<vid_t>(i);
    if (vdeg_global) cdeg[i] = vdeg_global[i];
    csize[i] = 1;
    ddeg[i] = 0.0;
    dsize[i] = 0;
  }
}

// ---- distUpdateLocalCinfo + distComputeModularity ---------------------------
// (louvain.cpp:2539-2552, 2433-2481; modularity.cu:36-100).  One streaming
// pass: Cinfo += delta and delta = 0 over [0, nv_global); a2 = sum degree^2
// over the owned communities [own0, own1); e = sum clusterWeight over the
// local vertices.  Two-level deterministic reduction: fixed grid, per-block
// partials, the last block to finish adds them in block order.
constexpr int kReduceBlock = 512;

__global__ void __launch_bounds__(kReduceBlock) k_apply_reduce(int64_t nv_global, int64_t own0, int64_t own1,
                                                              int64_t nv_local, double *__restrict__ cdeg,
                                                              int32_t *__restrict__ csize,
                                                              double *__restrict__ ddeg, int32_t *__restrict__ dsize,
                                                              const double *__restrict__ cw,
                                                              double *__restrict__ partials,
                                                              unsigned int *__restrict__ ticket,
                                                              double *__restrict__ out) {
  __shared__ double sh_e[kReduceBlock / 32], sh_a[kReduceBlock / 32];
  __shared__ bool last;
  double e = 0.0, a2 = 0.0;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t t0 = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  for (int64_t c = t0; c < nv_global; c += stride) {
    const double dd = ddeg[c];
    const int32_t ds = dsize[c];
    double d = cdeg[c];
    if (dd != 0.0 || ds != 0) {
      d += dd;
      cdeg[c] = d;
      csize[c] += ds;
      ddeg[c] = 0.0;
      dsize[c] = 0;
    }
    if (c >= own0 && c < own1) a2 += d * d;
  }
  for (int64_t i = t0; i < nv_local; i += stride) e += cw[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    e += __shfl_xor_sync(0xffffffffu, e, o);
    a2 += __shfl_xor_sync(0xffffffffu, a2, o);
  }
  if ((threadIdx.x & 31) == 0) {
    sh_e[threadIdx.x >> 5] = e;
    sh_a[threadIdx.x >> 5] = a2;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double be = 0.0, ba = 0.0;
    for (int w = 0; w < kReduceBlock / 32; ++w) {
      be += sh_e[w];
      ba += sh_a[w];
    }
    partials[2 * blockIdx.x] = be;
    partials[2 * blockIdx.x + 1] = ba;
    __threadfence();
    const unsigned int t = atomicAdd(ticket, 1u);
    last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (last) {  // block-uniform: the last block adds the partials in a fixed order
    __threadfence();
    double te = 0.0, ta = 0.0;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += kReduceBlock) {
      te += __ldcg(&partials[2 * b]);
      ta += __ldcg(&partials[2 * b + 1]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      te += __shfl_xor_sync(0xffffffffu, te, o);
      ta += __shfl_xor_sync(0xffffffffu, ta, o);
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) {
      sh_e[threadIdx.x >> 5] = te;
      sh_a[threadIdx.x >> 5] = ta;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double fe = 0.0, fa = 0.0;
      for (int w = 0; w < kReduceBlock / 32; ++w) {
        fe += sh_e[w];
        fa += sh_a[w];
      }
      out[0] = fe;
      out[1] = fa;
      *ticket = 0;
    }
  }
}

// ---- labels-only exchange (N > 1) ------------------------------------------------
// Every rank holds the replicated labels before (curr) and after (next) the
// sweep, so it derives the community deltas of ALL ranks' moves locally:
// cdeg[next] += vdeg, cdeg[curr] -= vdeg, sizes +-1 for every vertex that moved.
// Exact for integer degrees; work and atomic traffic are proportional to the
// number of moved vertices, and nothing but the labels crossed NVLink.
__global__ void k_apply_label_diff(int64_t nv_global, const vid_t *__restrict__ curr, const vid_t *__restrict__ next,
                                   const double *__restrict__ vdeg_global, double *__restrict__ cdeg,
                                   int32_t *__restrict__ csize) {
  for (int64_t v = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; v < nv_global;
       v += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const vid_t a = curr[v], b = next[v];
    if (a != b) {
      const double d = vdeg_global[v];
      atomicAdd(&cdeg[b], d);
      atomicAdd(&cdeg[a], -d);
      atomicAdd(&csize[b], 1);
      atomicAdd(&csize[a], -1);
    }
  }
}

// Reduction only (labels-only exchange: the deltas were applied from the label diff):
// a2 over the owned communities, e over the local vertices; same two-level scheme.
__global__ void __launch_bounds__(kReduceBlock) k_reduce_only(int64_t own0, int64_t own1, int64_t nv_local,
                                                             const double *__restrict__ cdeg,
                                                             const double *__restrict__ cw,
                                                             double *__restrict__ partials,
                                                             unsigned int *__restrict__ ticket,
                                                             double *__restrict__ out) {
  __shared__ double sh_e[kReduceBlock / 32], sh_a[kReduceBlock / 32];
  __shared__ bool last;
  double e = 0.0, a2 = 0.0;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t t0 = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  for (int64_t c = own0 + t0; c < own1; c += stride) {
    const double d = cdeg[c];
    a2 += d * d;
  }
  for (int64_t i = t0; i < nv_local; i += stride) e += cw[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    e += __shfl_xor_sync(0xffffffffu, e, o);
    a2 += __shfl_xor_sync(0xffffffffu, a2, o);
  }
  if ((threadIdx.x & 31) == 0) {
    sh_e[threadIdx.x >> 5] = e;
    sh_a[threadIdx.x >> 5] = a2;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double be = 0.0, ba = 0.0;
    for (int w = 0; w < kReduceBlock / 32; ++w) {
      be += sh_e[w];
      ba += sh_a[w];
    }
    partials[2 * blockIdx.x] = be;
    partials[2 * blockIdx.x + 1] = ba;
    __threadfence();
    last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (last) {
    __threadfence();
    double te = 0.0, ta = 0.0;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += kReduceBlock) {
      te += __ldcg(&partials[2 * b]);
      ta += __ldcg(&partials[2 * b + 1]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      te += __shfl_xor_sync(0xffffffffu, te, o);
      ta += __shfl_xor_sync(0xffffffffu, ta, o);
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) {
      sh_e[threadIdx.x >> 5] = te;
      sh_a[threadIdx.x >> 5] = ta;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double fe = 0.0, fa = 0.0;
      for (int w = 0; w < kReduceBlock / 32; ++w) {
        fe += sh_e[w];
        fa += sh_a[w];
      }
      out[0] = fe;
      out[1] = fa;
      *ticket = 0;
    }
  }
}

// ---- early termination (options -t 1..4 / -a; louvain.cpp:7-204, 206-423) ------------------
// After a sweep that ran with the delta atomics switched off: frozen vertices get their frozen
// target / clusterWeight back (louvain.cpp:97-101: an inactive vertex is skipped, its clusterWeight
// is the frozen one and it contributes no community update); active vertices record the new
// clusterWeight as the frozen one (louvain.cpp:107) and contribute their move.
__global__ void k_et_fix(int64_t nvl, const unsigned char *__restrict__ active, vid_t *__restrict__ target,
                         const vid_t *__restrict__ curr, const vid_t *__restrict__ frozen_target,
                         double *__restrict__ cw, double *__restrict__ frozen_cw, const double *__restrict__ vdeg,
                         double *__restrict__ ddeg, int32_t *__restrict__ dsize) {
  for (int64_t v = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; v < nvl;
       v += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    if (!active[v]) {
      target[v] = frozen_target[v];
      cw[v] = frozen_cw[v];
    } else {
      frozen_cw[v] = cw[v];
      const vid_t t = target[v], c = curr[v];
      if (t != c) {
        const double d = vdeg[v];
        atomicAdd(&ddeg[t], d);
        atomicAdd(&dsize[t], 1);
        atomicAdd(&ddeg[c], -d);
        atomicAdd(&dsize[c], -1);
      }
    }
  }
}

// The same over a vertex list (a colour step of the coloured phase with early termination, louvain.cpp:1086-1098).
__global__ void k_et_fix_list(const vid_t *__restrict__ list, int64_t n, const unsigned char *__restrict__ active,
                              vid_t *__restrict__ target, const vid_t *__restrict__ curr,
                              const vid_t *__restrict__ frozen_target, double *__restrict__ cw,
                              double *__restrict__ frozen_cw, const double *__restrict__ vdeg,
                              double *__restrict__ ddeg, int32_t *__restrict__ dsize) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const vid_t v = list[i];
    if (!active[v]) {
      target[v] = frozen_target[v];
      cw[v] = frozen_cw[v];
    } else {
      frozen_cw[v] = cw[v];
      const vid_t t = target[v], c = curr[v];
      if (t != c) {
        const double d = vdeg[v];
        atomicAdd(&ddeg[t], d);
        atomicAdd(&dsize[t], 1);
        atomicAdd(&ddeg[c], -d);
        atomicAdd(&dsize[c], -1);
      }
    }
  }
}

// The per-vertex label rotation with the freeze rules, bug for bug:
//   types 1, 3 (louvain.cpp:174-186): EVERY vertex, frozen or not, either freezes — when
//     numIters > 2 and `targetComm[i] == currComm[i] == pastComm[i]`, which C++ parses as
//     (target == curr) == past, a bool compared with a community id — or rotates its labels;
//   types 2, 4 (louvain.cpp:375-398): a frozen vertex is skipped; an active one whose label did not
//     change (curr == past, numIters > 2) multiplies its probability by (1 - delta) and freezes at
//     p <= 0.02 (P_CUTOFF, louvain.hpp:80); active vertices rotate.  p_curr is only written on that
//     branch and the two arrays are swapped afterwards (louvain.cpp:401), stale values included.
__global__ void k_et_rotate(int64_t nvl, int num_iters, int prob, double one_minus_delta, vid_t *__restrict__ past,
                            vid_t *__restrict__ curr, vid_t *__restrict__ target, unsigned char *__restrict__ active,
                            vid_t *__restrict__ frozen_target, const double *__restrict__ p_prev,
                            double *__restrict__ p_curr, unsigned int *__restrict__ frozen_count) {
  for (int64_t v = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; v < nvl;
       v += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    bool rotate;
    if (!prob) {
      const bool freeze = num_iters > 2 && (static_cast<vid_t>(target[v] == curr[v]) == past[v]);
      if (freeze && active[v]) {
        active[v] = 0;
        frozen_target[v] = target[v];
        atomicAdd(frozen_count, 1u);
      }
      rotate = !freeze;
    } else {
      if (!active[v]) continue;
      if (num_iters > 2 && curr[v] == past[v]) {
        const double pc = p_prev[v] * one_minus_delta;
        p_curr[v] = pc;
        if (pc <= 0.02) {
          active[v] = 0;
          frozen_target[v] = target[v];
          atomicAdd(frozen_count, 1u);
        }
      }
      rotate = active[v] != 0;
    }
    if (rotate) {
      const vid_t tmp = past[v];
      past[v] = curr[v];
      curr[v] = target[v];
      target[v] = tmp;
      if (!active[v]) frozen_target[v] = target[v];  // types 1, 3: a frozen vertex that rotates carries its labels along
    }
  }
}

// ---- distance-1 colouring (option -c; coloring.cpp:3-200) -------------------------------------
// The reference's integer hash (coloring.cpp:74-85).
__host__ __device__ inline unsigned int color_hash(unsigned int a, unsigned int seed) {
  a ^= seed;
  a = (a + 0x7ed55d16u) + (a << 12);
  a = (a ^ 0xc761c23cu) + (a >> 19);
  a = (a + 0x165667b1u) + (a << 5);
  a = (a ^ 0xd3a2646cu) + (a << 9);
  a = (a + 0xfd7046c5u) + (a << 3);
  a = (a ^ 0xb55a4f09u) + (a >> 16);
  return a;
}

// One round of distColoringIteration (coloring.cpp:87-200), one warp per vertex.  An uncoloured vertex
// compares its nHash hash values with those of its neighbours that were not coloured in an EARLIER
// round (colours assigned in this round are >= next_color and count as uncoloured, so the round does
// not depend on the order in which vertices are visited); where it is the strict minimum / maximum of
// hash t it may take colour 2t / 2t+1 (+ next_color), and it picks the (id mod #possible)-th of them.
__global__ void k_color_round(const eid_t *__restrict__ offsets, const vid_t *__restrict__ tails, int64_t nv, vid_t base,
                              int n_hash, int next_color, unsigned int seed, int32_t *__restrict__ color,
                              unsigned int *__restrict__ colored_count) {
  const int64_t v = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const unsigned lane = threadIdx.x & 31u;
  if (v >= nv) return;
  if (color[v] != -1) return;  // warp-uniform
  const unsigned int self = static_cast<unsigned int>(v + base);
  unsigned int not_min = 0u, not_max = 0u;
  for (eid_t e = offsets[v] + lane; e < offsets[v + 1]; e += 32) {
    const vid_t t = tails[e];
    if (static_cast<unsigned int>(t) == self) continue;  // self-loops are ignored
    const int32_t ct = color[t - base];
    if (ct != -1 && ct < next_color) continue;  // coloured in an earlier round
    for (int h = 0; h < n_hash; ++h) {
      const unsigned int vh = color_hash(self, seed + 1043u * h);
      const unsigned int jh = color_hash(static_cast<unsigned int>(t), seed + 1043u * h);
      if (vh <= jh) not_max |= 1u << h;
      if (vh >= jh) not_min |= 1u << h;
    }
  }
  not_min = __reduce_or_sync(0xffffffffu, not_min);
  not_max = __reduce_or_sync(0xffffffffu, not_max);
  if (lane != 0) return;
  const int possible = 2 * n_hash - __popc(not_min) - __popc(not_max);
  if (possible <= 0) return;
  const int col_id = static_cast<int>(static_cast<int64_t>(self) % possible);
  int this_id = 0;
  for (int h = 0; h < n_hash; ++h) {
    if (!(not_min & (1u << h)) && col_id == this_id) {
      color[v] = 2 * h + next_color;
      atomicAdd(colored_count, 1u);
      return;
    }
    this_id += !(not_min & (1u << h));
    if (!(not_max & (1u << h)) && col_id == this_id) {
      color[v] = 2 * h + 1 + next_color;
      atomicAdd(colored_count, 1u);
      return;
    }
    this_id += !(not_max & (1u << h));
  }
}

// sort key of a vertex for the per-colour class lists: (colour bucket, degree class); uncoloured vertices go
// to the last bucket (louvain.cpp:813, 829)
__global__ void k_color_keys(const eid_t *__restrict__ offsets, const int32_t *__restrict__ color, int64_t nv,
                             int num_color, int32_t *__restrict__ keys, vid_t *__restrict__ ids) {
  for (int64_t v = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; v < nv;
       v += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int b = color[v] < 0 ? num_color - 1 : color[v];
    keys[v] = b * kNumClasses + degree_class(offsets[v + 1] - offsets[v]);
    ids[v] = static_cast<vid_t>(v);
  }
}
__global__ void k_key_histogram(const int32_t *__restrict__ keys, int64_t n, unsigned int *__restrict__ hist) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    atomicAdd(&hist[keys[i]], 1u);
}

// Big vertices (above 8192 entries) in a colour step: their kernels ran over ALL of them with the delta
// atomics off; vertices of the current colour keep the result and contribute their move, the others get
// the values of their own colour step (or of the start of the iteration) back.
// With early termination (active != nullptr) a frozen vertex of the colour takes its frozen target / clusterWeight
// and moves nothing; an active one records the new clusterWeight as the frozen one.
__global__ void k_color_fix_big(const vid_t *__restrict__ big_list, int nbig, const int32_t *__restrict__ color,
                                int num_color, int ci, vid_t *__restrict__ target, const vid_t *__restrict__ curr,
                                double *__restrict__ cw, vid_t *__restrict__ bak_t, double *__restrict__ bak_cw,
                                const double *__restrict__ vdeg, double *__restrict__ ddeg, int32_t *__restrict__ dsize,
                                const unsigned char *__restrict__ active, const vid_t *__restrict__ frozen_target,
                                double *__restrict__ frozen_cw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nbig) return;
  const vid_t v = big_list[i];
  const int b = color[v] < 0 ? num_color - 1 : color[v];
  if (b == ci) {
    if (active && !active[v]) {
      target[v] = frozen_target[v];
      cw[v] = frozen_cw[v];
      bak_t[i] = target[v];
      bak_cw[i] = cw[v];
      return;
    }
    if (active) frozen_cw[v] = cw[v];
    bak_t[i] = target[v];
    bak_cw[i] = cw[v];
    const vid_t t = target[v], c = curr[v];
    if (t != c) {
      const double d = vdeg[v];
      atomicAdd(&ddeg[t], d);
      atomicAdd(&dsize[t], 1);
      atomicAdd(&ddeg[c], -d);
      atomicAdd(&dsize[c], -1);
    }
  } else {
    target[v] = bak_t[i];
    cw[v] = bak_cw[i];
  }
}
__global__ void k_color_bak_init(const vid_t *__restrict__ big_list, int nbig, const vid_t *__restrict__ target,
                                 vid_t *__restrict__ bak_t, double *__restrict__ bak_cw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nbig) return;
  bak_t[i] = target[big_list[i]];
  bak_cw[i] = 0.0;
}

// ---- boundary-call helpers ---------------------------------------------------
__global__ void k_scatter_labels(int64_t n, const int64_t *__restrict__ idx, const int64_t *__restrict__ val,
                                 vid_t *__restrict__ dst) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    dst[idx[i]] = static_cast<vid_t>(val[i]);
}

__global__ void k_scatter_cinfo(int64_t n, const int64_t *__restrict__ idx, const int64_t *__restrict__ size,
                                const double *__restrict__ deg, int64_t add, int32_t *__restrict__ csize,
                                double *__restrict__ cdeg) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t c = idx ? idx[i] : i + add;
    csize[c] = static_cast<int32_t>(size[i]);
    cdeg[c] = deg[i];
  }
}

// out_size/out_deg[i] = delta at community idx[i] (or add+i); the delta is
// cleared afterwards so the arrays are zero for the next call.
__global__ void k_gather_deltas(int64_t n, const int64_t *__restrict__ idx, int64_t add,
                                double *__restrict__ ddeg, int32_t *__restrict__ dsize,
                                int64_t *__restrict__ out_size, double *__restrict__ out_deg) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t c = idx ? idx[i] : i + add;
    out_size[i] = dsize[c];
    out_deg[i] = ddeg[c];
    dsize[c] = 0;
    ddeg[c] = 0.0;
  }
}

}  // namespace cvb
