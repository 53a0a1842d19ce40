// sweep_u32.cuh — the fast local-move sweep for CVB_WEIGHTS_U32 graphs (all
// weights non-negative integers, total < 2^32: every sum below is exact, so the
// result is independent of accumulation order and bit-identical to the
// reference's double arithmetic, louvain.cpp:2185-2431).
//
// Design, from measurements on B200 (profiles/r01_microbench_b200.txt):
//   * shared-memory atomics run at ~8.9 lane-ops/clk/SM, as fast as plain
//     random LDS, while match.any on 32 distinct keys costs 64 clk per warp —
//     so the per-community accumulation is an open-addressing hash in shared
//     memory claimed with atomicCAS and accumulated with atomicAdd;
//   * the binding resource is the random per-entry gather (~290 G sectors/s out
//     of L2) — so labels are PACKED: one 8-byte entry per vertex holds the
//     vertex's community, a "community is a singleton" bit and the community's
//     degree, and a single gather per CSR entry brings everything the gain needs
//     (no second gather into the community-info arrays, no size lookups);
//   * cost after the insert phase is proportional to the number of DISTINCT
//     neighbour communities (the lane that claimed a slot evaluates it), not to
//     the table capacity.
#pragma once

#include <climits>

#include "common.cuh"
#include "sweep.cuh"

namespace cvb {

constexpr uint32_t kEmptyKeyU = 0xFFFFFFFFu;
constexpr uint32_t kLabelMask = 0x7FFFFFFFu;
constexpr uint32_t kSingletonBit = 0x80000000u;

struct SweepU32Params {
  const eid_t *__restrict__ offsets;    // [nv_local+1]
  const vid_t *__restrict__ tails;      // [ne]
  const uint32_t *__restrict__ weights; // [ne]
  const uint2 *__restrict__ packed;     // [nv_global] {label | singleton<<31, degree of that community}
  const double *__restrict__ cdeg;      // [nv_global] community degrees (giant-table evaluation only)
  const double *__restrict__ vdeg;      // [nv_local]
  double constant;
  vid_t base;
  vid_t *__restrict__ target;  // [nv_local]
  double *__restrict__ cw;     // [nv_local]
  double *__restrict__ ddeg;   // [nv_global]
  int32_t *__restrict__ dsize; // [nv_global]
  // split-gather variant (CVB_SPLIT=1): 4-byte label gather per entry, community degree per DISTINCT community
  const vid_t *__restrict__ labels;      // [nv_global] the labels every sweep reads
  const uint32_t *__restrict__ cdeg32;   // [nv_global] (uint32) cdeg
  const int32_t *__restrict__ csize;     // [nv_global]
  // degree-ordered gather layout (lm.cuh): packed[] is indexed by rank[global id] and `tails` holds ranks
  const vid_t *__restrict__ rank;        // [nv_global] or nullptr (identity)
  int simple_rows;                       // every row holds each tail at most once (singleton neighbours bypass the table)
};

// index of local vertex v's own record in packed[] (and the value its self-loop entries carry in tails[])
__device__ __forceinline__ vid_t self_of(const SweepU32Params &p, vid_t v) {
  const vid_t g = v + p.base;
  return p.rank != nullptr ? p.rank[g] : g;
}

__global__ void k_cdeg32(int64_t nv_global, const double *__restrict__ cdeg, uint32_t *__restrict__ cdeg32) {
  for (int64_t c = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; c < nv_global;
       c += static_cast<int64_t>(gridDim.x) * blockDim.x)
    cdeg32[c] = static_cast<uint32_t>(cdeg[c]);
}

// packed[v] = {labels[v] | (csize[labels[v]] == 1) << 31, (uint32) cdeg[labels[v]]}
// rank (optional): the record of vertex v is written at packed[rank[v]] (degree-ordered gather layout, lm.cuh)
__global__ void k_pack_labels(int64_t nv_global, const vid_t *__restrict__ labels, const double *__restrict__ cdeg,
                              const int32_t *__restrict__ csize, uint2 *__restrict__ packed,
                              const vid_t *__restrict__ rank = nullptr) {
  for (int64_t v = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; v < nv_global;
       v += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const vid_t l = labels[v];
    const uint32_t flag = (csize[l] == 1) ? kSingletonBit : 0u;
    packed[rank != nullptr ? rank[v] : v] = make_uint2(static_cast<uint32_t>(l) | flag, static_cast<uint32_t>(cdeg[l]));
  }
}

__device__ __forceinline__ uint32_t hash_key(uint32_t key, int log2cap) {
  return ((key & kLabelMask) * 2654435761u) >> (32 - log2cap);
}

// best-candidate bookkeeping: ids carry the singleton bit in bit 31
__device__ __forceinline__ bool better_u(double g2, uint32_t id2, double g, uint32_t id) {
  return (g2 > g) || (g2 == g && (id2 & kLabelMask) < (id & kLabelMask));
}

__device__ __forceinline__ void commit_vertex_u32(const SweepU32Params &p, vid_t v, uint32_t me_key, double best_gain,
                                                  uint32_t best_key, double vd, double counter0) {
  const vid_t cc = static_cast<vid_t>(me_key & kLabelMask);
  vid_t best = (best_gain > 0.0) ? static_cast<vid_t>(best_key & kLabelMask) : cc;
  // louvain.cpp:2240-2241: two singletons only ever merge towards the smaller id
  if (best != cc && best > cc && (best_key & kSingletonBit) && (me_key & kSingletonBit)) best = cc;
  p.target[v] = best;
  p.cw[v] = counter0;
  if (best != cc && p.ddeg != nullptr) {
    atomicAdd(&p.ddeg[best], vd);
    atomicAdd(&p.dsize[best], 1);
    atomicAdd(&p.ddeg[cc], -vd);
    atomicAdd(&p.dsize[cc], -1);
  }
}

// ---------------------------------------------------------------------------
// deg 1..4: one thread per vertex, everything in registers.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_sweep_u32_t4(SweepU32Params p, const vid_t *__restrict__ list, int n) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const vid_t v = list[i];
  const eid_t e0 = p.offsets[v];
  const int deg = static_cast<int>(p.offsets[v + 1] - e0);
  const vid_t self = self_of(p, v);
  const uint2 me = p.packed[self];
  const uint32_t cc = me.x & kLabelMask;
  const double vd = p.vdeg[v];
  uint32_t key[4], w[4], ay[4];
  uint32_t selfloop = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    key[j] = kEmptyKeyU;
    w[j] = 0;
    ay[j] = 0;
    if (j < deg) {
      const vid_t t = p.tails[e0 + j];
      w[j] = p.weights[e0 + j];
      const uint2 pk = p.packed[t];
      key[j] = pk.x;
      ay[j] = pk.y;
      if (t == self) selfloop += w[j];
    }
  }
  // merge duplicates into the first occurrence
#pragma unroll
  for (int j = 1; j < 4; ++j)
#pragma unroll
    for (int k = 0; k < j; ++k)
      if (key[j] != kEmptyKeyU && key[j] == key[k]) {
        w[k] += w[j];
        key[j] = kEmptyKeyU;
      }
  uint32_t c0 = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (key[j] != kEmptyKeyU && (key[j] & kLabelMask) == cc) c0 = w[j];
  const double counter0 = u32_to_f64(c0);
  const double eix = counter0 - u32_to_f64(selfloop);
  const double ax = u32_to_f64(me.y) - vd;
  double g = -1.0;
  uint32_t id = kLabelMask;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (key[j] != kEmptyKeyU && (key[j] & kLabelMask) != cc) {
      const double g2 = modularity_gain(u32_to_f64(w[j]), eix, vd, u32_to_f64(ay[j]), ax, p.constant);
      if (better_u(g2, key[j], g, id)) {
        g = g2;
        id = key[j];
      }
    }
  commit_vertex_u32(p, v, me.x, g, id, vd, counter0);
}

// ---------------------------------------------------------------------------
// deg 5..32: G lanes per vertex, one batch, duplicates merged with match.any.
// ---------------------------------------------------------------------------
template <int G>
__global__ void __launch_bounds__(256) k_sweep_u32_group(SweepU32Params p, const vid_t *__restrict__ list, int n) {
  const unsigned lane = threadIdx.x & 31u;
  const int lig = static_cast<int>(lane) & (G - 1);
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~static_cast<unsigned>(G - 1)));
  const int64_t gid = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) / G;
  if (gid >= n) return;
  const vid_t v = list[gid];
  const eid_t e0 = p.offsets[v];
  const int deg = static_cast<int>(p.offsets[v + 1] - e0);
  const vid_t self = self_of(p, v);
  const uint2 me = p.packed[self];
  const uint32_t cc = me.x & kLabelMask;
  const bool valid = lig < deg;
  vid_t tail = -1;
  uint32_t w = 0, ay = 0;
  uint32_t key = 0xFFFFFF00u + lig;  // distinct idle keys (never a real label)
  if (valid) {
    tail = p.tails[e0 + lig];
    w = p.weights[e0 + lig];
    const uint2 pk = p.packed[tail];
    key = pk.x;
    ay = pk.y;
  }
  const unsigned m = __match_any_sync(gmask, key);
  const bool leader = valid && (static_cast<unsigned>(__ffs(m) - 1) == lane);
  const uint32_t s = group_sum_ordered<uint32_t>(gmask, m, w, 0u, lane);
  const unsigned ccb = __ballot_sync(gmask, leader && (key & kLabelMask) == cc);
  const uint32_t c0 = __shfl_sync(gmask, s, ccb ? (__ffs(ccb) - 1) : static_cast<int>(lane));
  const double counter0 = ccb ? u32_to_f64(c0) : 0.0;
  unsigned slb = __ballot_sync(gmask, valid && tail == self);
  double selfloop = 0.0;
  while (slb) {
    selfloop += u32_to_f64(__shfl_sync(gmask, w, __ffs(slb) - 1));
    slb &= slb - 1u;
  }
  const double vd = p.vdeg[v];
  const double eix = counter0 - selfloop;
  const double ax = u32_to_f64(me.y) - vd;
  double g = -1.0;
  uint32_t id = kLabelMask;
  if (leader && (key & kLabelMask) != cc) {
    g = modularity_gain(u32_to_f64(s), eix, vd, u32_to_f64(ay), ax, p.constant);
    id = key;
  }
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) {
    const double g2 = __shfl_xor_sync(gmask, g, o);
    const uint32_t id2 = __shfl_xor_sync(gmask, id, o);
    if (better_u(g2, id2, g, id)) {
      g = g2;
      id = id2;
    }
  }
  if (lig == 0) commit_vertex_u32(p, v, me.x, g, id, vd, counter0);
}

// ---------------------------------------------------------------------------
// deg 33..8192: TPV threads per vertex, K CSR entries per thread, hash table
// {keys, wts} in shared memory.
//   TPV == 32 : VPB warps per block, each warp owns a vertex and a table
//   TPV  > 32 : the whole block owns one vertex
// Persistent blocks stride over the class list.  A lane that CLAIMS a slot for
// a new key remembers the slot in a register and is the one that evaluates that
// community's gain after the last insert, then frees the slot — so there is no
// per-table list and no table scan, and the cost after the insert phase is
// proportional to the number of distinct neighbour communities.
// Per table: keys[CAP] wts[CAP] (uint32), meta[4] = {-, selfloop, counter0, -}.
// ---------------------------------------------------------------------------
// dynamic shared memory of one block of k_sweep_u32_hash<TPV, KC, NCH>
// TPV >= 128 (one block per vertex) also stages the vertex's tails and weights tile
// (KC*NCH*TPV + 4 entries each, filled by TMA) and one mbarrier.
template <int TPV, int KC, int NCH, bool TMA = false>
__host__ __device__ constexpr size_t hash_block_smem() {
  return (TPV <= 32) ? static_cast<size_t>(256 / TPV) * (2 * KC * NCH * TPV) * 8
                     : static_cast<size_t>(2 * KC * NCH * TPV) * 8 + 16 +
                           (TMA ? 2 * static_cast<size_t>(KC * NCH * TPV + 4) * 4 + 16 : 0);
}

template <int TPV>
__device__ __forceinline__ void tile_sync() {
  if (TPV == 32)
    __syncwarp();
  else
    __syncthreads();
}

// Warp argmax under the (gain desc, id asc) order, gains > 0 only, with three
// REDUX operations: positive doubles order like their bit patterns.
// Returns the winning key (singleton bit included) or kLabelMask when no lane
// holds a positive gain; *gain_out gets the winning gain (or -1).
__device__ __forceinline__ uint32_t warp_argmax(double g, uint32_t key, double *gain_out,
                                                unsigned gmask = 0xffffffffu) {
  const unsigned long long gb = (g > 0.0) ? static_cast<unsigned long long>(__double_as_longlong(g)) : 0ull;
  const uint32_t hi = static_cast<uint32_t>(gb >> 32), lo = static_cast<uint32_t>(gb);
  const uint32_t mhi = __reduce_max_sync(gmask, hi);
  const uint32_t mlo = __reduce_max_sync(gmask, hi == mhi ? lo : 0u);
  const bool top = gb != 0ull && hi == mhi && lo == mlo;
  const uint32_t mid = __reduce_min_sync(gmask, top ? (key & kLabelMask) : 0xFFFFFFFFu);
  const unsigned who = __ballot_sync(gmask, top && (key & kLabelMask) == mid);
  *gain_out = (who == 0u) ? -1.0
                          : __longlong_as_double(static_cast<long long>((static_cast<unsigned long long>(mhi) << 32) | mlo));
  const uint32_t wkey = __shfl_sync(gmask, key, who ? (__ffs(who) - 1) : 0);
  return who ? wkey : kLabelMask;
}

// Shared-memory accesses by 32-bit shared address (one address register +
// immediate offset in SASS instead of re-deriving generic pointers).
__device__ __forceinline__ uint32_t smem_addr(const void *p) {
  uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  // opaque move: keeps the address in one register instead of letting ptxas
  // re-derive the shared window base (S2UR/UMOV/ULEA) at every access
  asm volatile("mov.u32 %0, %0;" : "+r"(a));
  return a;
}
__device__ __forceinline__ uint32_t atoms_cas(uint32_t addr, uint32_t cmp, uint32_t val) {
  uint32_t old;
  asm volatile("atom.shared.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "r"(addr), "r"(cmp), "r"(val) : "memory");
  return old;
}
// Predicated forms (@p atom / @p red): no branch, no BSSY/BSYNC around a single memory op.
__device__ __forceinline__ uint32_t atoms_cas_if(bool pred, uint32_t addr, uint32_t cmp, uint32_t val, uint32_t dflt) {
  uint32_t old;
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %4, 0;\n\tmov.u32 %0, %5;\n\t@p atom.shared.cas.b32 %0, [%1], %2, %3;\n\t}"
      : "=r"(old)
      : "r"(addr), "r"(cmp), "r"(val), "r"(static_cast<uint32_t>(pred)), "r"(dflt)
      : "memory");
  return old;
}
__device__ __forceinline__ void reds_add_if(bool pred, uint32_t addr, uint32_t v) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p red.shared.add.u32 [%0], %1;\n\t}" ::"r"(addr), "r"(v),
               "r"(static_cast<uint32_t>(pred))
               : "memory");
}
__device__ __forceinline__ void reds_add(uint32_t addr, uint32_t v) {
  asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}

// ---- TMA (bulk async copy) + mbarrier: stage a vertex's tails / weights tile in shared memory ----
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t mbar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra WAIT_%=;\n\t}" ::"r"(mbar),
      "r"(parity)
      : "memory");
}
// 1-D bulk copy global -> shared, completion signalled on the mbarrier (UBLKCP in SASS).
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void *src, uint32_t bytes, uint32_t mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(mbar)
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Second hash for double hashing: an odd step, so every slot of the
// power-of-two table is reachable.
__device__ __forceinline__ uint32_t hash_step(uint32_t key, int log2cap) {
  return ((((key & kLabelMask) * 0x85EBCA6Bu) >> (32 - log2cap)) | 1u);
}

// Reductions over the GL contiguous lanes of one vertex.  GL == 32 uses REDUX;
// smaller groups use xor shuffles (REDUX with a per-group mask makes ptxas
// emit a match.any loop over the distinct masks).
template <int GL>
__device__ __forceinline__ uint32_t group_sum(uint32_t v, unsigned gmask) {
  if (GL == 32) return __reduce_add_sync(0xffffffffu, v);
#pragma unroll
  for (int o = GL / 2; o > 0; o >>= 1) v += __shfl_xor_sync(gmask, v, o);
  return v;
}

template <int GL>
__device__ __forceinline__ uint32_t group_argmax(double g, uint32_t key, double *gain_out, unsigned gmask) {
  if (GL == 32) return warp_argmax(g, key, gain_out, 0xffffffffu);
#pragma unroll
  for (int o = GL / 2; o > 0; o >>= 1) {
    const double g2 = __shfl_xor_sync(gmask, g, o);
    const uint32_t k2 = __shfl_xor_sync(gmask, key, o);
    if (better_u(g2, k2, g, key)) {
      g = g2;
      key = k2;
    }
  }
  *gain_out = g;
  return key;
}

// Insert KC entries per lane into the open-addressing table.  kaddr is the
// shared address of keys[]; wts[] follows at +capb bytes.  All KC first-probe
// claims are issued back to back (independent shared-memory atomics in flight);
// a collided entry then walks its own short probe loop (double hashing).  Bit k
// of the return value is set when this lane claimed a fresh slot for entry k;
// slb[k] is the entry's final slot as a BYTE offset.
template <int KC>
__device__ __forceinline__ uint32_t table_insert_chunk(uint32_t kaddr, uint32_t capb, const uint32_t (&key)[KC],
                                                       const uint32_t (&w)[KC], uint32_t active, int log2cap,
                                                       uint32_t maskb, uint32_t (&slb)[KC]) {
  uint32_t old[KC];
  uint32_t fresh = 0u, pend = 0u;
#pragma unroll
  for (int k = 0; k < KC; ++k) {
    slb[k] = hash_key(key[k], log2cap) << 2;
    old[k] = atoms_cas_if((active >> k) & 1u, kaddr + slb[k], kEmptyKeyU, key[k], key[k]);
  }
#pragma unroll
  for (int k = 0; k < KC; ++k) {
    const bool act = (active >> k) & 1u;
    fresh |= (act && old[k] == kEmptyKeyU) ? (1u << k) : 0u;
    pend |= (act && old[k] != kEmptyKeyU && old[k] != key[k]) ? (1u << k) : 0u;
  }
  if (pend) {  // collided first probes: rare lanes, short loops
#pragma unroll
    for (int k = 0; k < KC; ++k)
      if (pend & (1u << k)) {
        const uint32_t step = hash_step(key[k], log2cap) << 2;
        while (true) {
          slb[k] = (slb[k] + step) & maskb;
          const uint32_t o = atoms_cas(kaddr + slb[k], kEmptyKeyU, key[k]);
          if (o == kEmptyKeyU) {
            fresh |= 1u << k;
            break;
          }
          if (o == key[k]) break;
        }
      }
  }
#pragma unroll
  for (int k = 0; k < KC; ++k) reds_add_if((active >> k) & 1u, kaddr + capb + slb[k], w[k]);
  return fresh;
}

// split-gather variant: keys carry no singleton flag; the committing lane reads the winner's community size
// (only needed when the vertex itself is a singleton: louvain.cpp:2240-2241)
template <bool SPLIT>
__device__ __forceinline__ uint32_t winner_with_singleton_bit(const SweepU32Params &p, uint32_t me_key, uint32_t wid,
                                                              double wg) {
  if (SPLIT && wg > 0.0 && (me_key & kSingletonBit) && wid != kLabelMask && p.csize[wid & kLabelMask] == 1)
    wid |= kSingletonBit;
  return wid;
}

struct VertexHeader {
  vid_t v;
  eid_t e0, e1;
  uint2 me;
  double vd;
};

__device__ __forceinline__ VertexHeader fetch_header(const SweepU32Params &p, vid_t v) {
  VertexHeader h;
  h.v = v;
  h.e0 = p.offsets[v];
  h.e1 = p.offsets[v + 1];
  h.me = p.packed[self_of(p, v)];
  h.vd = p.vdeg[v];
  return h;
}

// K = KC * NCH entries per thread, processed in NCH chunks of KC; the table of a
// vertex has 2*K*TPV slots.  TPV < 32: 32/TPV vertices share a warp (lanes of a
// vertex are contiguous; every collective carries the vertex's lane mask), so
// the per-vertex bookkeeping instructions are issued once for several
// vertices; TPV >= 64: one block per vertex.
// TMA = true (block-per-vertex classes, CVB_TMA=1): the tails / weights tile of the next
// vertex is brought into shared memory by two cp.async.bulk copies signalled on an mbarrier
// while the current vertex is evaluated, and the insert phase reads it with LDS.  Measured on
// B200 (scale-20) this single-stage staging is ~10 % SLOWER than per-thread streaming loads
// (the copy is issued only after the previous tile is consumed, and the tile costs occupancy),
// so it is off by default; see DESIGN.md.
// SPLIT = true (CVB_SPLIT=1, A/B variant added at the end of round 1, not measured yet): the per-entry gather
// reads the 4-byte label array instead of the 8-byte `packed` record — half the L2 footprint, which is what
// decides the hit rate once nv_global * 8 B approaches the L2 size (scale 24; 8-GPU weak scaling) — and the
// community degree is gathered once per DISTINCT neighbour community at evaluation time (no cay[] registers
// during the insert phase); the singleton flag of the winner is read at commit time.
template <int TPV, int KC, int NCH, int MINB, bool TMA = false, bool SPLIT = false>
__global__ void __launch_bounds__((TPV <= 32 ? 256 : TPV), MINB)
    k_sweep_u32_hash(SweepU32Params p, const vid_t *__restrict__ list, int n) {
  extern __shared__ __align__(16) uint32_t smem_u32[];
  __shared__ double red_g[32];
  __shared__ uint32_t red_id[32];
  constexpr int kWarpsPerVertex = TPV / 32;
  constexpr int K = KC * NCH;
  constexpr int CAP = 2 * K * TPV;  // slots per vertex (power of two)
  constexpr int VPB = TPV <= 32 ? 256 / TPV : 1;
  constexpr int GL = TPV < 32 ? TPV : 32;  // lanes of one vertex inside a warp
  constexpr uint32_t capb = CAP * 4u;
  constexpr int kWords = CAP * 2 + (TPV > 32 ? 4 : 0);
  const unsigned lane = threadIdx.x & 31u;
  const unsigned gmask = (GL == 32) ? 0xffffffffu : (((1u << GL) - 1u) << (lane & ~static_cast<unsigned>(GL - 1)));
  const int tid = (TPV <= 32) ? static_cast<int>(threadIdx.x & (TPV - 1)) : static_cast<int>(threadIdx.x);
  const int vslot = (TPV <= 32) ? static_cast<int>(threadIdx.x / TPV) : 0;
  const int wiv = (TPV <= 32) ? 0 : static_cast<int>(threadIdx.x >> 5);  // warp in vertex
  uint32_t *keys = smem_u32 + kWords * vslot;
  uint32_t *wts = keys + CAP;
  uint32_t *meta = keys + 2 * CAP;  // only for TPV > 32
  const uint32_t kaddr = smem_addr(keys);
  // TMA-staged tile of the current vertex (block-per-vertex classes)
  constexpr bool kTma = TMA && TPV >= 128;
  constexpr int kTileWords = K * TPV + 4;  // +4: the row start is only 4-byte aligned, copies are 16-byte aligned
  uint32_t *tile_tails = meta + 4;
  uint32_t *tile_weights = tile_tails + kTileWords;
  const uint32_t mbar = smem_addr(tile_weights + kTileWords);
  uint32_t tile_parity = 0u;
  for (int s = tid; s < CAP; s += TPV) {
    keys[s] = kEmptyKeyU;
    wts[s] = 0u;
  }
  if (TPV > 32) {
    if (tid < 4) meta[tid] = 0u;
    if (kTma && tid == 0) {
      mbar_init(mbar, 1u);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
  } else {
    __syncwarp(gmask);
  }

  const int64_t stride = static_cast<int64_t>(gridDim.x) * VPB;
  int64_t item = static_cast<int64_t>(blockIdx.x) * VPB + vslot;
  if (item >= n) return;  // all lanes of a vertex leave together
  // two-deep software pipeline over the class list: the id of vertex i+2 and the
  // header (offsets, own packed entry, degree) of vertex i+1 are in flight
  // while vertex i is processed
  VertexHeader hc = fetch_header(p, list[item]);
  vid_t v_next = (item + stride < n) ? list[item + stride] : -1;
  // stage the tails/weights of [e0, e1): one elected thread arms the mbarrier and issues two bulk copies
  auto issue_tile = [&](eid_t e0, eid_t e1) {
    const eid_t a0 = e0 & ~static_cast<eid_t>(3);
    const uint32_t bytes = static_cast<uint32_t>(((e1 - a0) * 4 + 15) & ~static_cast<eid_t>(15));
    fence_proxy_async();  // generic-proxy reads of the previous tile are ordered before the async refill
    mbar_expect_tx(mbar, 2u * bytes);
    tma_load_1d(smem_addr(tile_tails), p.tails + a0, bytes, mbar);
    tma_load_1d(smem_addr(tile_weights), p.weights + a0, bytes, mbar);
  };
  if (kTma && tid == 0) issue_tile(hc.e0, hc.e1);
  for (; item < n; item += stride) {
    VertexHeader hn = hc;
    if (v_next >= 0) hn = fetch_header(p, v_next);
    const vid_t v_next2 = (item + 2 * stride < n) ? list[item + 2 * stride] : -1;

    const vid_t v = hc.v;
    const eid_t e0 = hc.e0, e1 = hc.e1;
    const int deg = static_cast<int>(e1 - e0);
    const vid_t self = self_of(p, v);
    const uint2 me = hc.me;
    const uint32_t cc = me.x & kLabelMask;
    const double vd = hc.vd;
    int log2cap = 32 - __clz(2 * deg - 1);  // ceil(log2(2*deg)): load factor <= 0.5
    log2cap = log2cap < 4 ? 4 : (log2cap > 31 - __clz(CAP) ? 31 - __clz(CAP) : log2cap);
    const uint32_t maskb = ((1u << log2cap) - 1u) << 2;

    const vid_t *__restrict__ row_tails = p.tails + e0;
    const uint32_t *__restrict__ row_weights = p.weights + e0;
    const int tile_off = static_cast<int>(e0 & 3);
    if (kTma) {
      mbar_wait(mbar, tile_parity);  // the tile of this vertex has landed
      tile_parity ^= 1u;
    }
    // claims of this lane: key, community degree and slot of every entry that
    // opened a new table slot (valid where the fresh bit is set)
    uint32_t ckey[K], cay[K];
    uint16_t cslot[K];
    uint32_t freshmask = 0u, ownsum = 0u, selfsum = 0u;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      {  // every class spans (K*TPV/2, K*TPV] entries, so all chunks are populated
        uint32_t key[KC], w[KC], slb[KC];
        vid_t tail[KC];
        uint32_t validm = 0u;
#pragma unroll
        for (int k = 0; k < KC; ++k) {
          const int i = tid + (c * KC + k) * TPV;  // 32-bit position inside the row
          tail[k] = self;  // a safe address for the gather of idle lanes
          w[k] = 0u;
          if (i < deg) {
            if (kTma) {
              tail[k] = static_cast<vid_t>(tile_tails[tile_off + i]);
              w[k] = tile_weights[tile_off + i];
            } else {
              tail[k] = __ldcs(&row_tails[i]);
              w[k] = __ldcs(&row_weights[i]);
            }
            validm |= 1u << k;
          }
        }
        uint32_t active = 0u;
#pragma unroll
        for (int k = 0; k < KC; ++k) {
          const bool valid = (validm >> k) & 1u;
          if (SPLIT) {
            const uint32_t lb = static_cast<uint32_t>(p.labels[tail[k]]);
            key[k] = valid ? lb : kEmptyKeyU;
          } else {
            const uint2 pk = p.packed[tail[k]];
            key[k] = valid ? pk.x : kEmptyKeyU;
            cay[c * KC + k] = pk.y;
          }
          const bool own = (key[k] & kLabelMask) == cc;  // idle lanes: 0x7fffffff is never a label
          ownsum += own ? w[k] : 0u;
          selfsum += (valid && tail[k] == self) ? w[k] : 0u;
          active |= (valid && !own) ? (1u << k) : 0u;
        }
        const uint32_t fresh = table_insert_chunk<KC>(kaddr, capb, key, w, active, log2cap, maskb, slb);
        freshmask |= fresh << (c * KC);
#pragma unroll
        for (int k = 0; k < KC; ++k) {
          ckey[c * KC + k] = key[k];
          cslot[c * KC + k] = static_cast<uint16_t>(slb[k] >> 2);
        }
      }
    }
    // own-community weight and self-loop weight of the vertex
    uint32_t c0u = ownsum, slu = selfsum;
    if (TPV > 1) {
      c0u = group_sum<GL>(ownsum, gmask);
      slu = group_sum<GL>(selfsum, gmask);
    }
    if (TPV > 32) {
      if (lane == 0) {
        if (c0u) atomicAdd(&meta[2], c0u);
        if (slu) atomicAdd(&meta[1], slu);
      }
      __syncthreads();
      c0u = meta[2];
      slu = meta[1];
      // every thread has consumed this vertex's tile: refill it for the next vertex while
      // this one is evaluated and committed
      if (kTma && tid == 0 && v_next >= 0) issue_tile(hn.e0, hn.e1);
    } else {
      __syncwarp(gmask);
    }

    // ---- evaluate: the claimer of each key reads its total and frees the slot ----
    const double counter0 = u32_to_f64(c0u);
    const double eix = counter0 - u32_to_f64(slu);
    const double ax = u32_to_f64(me.y) - vd;
    double g = -1.0;
    uint32_t id = kLabelMask;
    if (SPLIT) {  // community degrees of this lane's claims: all gathers in flight before the first use
#pragma unroll
      for (int q = 0; q < K; ++q) cay[q] = (freshmask & (1u << q)) ? p.cdeg32[ckey[q] & kLabelMask] : 0u;
    }
#pragma unroll
    for (int q = 0; q < K; ++q)
      if (freshmask & (1u << q)) {
        const uint32_t a = kaddr + (static_cast<uint32_t>(cslot[q]) << 2);
        const uint32_t tot = lds_u32(a + capb);
        sts_u32(a, kEmptyKeyU);
        sts_u32(a + capb, 0u);
        const double g2 = modularity_gain(u32_to_f64(tot), eix, vd, u32_to_f64(cay[q]), ax, p.constant);
        if (better_u(g2, ckey[q], g, id)) {
          g = g2;
          id = ckey[q];
        }
      }
    double wg = g;
    uint32_t wid = id;
    if (TPV > 1) wid = group_argmax<GL>(g, id, &wg, gmask);
    if (TPV <= 32) {
      if (tid == 0) commit_vertex_u32(p, v, me.x, wg, winner_with_singleton_bit<SPLIT>(p, me.x, wid, wg), vd, counter0);
      __syncwarp(gmask);
    } else {
      if (lane == 0) {
        red_g[wiv] = wg;
        red_id[wiv] = wid;
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        for (int i = 1; i < kWarpsPerVertex; ++i)
          if (better_u(red_g[i], red_id[i], wg, wid)) {
            wg = red_g[i];
            wid = red_id[i];
          }
        commit_vertex_u32(p, v, me.x, wg, winner_with_singleton_bit<SPLIT>(p, me.x, wid, wg), vd, counter0);
        meta[1] = 0u;
        meta[2] = 0u;
      }
      __syncthreads();
    }
    hc = hn;
    v_next = v_next2;
  }
}

// ---------------------------------------------------------------------------
// k_sweep_u32_hash2 — the round-2 revision of k_sweep_u32_hash (same work decomposition: K = KC*NCH
// register-resident claims per thread, two-deep header pipeline), cut where the round-2 profiles put the
// time: the L1TEX pipe (random gathers + random shared-memory accesses), not the issue slots.
//   * slots are interleaved {key, weight} 8-byte records: the evaluation frees a slot with ONE
//     st.shared.v2 instead of two stores;
//   * UNIT (every weight of the graph is 1): the weights array is never read, and a lane that opens a
//     slot adds nothing to it — the weight word only counts LATER entries of the same community
//     (total = 1 + count), so the common case is one CAS + one LDS + one STS per entry;
//   * on graphs whose rows hold each tail once (p.simple_rows), a neighbour whose community is a
//     singleton cannot share it with another entry of the row: it bypasses the table and is evaluated
//     with its own weight.
// Per table: CAP slots of 8 bytes, then meta[4] = {-, selfloop, counter0, -} for TPV > 32.
// ---------------------------------------------------------------------------
template <int TPV, int KC, int NCH>
__host__ __device__ constexpr size_t hash2_block_smem() {
  return (TPV <= 32) ? static_cast<size_t>(256 / TPV) * (2 * KC * NCH * TPV) * 8
                     : static_cast<size_t>(2 * KC * NCH * TPV) * 8 + 16;
}

__device__ __forceinline__ void sts64_u32(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}

template <int TPV, int KC, int NCH, int MINB, bool UNIT>
__global__ void __launch_bounds__((TPV <= 32 ? 256 : TPV), MINB)
    k_sweep_u32_hash2(SweepU32Params p, const vid_t *__restrict__ list, int n) {
  extern __shared__ __align__(16) uint32_t smem_u32[];
  __shared__ double red_g[32];
  __shared__ uint32_t red_id[32];
  constexpr int kWarpsPerVertex = TPV / 32;
  constexpr int K = KC * NCH;
  constexpr int CAP = 2 * K * TPV;  // slots per vertex (power of two)
  constexpr int VPB = TPV <= 32 ? 256 / TPV : 1;
  constexpr int GL = TPV < 32 ? TPV : 32;  // lanes of one vertex inside a warp
  constexpr int kWords = CAP * 2 + (TPV > 32 ? 4 : 0);
  const unsigned lane = threadIdx.x & 31u;
  const unsigned gmask = (GL == 32) ? 0xffffffffu : (((1u << GL) - 1u) << (lane & ~static_cast<unsigned>(GL - 1)));
  const int tid = (TPV <= 32) ? static_cast<int>(threadIdx.x & (TPV - 1)) : static_cast<int>(threadIdx.x);
  const int vslot = (TPV <= 32) ? static_cast<int>(threadIdx.x / TPV) : 0;
  const int wiv = (TPV <= 32) ? 0 : static_cast<int>(threadIdx.x >> 5);  // warp in vertex
  uint32_t *slots = smem_u32 + kWords * vslot;
  uint32_t *meta = slots + 2 * CAP;  // only for TPV > 32
  const uint32_t kaddr = smem_addr(slots);
  for (int s = tid; s < CAP; s += TPV) {
    slots[2 * s] = kEmptyKeyU;
    slots[2 * s + 1] = 0u;
  }
  if (TPV > 32) {
    if (tid < 4) meta[tid] = 0u;
    __syncthreads();
  } else {
    __syncwarp(gmask);
  }
  const bool bypass = p.simple_rows != 0;

  const int64_t stride = static_cast<int64_t>(gridDim.x) * VPB;
  int64_t item = static_cast<int64_t>(blockIdx.x) * VPB + vslot;
  if (item >= n) return;  // all lanes of a vertex leave together
  VertexHeader hc = fetch_header(p, list[item]);
  vid_t v_next = (item + stride < n) ? list[item + stride] : -1;
  for (; item < n; item += stride) {
    VertexHeader hn = hc;
    if (v_next >= 0) hn = fetch_header(p, v_next);
    const vid_t v_next2 = (item + 2 * stride < n) ? list[item + 2 * stride] : -1;

    const vid_t v = hc.v;
    const eid_t e0 = hc.e0, e1 = hc.e1;
    const int deg = static_cast<int>(e1 - e0);
    const vid_t self = self_of(p, v);
    const uint2 me = hc.me;
    const uint32_t cc = me.x & kLabelMask;
    const double vd = hc.vd;
    int log2cap = 32 - __clz(2 * deg - 1);  // ceil(log2(2*deg)): load factor <= 0.5
    log2cap = log2cap < 4 ? 4 : (log2cap > 31 - __clz(CAP) ? 31 - __clz(CAP) : log2cap);
    const uint32_t maskb = ((1u << log2cap) - 1u) << 3;

    const vid_t *__restrict__ row_tails = p.tails + e0;
    const uint32_t *__restrict__ row_weights = p.weights + e0;
    // claims of this lane: key, community degree and slot of every entry that opened a table slot
    // (freshmask) or bypassed the table (directmask)
    uint32_t ckey[K], cay[K], cw_[K];
    uint16_t cslot[K];
    uint32_t freshmask = 0u, directmask = 0u, ownsum = 0u, selfsum = 0u;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      uint32_t key[KC], w[KC], slb[KC], old[KC];
      vid_t tail[KC];
      uint32_t validm = 0u;
#pragma unroll
      for (int k = 0; k < KC; ++k) {
        const int i = tid + (c * KC + k) * TPV;  // 32-bit position inside the row
        tail[k] = self;  // a safe address for the gather of idle lanes
        w[k] = UNIT ? 1u : 0u;
        if (i < deg) {
          tail[k] = __ldcs(&row_tails[i]);
          if (!UNIT) w[k] = __ldcs(&row_weights[i]);
          validm |= 1u << k;
        }
      }
      uint32_t active = 0u;
#pragma unroll
      for (int k = 0; k < KC; ++k) {
        const bool valid = (validm >> k) & 1u;
        const uint2 pk = p.packed[tail[k]];
        key[k] = valid ? pk.x : kEmptyKeyU;
        cay[c * KC + k] = pk.y;
        cw_[c * KC + k] = w[k];
        const bool own = (key[k] & kLabelMask) == cc;  // idle lanes: 0x7fffffff is never a label
        ownsum += (valid && own) ? w[k] : 0u;
        selfsum += (valid && tail[k] == self) ? w[k] : 0u;
        const bool direct = bypass && valid && !own && (pk.x & kSingletonBit);
        directmask |= direct ? (1u << (c * KC + k)) : 0u;
        active |= (valid && !own && !direct) ? (1u << k) : 0u;
      }
      // first probes of the chunk back to back, collided entries walk their own short loops (double hashing)
      uint32_t fresh = 0u, pend = 0u;
#pragma unroll
      for (int k = 0; k < KC; ++k) {
        slb[k] = hash_key(key[k], log2cap) << 3;
        old[k] = atoms_cas_if((active >> k) & 1u, kaddr + slb[k], kEmptyKeyU, key[k], key[k]);
      }
#pragma unroll
      for (int k = 0; k < KC; ++k) {
        const bool act = (active >> k) & 1u;
        fresh |= (act && old[k] == kEmptyKeyU) ? (1u << k) : 0u;
        pend |= (act && old[k] != kEmptyKeyU && old[k] != key[k]) ? (1u << k) : 0u;
      }
      if (pend) {
#pragma unroll
        for (int k = 0; k < KC; ++k)
          if (pend & (1u << k)) {
            const uint32_t step = hash_step(key[k], log2cap) << 3;
            while (true) {
              slb[k] = (slb[k] + step) & maskb;
              const uint32_t o = atoms_cas(kaddr + slb[k], kEmptyKeyU, key[k]);
              if (o == kEmptyKeyU) {
                fresh |= 1u << k;
                break;
              }
              if (o == key[k]) break;
            }
          }
      }
      // unit weights: the slot counts the entries AFTER the claim; weighted: every entry adds
#pragma unroll
      for (int k = 0; k < KC; ++k)
        reds_add_if(((active >> k) & 1u) && !(UNIT && ((fresh >> k) & 1u)), kaddr + slb[k] + 4u, w[k]);
      freshmask |= fresh << (c * KC);
#pragma unroll
      for (int k = 0; k < KC; ++k) {
        ckey[c * KC + k] = key[k];
        cslot[c * KC + k] = static_cast<uint16_t>(slb[k] >> 3);
      }
    }
    // own-community weight and self-loop weight of the vertex
    uint32_t c0u = ownsum, slu = selfsum;
    if (TPV > 1) {
      c0u = group_sum<GL>(ownsum, gmask);
      slu = group_sum<GL>(selfsum, gmask);
    }
    if (TPV > 32) {
      if (lane == 0) {
        if (c0u) atomicAdd(&meta[2], c0u);
        if (slu) atomicAdd(&meta[1], slu);
      }
      __syncthreads();
      c0u = meta[2];
      slu = meta[1];
    } else {
      __syncwarp(gmask);
    }

    // ---- evaluate: the claimer of each key reads its total and frees the slot ----
    const double counter0 = u32_to_f64(c0u);
    const double eix = counter0 - u32_to_f64(slu);
    const double ax = u32_to_f64(me.y) - vd;
    double g = -1.0;
    uint32_t id = kLabelMask;
#pragma unroll
    for (int q = 0; q < K; ++q)
      if ((freshmask | directmask) & (1u << q)) {
        uint32_t tot = cw_[q];
        if (freshmask & (1u << q)) {
          const uint32_t a = kaddr + (static_cast<uint32_t>(cslot[q]) << 3);
          tot = lds_u32(a + 4u) + (UNIT ? 1u : 0u);
          sts64_u32(a, kEmptyKeyU, 0u);
        }
        const double g2 = modularity_gain(u32_to_f64(tot), eix, vd, u32_to_f64(cay[q]), ax, p.constant);
        if (better_u(g2, ckey[q], g, id)) {
          g = g2;
          id = ckey[q];
        }
      }
    double wg = g;
    uint32_t wid = id;
    if (TPV > 1) wid = group_argmax<GL>(g, id, &wg, gmask);
    if (TPV <= 32) {
      if (tid == 0) commit_vertex_u32(p, v, me.x, wg, wid, vd, counter0);
      __syncwarp(gmask);
    } else {
      if (lane == 0) {
        red_g[wiv] = wg;
        red_id[wiv] = wid;
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        for (int i = 1; i < kWarpsPerVertex; ++i)
          if (better_u(red_g[i], red_id[i], wg, wid)) {
            wg = red_g[i];
            wid = red_id[i];
          }
        commit_vertex_u32(p, v, me.x, wg, wid, vd, counter0);
        meta[1] = 0u;
        meta[2] = 0u;
      }
      __syncthreads();
    }
    hc = hn;
    v_next = v_next2;
  }
}

// ---------------------------------------------------------------------------
// deg 8193..65536: one THREAD-BLOCK CLUSTER per vertex (2, 4 or 8 CTAs on as many
// SMs), the hash table distributed over the cluster's shared memories (DSMEM).
// CTA r streams entries [r*8192, (r+1)*8192) of the adjacency; a key's slot
// lives in the CTA named by the top bits of its hash and is claimed / added to
// with shared::cluster atomics, so the whole accumulation stays on chip and no
// global table, no L2 atomic and no second pass over memory is needed.
// Per CTA: keys[16384] wts[16384] (128 KB), meta[4]; rank 0 also gathers the
// per-CTA candidates.
// ---------------------------------------------------------------------------
constexpr int kClusterSliceLog2 = 14;  // 16384 slots per CTA
constexpr int kClusterSlice = 1 << kClusterSliceLog2;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_nclusters_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t d;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(d) : "r"(addr), "r"(rank));
  return d;
}
__device__ __forceinline__ uint32_t atomc_cas(uint32_t addr, uint32_t cmp, uint32_t val) {
  uint32_t old;
  asm volatile("atom.shared::cluster.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "r"(addr), "r"(cmp), "r"(val) : "memory");
  return old;
}
__device__ __forceinline__ void redc_add(uint32_t addr, uint32_t v) {
  asm volatile("red.shared::cluster.add.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ldc_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void stc_u32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void stc_f64(uint32_t addr, double v) {
  asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int CS>
__global__ void __launch_bounds__(1024, 1) k_sweep_u32_cluster(SweepU32Params p, const vid_t *__restrict__ list, int n) {
  constexpr int TPB = 1024, KC = 4, NCH = 2, K = KC * NCH;
  constexpr int kLog2Cap = kClusterSliceLog2 + (CS == 2 ? 1 : CS == 4 ? 2 : 3);
  constexpr uint32_t capb = kClusterSlice * 4u;          // bytes of one keys[] slice
  constexpr uint32_t maskw = (1u << kLog2Cap) - 1u;      // slot mask over the whole cluster
  extern __shared__ __align__(16) uint32_t smem_u32[];
  __shared__ double red_g[32];
  __shared__ uint32_t red_id[32];
  __shared__ double cl_g[8];     // rank 0: one candidate per CTA
  __shared__ uint32_t cl_id[8];
  uint32_t *keys = smem_u32, *wts = smem_u32 + kClusterSlice, *meta = smem_u32 + 2 * kClusterSlice;
  const uint32_t kaddr = smem_addr(keys);
  const uint32_t maddr = smem_addr(meta);
  const uint32_t rank = cluster_ctarank();
  const unsigned lane = threadIdx.x & 31u;
  const int wib = threadIdx.x >> 5;
  const int tid = threadIdx.x;
  for (int s = tid; s < kClusterSlice; s += TPB) {
    keys[s] = kEmptyKeyU;
    wts[s] = 0u;
  }
  if (tid < 4) meta[tid] = 0u;
  cluster_sync_all();
  const uint32_t maddr0 = mapa_u32(maddr, 0);  // rank 0's meta: {-, selfloop, counter0, -}

  for (int64_t item = cluster_id_x(); item < n; item += cluster_nclusters_x()) {
    const vid_t v = list[item];
    const eid_t e0 = p.offsets[v], e1 = p.offsets[v + 1];
    const int deg = static_cast<int>(e1 - e0);
    const vid_t self = self_of(p, v);
    const uint2 me = p.packed[self];
    const uint32_t cc = me.x & kLabelMask;
    const double vd = p.vdeg[v];
    const vid_t *__restrict__ row_tails = p.tails + e0;
    const uint32_t *__restrict__ row_weights = p.weights + e0;
    const int i0 = static_cast<int>(rank) * (TPB * K);  // this CTA's window of the row

    uint32_t ckey[K], cay[K], caddr[K];
    uint32_t freshmask = 0u, ownsum = 0u, selfsum = 0u;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      uint32_t key[KC], w[KC];
      vid_t tail[KC];
      uint32_t validm = 0u;
#pragma unroll
      for (int k = 0; k < KC; ++k) {
        const int i = i0 + tid + (c * KC + k) * TPB;
        tail[k] = self;
        w[k] = 0u;
        if (i < deg) {
          tail[k] = __ldcs(&row_tails[i]);
          w[k] = __ldcs(&row_weights[i]);
          validm |= 1u << k;
        }
      }
      uint32_t active = 0u;
#pragma unroll
      for (int k = 0; k < KC; ++k) {
        const uint2 pk = p.packed[tail[k]];
        const bool valid = (validm >> k) & 1u;
        key[k] = valid ? pk.x : kEmptyKeyU;
        cay[c * KC + k] = pk.y;
        const bool own = (key[k] & kLabelMask) == cc;
        ownsum += own ? w[k] : 0u;
        selfsum += (valid && tail[k] == self) ? w[k] : 0u;
        active |= (valid && !own) ? (1u << k) : 0u;
      }
      // claim: all first probes in flight, then round-outer retries (one DSMEM round trip per round)
      uint32_t sl[KC], old[KC], addr[KC];
      uint32_t pend = 0u, fresh = 0u;
#pragma unroll
      for (int k = 0; k < KC; ++k) {
        sl[k] = hash_key(key[k], kLog2Cap);
        addr[k] = mapa_u32(kaddr, sl[k] >> kClusterSliceLog2) + ((sl[k] & (kClusterSlice - 1)) << 2);
        old[k] = key[k];
        if (active & (1u << k)) old[k] = atomc_cas(addr[k], kEmptyKeyU, key[k]);
      }
#pragma unroll
      for (int k = 0; k < KC; ++k) {
        const bool act = (active >> k) & 1u;
        fresh |= (act && old[k] == kEmptyKeyU) ? (1u << k) : 0u;
        pend |= (act && old[k] != kEmptyKeyU && old[k] != key[k]) ? (1u << k) : 0u;
      }
      while (pend) {
#pragma unroll
        for (int k = 0; k < KC; ++k)
          if (pend & (1u << k)) {
            sl[k] = (sl[k] + hash_step(key[k], kLog2Cap)) & maskw;
            addr[k] = mapa_u32(kaddr, sl[k] >> kClusterSliceLog2) + ((sl[k] & (kClusterSlice - 1)) << 2);
            old[k] = atomc_cas(addr[k], kEmptyKeyU, key[k]);
          }
#pragma unroll
        for (int k = 0; k < KC; ++k)
          if (pend & (1u << k)) {
            if (old[k] == kEmptyKeyU) {
              fresh |= 1u << k;
              pend &= ~(1u << k);
            } else if (old[k] == key[k]) {
              pend &= ~(1u << k);
            }
          }
      }
#pragma unroll
      for (int k = 0; k < KC; ++k) {
        if (active & (1u << k)) redc_add(addr[k] + capb, w[k]);
        ckey[c * KC + k] = key[k];
        caddr[c * KC + k] = addr[k];
      }
      freshmask |= fresh << (c * KC);
    }
    {
      const uint32_t c0u = __reduce_add_sync(0xffffffffu, ownsum);
      const uint32_t slu = __reduce_add_sync(0xffffffffu, selfsum);
      if (lane == 0) {
        if (c0u) redc_add(maddr0 + 8, c0u);
        if (slu) redc_add(maddr0 + 4, slu);
      }
    }
    cluster_sync_all();  // every insert of every CTA has landed

    const double counter0 = u32_to_f64(ldc_u32(maddr0 + 8));
    const double eix = counter0 - u32_to_f64(ldc_u32(maddr0 + 4));
    const double ax = u32_to_f64(me.y) - vd;
    double g = -1.0;
    uint32_t id = kLabelMask;
#pragma unroll
    for (int q = 0; q < K; ++q)
      if (freshmask & (1u << q)) {
        const uint32_t tot = ldc_u32(caddr[q] + capb);
        stc_u32(caddr[q], kEmptyKeyU);
        stc_u32(caddr[q] + capb, 0u);
        const double g2 = modularity_gain(u32_to_f64(tot), eix, vd, u32_to_f64(cay[q]), ax, p.constant);
        if (better_u(g2, ckey[q], g, id)) {
          g = g2;
          id = ckey[q];
        }
      }
    double wg;
    uint32_t wid = warp_argmax(g, id, &wg);
    if (lane == 0) {
      red_g[wib] = wg;
      red_id[wib] = wid;
    }
    __syncthreads();
    if (tid == 0) {
      for (int i = 1; i < TPB / 32; ++i)
        if (better_u(red_g[i], red_id[i], wg, wid)) {
          wg = red_g[i];
          wid = red_id[i];
        }
      stc_f64(mapa_u32(smem_addr(&cl_g[rank]), 0), wg);
      stc_u32(mapa_u32(smem_addr(&cl_id[rank]), 0), wid);
    }
    cluster_sync_all();  // candidates and slot clears visible cluster-wide
    if (rank == 0 && tid == 0) {
      wg = cl_g[0];
      wid = cl_id[0];
      for (int r = 1; r < CS; ++r)
        if (better_u(cl_g[r], cl_id[r], wg, wid)) {
          wg = cl_g[r];
          wid = cl_id[r];
        }
      commit_vertex_u32(p, v, me.x, wg, wid, vd, counter0);
      meta[1] = 0u;
      meta[2] = 0u;
    }
    cluster_sync_all();  // meta reset before the next vertex's sums arrive
  }
}

// ---------------------------------------------------------------------------
// deg 8193..16384: one block per vertex, the adjacency is read TWICE and each pass
// accumulates the neighbour communities of one half of the key space (a mixing
// hash of the community id picks the half) in a 16384-slot shared-memory table.
// The second read of the row (<= 128 KB) comes from L2; no global table, no L2
// atomic, no second kernel.  A table can never overflow: a row has at most 16384
// entries, hence at most 16384 distinct communities in total.
// After each pass the table is scanned (16 slots per thread), every community
// found is evaluated and its slot freed.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t key_half(uint32_t key) {
  uint32_t m = (key & kLabelMask) * 0xC2B2AE35u;
  m ^= m >> 15;
  m *= 0x27D4EB2Fu;
  return m >> 31;
}

__global__ void __launch_bounds__(1024, 1) k_sweep_u32_twopass(SweepU32Params p, const vid_t *__restrict__ list, int n) {
  constexpr int TPB = 1024, KC = 4, CAP = 16384, kLog2Cap = 14;
  constexpr uint32_t capb = CAP * 4u, maskb = (CAP - 1u) << 2;
  extern __shared__ __align__(16) uint32_t smem_u32[];
  __shared__ double red_g[32];
  __shared__ uint32_t red_id[32];
  __shared__ uint32_t meta[4];
  uint32_t *keys = smem_u32, *wts = smem_u32 + CAP;
  const uint32_t kaddr = smem_addr(keys);
  const unsigned lane = threadIdx.x & 31u;
  const int wib = threadIdx.x >> 5;
  const int tid = threadIdx.x;
  for (int s = tid; s < CAP; s += TPB) {
    keys[s] = kEmptyKeyU;
    wts[s] = 0u;
  }
  if (tid < 4) meta[tid] = 0u;
  __syncthreads();
  for (int64_t item = blockIdx.x; item < n; item += gridDim.x) {
    const vid_t v = list[item];
    const eid_t e0 = p.offsets[v];
    const int deg = static_cast<int>(p.offsets[v + 1] - e0);
    const vid_t self = self_of(p, v);
    const uint2 me = p.packed[self];
    const uint32_t cc = me.x & kLabelMask;
    const double vd = p.vdeg[v];
    const double ax = u32_to_f64(me.y) - vd;
    const vid_t *__restrict__ row_tails = p.tails + e0;
    const uint32_t *__restrict__ row_weights = p.weights + e0;
    double g = -1.0, eix = 0.0, counter0 = 0.0;
    uint32_t id = kLabelMask;
    for (uint32_t half = 0; half < 2u; ++half) {
      uint32_t ownsum = 0u, selfsum = 0u;
      for (int i0 = 0; i0 < deg; i0 += TPB * KC) {
        uint32_t key[KC], w[KC], slb[KC];
        vid_t tail[KC];
        uint32_t validm = 0u, active = 0u;
#pragma unroll
        for (int k = 0; k < KC; ++k) {
          const int i = i0 + tid + k * TPB;
          tail[k] = self;
          w[k] = 0u;
          if (i < deg) {
            // first pass: default loads leave the row in L2 for the second pass, which streams it out
            tail[k] = half ? __ldcs(&row_tails[i]) : __ldg(&row_tails[i]);
            w[k] = half ? __ldcs(&row_weights[i]) : __ldg(&row_weights[i]);
            validm |= 1u << k;
          }
        }
#pragma unroll
        for (int k = 0; k < KC; ++k) {
          const uint2 pk = p.packed[tail[k]];
          const bool valid = (validm >> k) & 1u;
          key[k] = valid ? pk.x : kEmptyKeyU;
          const bool own = (key[k] & kLabelMask) == cc;
          ownsum += own ? w[k] : 0u;
          selfsum += (valid && tail[k] == self) ? w[k] : 0u;
          active |= (valid && !own && key_half(key[k]) == half) ? (1u << k) : 0u;
        }
        if (__any_sync(0xffffffffu, active != 0u))
          (void)table_insert_chunk<KC>(kaddr, capb, key, w, active, kLog2Cap, maskb, slb);
      }
      if (half == 0u) {  // own-community and self-loop weight: counted once
        const uint32_t c0u = __reduce_add_sync(0xffffffffu, ownsum);
        const uint32_t slu = __reduce_add_sync(0xffffffffu, selfsum);
        if (lane == 0) {
          if (c0u) atomicAdd(&meta[2], c0u);
          if (slu) atomicAdd(&meta[1], slu);
        }
      }
      __syncthreads();
      if (half == 0u) {
        counter0 = u32_to_f64(meta[2]);
        eix = counter0 - u32_to_f64(meta[1]);
      }
      // evaluate and free this half's table: 16 slots per thread, 4 records in flight
#pragma unroll
      for (int j0 = 0; j0 < CAP / TPB; j0 += 4) {
        uint32_t rk[4], rw[4];
        double ayv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t a = kaddr + (static_cast<uint32_t>(tid + (j0 + j) * TPB) << 2);
          rk[j] = lds_u32(a);
          rw[j] = lds_u32(a + capb);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          ayv[j] = 0.0;
          if (rk[j] != kEmptyKeyU) {
            ayv[j] = p.cdeg[rk[j] & kLabelMask];
            const uint32_t a = kaddr + (static_cast<uint32_t>(tid + (j0 + j) * TPB) << 2);
            sts_u32(a, kEmptyKeyU);
            sts_u32(a + capb, 0u);
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (rk[j] != kEmptyKeyU) {
            const double g2 = modularity_gain(u32_to_f64(rw[j]), eix, vd, ayv[j], ax, p.constant);
            if (better_u(g2, rk[j], g, id)) {
              g = g2;
              id = rk[j];
            }
          }
      }
      __syncthreads();  // the table is clean before the next pass / vertex inserts
    }
    double wg;
    uint32_t wid = warp_argmax(g, id, &wg);
    if (lane == 0) {
      red_g[wib] = wg;
      red_id[wib] = wid;
    }
    __syncthreads();
    if (tid == 0) {
      for (int i = 1; i < TPB / 32; ++i)
        if (better_u(red_g[i], red_id[i], wg, wid)) {
          wg = red_g[i];
          wid = red_id[i];
        }
      commit_vertex_u32(p, v, me.x, wg, wid, vd, counter0);
      meta[1] = 0u;
      meta[2] = 0u;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------
// deg > 8192: tables in global memory, three kernels.
//   k_sweep_u32_giant_insert : blocks take 4096-entry chunks of giant
//       adjacencies (chunk -> vertex by binary search in chunk_start) and
//       insert into that vertex's table with L2 atomics;
//   k_sweep_u32_giant_eval   : blocks take 8192-slot pieces of the tables,
//       evaluate the gains of the keys found there, free the slots and leave one
//       (gain, key) candidate per piece;
//   k_sweep_u32_giant_final  : one warp per giant vertex reduces its pieces'
//       candidates and commits.
// gmeta[3*i + {1,2}] = {selfloop, counter0} of giant i.
// ---------------------------------------------------------------------------
constexpr int kGiantChunk = 4096;
constexpr int kGiantPiece = 8192;  // table capacities are multiples of this

struct GiantTables {
  const int32_t *__restrict__ chunk_owner;  // [total_chunks] giant index of every chunk
  const int64_t *__restrict__ chunk_start;  // [ngiant+1] prefix of chunks per vertex
  const int64_t *__restrict__ table_start;  // [ngiant+1] prefix of table slots per vertex
  uint2 *__restrict__ slots;  // {key, weight} arena shared by the batches: slot s of the layout lives at slots[s - slot_base]
  int64_t slot_base;          // first layout slot of the batch being processed
  // dense-accumulator variant (default): vertex gi of the batch adds into
  // dense[(gi - vert_lo) * dense_stride + community] and appends the communities it met to
  // klist[chunk_start[gi] * kGiantChunk ...], kcount[gi] entries
  uint32_t *__restrict__ dense;
  int64_t dense_stride;
  int vert_lo;
  uint32_t *__restrict__ kcount;
  uint32_t *__restrict__ klist;
  uint32_t *__restrict__ gmeta;
  double *__restrict__ piece_gain;   // [total_slots / kGiantPiece]
  uint32_t *__restrict__ piece_key;  // [total_slots / kGiantPiece]
};

// One warp (<<<1, 32>>>): exclusive prefix sums of the chunk and table-slot counts of the listed vertices.
__global__ void k_giant_layout(const eid_t *__restrict__ offsets, const vid_t *__restrict__ list, int n,
                               int64_t *__restrict__ chunk_start, int64_t *__restrict__ table_start) {
  if (blockIdx.x != 0 || threadIdx.x >= 32) return;
  const unsigned lane = threadIdx.x;
  int64_t cbase = 0, sbase = 0;
  for (int i0 = 0; i0 < n; i0 += 32) {
    const int i = i0 + static_cast<int>(lane);
    int64_t dc = 0, ds = 0;
    if (i < n) {
      const int64_t deg = offsets[list[i] + 1] - offsets[list[i]];
      dc = (deg + kGiantChunk - 1) / kGiantChunk;
      int64_t cap = kGiantPiece;
      while (2 * cap < 3 * deg) cap <<= 1;  // load factor <= 2/3 (double hashing)
      ds = cap;
    }
    int64_t pc = dc, ps = ds;  // inclusive warp scan
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int64_t tc = __shfl_up_sync(0xffffffffu, pc, d), ts = __shfl_up_sync(0xffffffffu, ps, d);
      if (lane >= static_cast<unsigned>(d)) {
        pc += tc;
        ps += ts;
      }
    }
    if (i < n) {
      chunk_start[i] = cbase + pc - dc;
      table_start[i] = sbase + ps - ds;
    }
    cbase += __shfl_sync(0xffffffffu, pc, 31);
    sbase += __shfl_sync(0xffffffffu, ps, 31);
  }
  if (lane == 0) {
    chunk_start[n] = cbase;
    table_start[n] = sbase;
  }
}

__global__ void k_giant_init_slots(uint2 *__restrict__ slots, int64_t n) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    slots[i] = make_uint2(kEmptyKeyU, 0u);
}

__global__ void k_giant_chunk_owner(const int64_t *__restrict__ chunk_start, int n, int32_t *__restrict__ owner) {
  const int gi = blockIdx.x;
  if (gi >= n) return;
  for (int64_t c = chunk_start[gi] + threadIdx.x; c < chunk_start[gi + 1]; c += blockDim.x) owner[c] = gi;
}

__device__ __forceinline__ int giant_owner(const int64_t *__restrict__ start, int n, int64_t x) {
  int lo = 0, hi = n;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (start[mid] <= x) lo = mid; else hi = mid;
  }
  return lo;
}

// Global-table twin of table_insert_chunk: K independent L2 atomics in flight,
// collided entries retry in their own short loops.
template <int K>
__device__ __forceinline__ void gtable_insert_chunk(uint2 *__restrict__ slots, const uint32_t (&key)[K],
                                                    const uint32_t (&w)[K], uint32_t active, int log2cap,
                                                    uint32_t mask) {
  uint32_t old[K], sl[K];
  uint32_t fresh = 0u, pend = 0u;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    sl[k] = hash_key(key[k], log2cap);
    old[k] = key[k];
    if (active & (1u << k)) old[k] = atomicCAS(&slots[sl[k]].x, kEmptyKeyU, key[k]);
  }
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const bool act = (active >> k) & 1u;
    fresh |= (act && old[k] == kEmptyKeyU) ? (1u << k) : 0u;
    pend |= (act && old[k] != kEmptyKeyU && old[k] != key[k]) ? (1u << k) : 0u;
  }
  // collided entries: every round issues ALL outstanding claims back to back, so a
  // round costs one L2 round trip however many entries are still pending
  while (pend) {
#pragma unroll
    for (int k = 0; k < K; ++k)
      if (pend & (1u << k)) {
        sl[k] = (sl[k] + hash_step(key[k], log2cap)) & mask;
        old[k] = atomicCAS(&slots[sl[k]].x, kEmptyKeyU, key[k]);
      }
#pragma unroll
    for (int k = 0; k < K; ++k)
      if (pend & (1u << k)) {
        if (old[k] == kEmptyKeyU) {
          fresh |= 1u << k;
          pend &= ~(1u << k);
        } else if (old[k] == key[k]) {
          pend &= ~(1u << k);
        }
      }
  }
#pragma unroll
  for (int k = 0; k < K; ++k)
    if (active & (1u << k)) atomicAdd(&slots[sl[k]].y, w[k]);
  (void)fresh;
}

// One 4096-entry chunk per block iteration: the chunk is first merged in a
// shared-memory table (cheap atomics; duplicates collapse), then each claiming
// lane flushes its distinct (community, weight, degree) triples into the
// vertex's global table — so the number of L2 atomics is the number of distinct
// communities per chunk, and they are issued with 8 independent ones in flight.
template <bool kDense>
__global__ void __launch_bounds__(512, 2) k_sweep_u32_giant_insert(SweepU32Params p, const vid_t *__restrict__ list,
                                                                  int n, GiantTables gt, int64_t chunk_lo,
                                                                  int64_t chunk_hi) {
  constexpr int TPV = 512, KC = 4, NCH = kGiantChunk / (TPV * KC), K = KC * NCH;
  constexpr int CAP = 2 * kGiantChunk;
  constexpr uint32_t capb = CAP * 4u;
  extern __shared__ __align__(16) uint32_t smem_u32[];
  uint32_t *skeys = smem_u32, *swts = smem_u32 + CAP;
  const uint32_t kaddr = smem_addr(skeys);
  const unsigned lane = threadIdx.x & 31u;
  const int tid = threadIdx.x;
  for (int s = tid; s < CAP; s += TPV) {
    skeys[s] = kEmptyKeyU;
    swts[s] = 0u;
  }
  __syncthreads();
  constexpr int log2cap_s = 13;  // CAP = 8192
  constexpr uint32_t maskb = (CAP - 1u) << 2;
  for (int64_t chunk = chunk_lo + blockIdx.x; chunk < chunk_hi; chunk += gridDim.x) {
    const int gi = gt.chunk_owner[chunk];
    const vid_t v = list[gi];
    const eid_t ve0 = p.offsets[v], ve1 = p.offsets[v + 1];
    const eid_t e0 = ve0 + (chunk - gt.chunk_start[gi]) * kGiantChunk;
    const eid_t e1 = (e0 + kGiantChunk < ve1) ? e0 + kGiantChunk : ve1;
    const vid_t self = self_of(p, v);
    const uint32_t cc = p.packed[self].x & kLabelMask;
    const int64_t tbase = gt.table_start[gi];
    const int64_t gcap = gt.table_start[gi + 1] - tbase;
    const int glog2 = 63 - __clzll(gcap);
    const uint32_t gmask = static_cast<uint32_t>(gcap - 1);
    uint32_t *meta = gt.gmeta + 3 * static_cast<int64_t>(gi);

    uint32_t ckey[K], cay[K];
    uint16_t cslot[K];
    uint32_t freshmask = 0u, ownsum = 0u, selfsum = 0u;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      uint32_t key[KC], w[KC], slb[KC];
      vid_t tail[KC];
      uint32_t validm = 0u;
#pragma unroll
      for (int k = 0; k < KC; ++k) {
        const eid_t e = e0 + tid + static_cast<eid_t>(c * KC + k) * TPV;
        tail[k] = self;
        w[k] = 0u;
        if (e < e1) {
          tail[k] = __ldcs(&p.tails[e]);
          w[k] = __ldcs(&p.weights[e]);
          validm |= 1u << k;
        }
      }
      uint32_t active = 0u;
#pragma unroll
      for (int k = 0; k < KC; ++k) {
        const uint2 pk = p.packed[tail[k]];
        const bool valid = (validm >> k) & 1u;
        key[k] = valid ? pk.x : kEmptyKeyU;
        cay[c * KC + k] = pk.y;
        const bool own = (key[k] & kLabelMask) == cc;
        ownsum += own ? w[k] : 0u;
        selfsum += (valid && tail[k] == self) ? w[k] : 0u;
        active |= (valid && !own) ? (1u << k) : 0u;
      }
      const uint32_t fresh = table_insert_chunk<KC>(kaddr, capb, key, w, active, log2cap_s, maskb, slb);
      freshmask |= fresh << (c * KC);
#pragma unroll
      for (int k = 0; k < KC; ++k) {
        ckey[c * KC + k] = key[k];
        cslot[c * KC + k] = static_cast<uint16_t>(slb[k] >> 2);
      }
    }
    const uint32_t c0u = __reduce_add_sync(0xffffffffu, ownsum);
    const uint32_t slu = __reduce_add_sync(0xffffffffu, selfsum);
    if (lane == 0) {
      if (c0u) atomicAdd(&meta[2], c0u);
      if (slu) atomicAdd(&meta[1], slu);
    }
    __syncthreads();
    // flush this lane's claims: chunk totals -> global table
    uint32_t tot[K];
#pragma unroll
    for (int q = 0; q < K; ++q) {
      tot[q] = 0u;
      if (freshmask & (1u << q)) {
        const uint32_t a = kaddr + (static_cast<uint32_t>(cslot[q]) << 2);
        tot[q] = lds_u32(a + capb);
        sts_u32(a, kEmptyKeyU);
        sts_u32(a + capb, 0u);
      }
    }
    if (kDense) {
      // chunk totals -> the vertex's dense accumulator with fire-and-forget L2 reductions (no
      // probing, no return value to wait for); the communities this chunk met are appended to
      // the vertex's list — one cursor atomic per warp
      uint32_t *__restrict__ dense = gt.dense + static_cast<int64_t>(gi - gt.vert_lo) * gt.dense_stride;
      const uint32_t mine = __popc(freshmask);
      uint32_t incl = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= static_cast<unsigned>(o)) incl += t;
      }
      const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
      uint32_t base = 0u;
      if (lane == 0 && total) base = atomicAdd(&gt.kcount[gi], total);
      base = __shfl_sync(0xffffffffu, base, 0);
      uint32_t *__restrict__ out = gt.klist + gt.chunk_start[gi] * kGiantChunk + base + (incl - mine);
#pragma unroll
      for (int q = 0; q < K; ++q)
        if (freshmask & (1u << q)) {
          atomicAdd(&dense[ckey[q] & kLabelMask], tot[q]);
          *out++ = ckey[q];
        }
    } else {
      gtable_insert_chunk<K>(gt.slots + (tbase - gt.slot_base), ckey, tot, freshmask, glog2, gmask);
    }
    __syncthreads();
  }
}

// Dense-accumulator evaluation: block c takes list positions [4096*(c - first chunk), +4096)
// of its vertex; the first reader of a community's total (atomicExch -> 0) evaluates it, later
// duplicates of the same community (met in another chunk) read 0 and skip — which also leaves
// the accumulator clean for the next batch.  One (gain, key) candidate per chunk.
__global__ void __launch_bounds__(512) k_sweep_u32_giant_eval_dense(SweepU32Params p, const vid_t *__restrict__ list,
                                                                   int n, GiantTables gt, int64_t chunk_lo,
                                                                   int64_t chunk_hi, double *__restrict__ chunk_gain,
                                                                   uint32_t *__restrict__ chunk_key) {
  __shared__ double red_g[16];
  __shared__ uint32_t red_id[16];
  const unsigned lane = threadIdx.x & 31u;
  const int wib = threadIdx.x >> 5;
  for (int64_t chunk = chunk_lo + blockIdx.x; chunk < chunk_hi; chunk += gridDim.x) {
    const int gi = gt.chunk_owner[chunk];
    const int64_t first = gt.chunk_start[gi];
    const uint32_t cnt = __ldcg(&gt.kcount[gi]);
    const uint32_t off = static_cast<uint32_t>(chunk - first) * kGiantChunk;
    double g = -1.0;
    uint32_t id = kLabelMask;
    if (off < cnt) {  // block-uniform
      const vid_t v = list[gi];
      const uint2 me = p.packed[self_of(p, v)];
      const double vd = p.vdeg[v];
      const uint32_t *meta = gt.gmeta + 3 * static_cast<int64_t>(gi);
      const double counter0 = u32_to_f64(__ldcg(&meta[2]));
      const double eix = counter0 - u32_to_f64(__ldcg(&meta[1]));
      const double ax = u32_to_f64(me.y) - vd;
      uint32_t *__restrict__ dense = gt.dense + static_cast<int64_t>(gi - gt.vert_lo) * gt.dense_stride;
      const uint32_t *__restrict__ keys = gt.klist + first * kGiantChunk + off;
      const uint32_t m = (cnt - off < kGiantChunk) ? cnt - off : kGiantChunk;
#pragma unroll
      for (int j0 = 0; j0 < kGiantChunk / 512; j0 += 4) {
        uint32_t rk[4], rw[4];
        double ayv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t i = threadIdx.x + (j0 + j) * 512;
          rk[j] = (i < m) ? __ldcs(&keys[i]) : kEmptyKeyU;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          rw[j] = 0u;
          if (rk[j] != kEmptyKeyU) rw[j] = atomicExch(&dense[rk[j] & kLabelMask], 0u);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) ayv[j] = rw[j] ? p.cdeg[rk[j] & kLabelMask] : 0.0;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (rw[j]) {
            const double g2 = modularity_gain(u32_to_f64(rw[j]), eix, vd, ayv[j], ax, p.constant);
            if (better_u(g2, rk[j], g, id)) {
              g = g2;
              id = rk[j];
            }
          }
      }
    }
    double wg;
    uint32_t wid = warp_argmax(g, id, &wg);
    if (lane == 0) {
      red_g[wib] = wg;
      red_id[wib] = wid;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int i = 1; i < 16; ++i)
        if (better_u(red_g[i], red_id[i], wg, wid)) {
          wg = red_g[i];
          wid = red_id[i];
        }
      chunk_gain[chunk] = wg;
      chunk_key[chunk] = wid;
    }
    __syncthreads();
  }
}

// one warp per big vertex: reduce its chunks' candidates, commit, reset its counters
__global__ void __launch_bounds__(128) k_sweep_u32_giant_final_dense(SweepU32Params p, const vid_t *__restrict__ list,
                                                                    int n, GiantTables gt,
                                                                    const double *__restrict__ chunk_gain,
                                                                    const uint32_t *__restrict__ chunk_key) {
  const int gi = static_cast<int>((static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5);
  const unsigned lane = threadIdx.x & 31u;
  if (gi >= n) return;
  const vid_t v = list[gi];
  const uint2 me = p.packed[self_of(p, v)];
  double g = -1.0;
  uint32_t id = kLabelMask;
  for (int64_t q = gt.chunk_start[gi] + lane; q < gt.chunk_start[gi + 1]; q += 32)
    if (better_u(chunk_gain[q], chunk_key[q], g, id)) {
      g = chunk_gain[q];
      id = chunk_key[q];
    }
  double wg;
  const uint32_t wid = warp_argmax(g, id, &wg);
  if (lane == 0) {
    uint32_t *meta = gt.gmeta + 3 * static_cast<int64_t>(gi);
    commit_vertex_u32(p, v, me.x, wg, wid, p.vdeg[v], u32_to_f64(meta[2]));
    meta[1] = 0u;
    meta[2] = 0u;
    gt.kcount[gi] = 0u;
  }
}

// Variant without the shared-memory stage: every entry goes straight to the
// vertex's global table, 8 independent L2 atomics in flight per thread, 4 blocks
// per SM.  (A/B against k_sweep_u32_giant_insert: CVB_GIANT_DIRECT=1.)
__global__ void __launch_bounds__(512, 4) k_sweep_u32_giant_insert_direct(SweepU32Params p,
                                                                         const vid_t *__restrict__ list, int n,
                                                                         GiantTables gt, int64_t chunk_lo,
                                                                         int64_t chunk_hi) {
  constexpr int K = kGiantChunk / 512;
  const unsigned lane = threadIdx.x & 31u;
  for (int64_t chunk = chunk_lo + blockIdx.x; chunk < chunk_hi; chunk += gridDim.x) {
    const int gi = gt.chunk_owner[chunk];
    const vid_t v = list[gi];
    const eid_t ve0 = p.offsets[v], ve1 = p.offsets[v + 1];
    const eid_t e0 = ve0 + (chunk - gt.chunk_start[gi]) * kGiantChunk;
    const int cnt = static_cast<int>(((e0 + kGiantChunk < ve1) ? e0 + kGiantChunk : ve1) - e0);
    const vid_t self = self_of(p, v);
    const uint32_t cc = p.packed[self].x & kLabelMask;
    const int64_t tbase = gt.table_start[gi];
    const int64_t gcap = gt.table_start[gi + 1] - tbase;
    const int glog2 = 63 - __clzll(gcap);
    const uint32_t gmask = static_cast<uint32_t>(gcap - 1);
    uint32_t *meta = gt.gmeta + 3 * static_cast<int64_t>(gi);
    const vid_t *__restrict__ row_tails = p.tails + e0;
    const uint32_t *__restrict__ row_weights = p.weights + e0;
    uint32_t key[K], ay[K], w[K];
    vid_t tail[K];
    uint32_t validm = 0u, active = 0u, ownsum = 0u, selfsum = 0u;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int i = threadIdx.x + k * 512;
      tail[k] = self;
      w[k] = 0u;
      if (i < cnt) {
        tail[k] = __ldcs(&row_tails[i]);
        w[k] = __ldcs(&row_weights[i]);
        validm |= 1u << k;
      }
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const uint2 pk = p.packed[tail[k]];
      const bool valid = (validm >> k) & 1u;
      key[k] = valid ? pk.x : kEmptyKeyU;
      ay[k] = pk.y;
      const bool own = (key[k] & kLabelMask) == cc;
      ownsum += own ? w[k] : 0u;
      selfsum += (valid && tail[k] == self) ? w[k] : 0u;
      active |= (valid && !own) ? (1u << k) : 0u;
    }
    gtable_insert_chunk<K>(gt.slots + (tbase - gt.slot_base), key, w, active, glog2, gmask);
    const uint32_t c0u = __reduce_add_sync(0xffffffffu, ownsum);
    const uint32_t slu = __reduce_add_sync(0xffffffffu, selfsum);
    if (lane == 0) {
      if (c0u) atomicAdd(&meta[2], c0u);
      if (slu) atomicAdd(&meta[1], slu);
    }
  }
}

__global__ void __launch_bounds__(512) k_sweep_u32_giant_eval(SweepU32Params p, const vid_t *__restrict__ list, int n,
                                                             GiantTables gt, int64_t piece_lo, int64_t piece_hi) {
  __shared__ double red_g[16];
  __shared__ uint32_t red_id[16];
  const unsigned lane = threadIdx.x & 31u;
  const int wib = threadIdx.x >> 5;
  for (int64_t piece = piece_lo + blockIdx.x; piece < piece_hi; piece += gridDim.x) {
    const int64_t s0 = piece * kGiantPiece;
    const int gi = giant_owner(gt.table_start, n, s0);
    uint2 *__restrict__ tslots = gt.slots + (s0 - gt.slot_base);
    const vid_t v = list[gi];
    const uint2 me = p.packed[self_of(p, v)];
    const double vd = p.vdeg[v];
    const uint32_t *meta = gt.gmeta + 3 * static_cast<int64_t>(gi);
    const double counter0 = u32_to_f64(__ldcg(&meta[2]));
    const double eix = counter0 - u32_to_f64(__ldcg(&meta[1]));
    const double ax = u32_to_f64(me.y) - vd;
    double g = -1.0;
    uint32_t id = kLabelMask;
#pragma unroll
    for (int j0 = 0; j0 < kGiantPiece / 512; j0 += 4) {  // four records (and their degree gathers) in flight
      uint2 r[4];
      double ayv[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) r[j] = __ldcg(&tslots[threadIdx.x + (j0 + j) * 512]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        ayv[j] = 0.0;
        if (r[j].x != kEmptyKeyU) {
          ayv[j] = p.cdeg[r[j].x & kLabelMask];
          tslots[threadIdx.x + (j0 + j) * 512] = make_uint2(kEmptyKeyU, 0u);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (r[j].x != kEmptyKeyU) {
          const double g2 = modularity_gain(u32_to_f64(r[j].y), eix, vd, ayv[j], ax, p.constant);
          if (better_u(g2, r[j].x, g, id)) {
            g = g2;
            id = r[j].x;
          }
        }
      }
    }
    double wg;
    uint32_t wid = warp_argmax(g, id, &wg);
    if (lane == 0) {
      red_g[wib] = wg;
      red_id[wib] = wid;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int i = 1; i < 16; ++i)
        if (better_u(red_g[i], red_id[i], wg, wid)) {
          wg = red_g[i];
          wid = red_id[i];
        }
      gt.piece_gain[piece] = wg;
      gt.piece_key[piece] = wid;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(128) k_sweep_u32_giant_final(SweepU32Params p, const vid_t *__restrict__ list, int n,
                                                              GiantTables gt) {
  const int gi = static_cast<int>((static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5);
  const unsigned lane = threadIdx.x & 31u;
  if (gi >= n) return;
  const vid_t v = list[gi];
  const uint2 me = p.packed[self_of(p, v)];
  const int64_t p0 = gt.table_start[gi] / kGiantPiece, p1 = gt.table_start[gi + 1] / kGiantPiece;
  double g = -1.0;
  uint32_t id = kLabelMask;
  for (int64_t q = p0 + lane; q < p1; q += 32) {
    const double g2 = gt.piece_gain[q];
    const uint32_t k2 = gt.piece_key[q];
    if (better_u(g2, k2, g, id)) {
      g = g2;
      id = k2;
    }
  }
  double wg;
  const uint32_t wid = warp_argmax(g, id, &wg);
  if (lane == 0) {
    uint32_t *meta = gt.gmeta + 3 * static_cast<int64_t>(gi);
    commit_vertex_u32(p, v, me.x, wg, wid, p.vdeg[v], u32_to_f64(meta[2]));
    meta[1] = 0u;
    meta[2] = 0u;
  }
}

// ---------------------------------------------------------------------------
// First sweep of a phase (every vertex is its own community) on graphs whose
// rows hold each tail at most once: all neighbour communities are distinct, so
// nothing needs merging — no table, no atomics: eiy is the entry's weight and
// the own-community weight is the self-loop.  TPV threads per vertex stride
// over the adjacency and keep the running (gain, id) best.
// ---------------------------------------------------------------------------
template <int TPV>
__global__ void __launch_bounds__(256) k_sweep_u32_first(SweepU32Params p, const vid_t *__restrict__ list, int n) {
  constexpr int VPB = 256 / TPV;
  constexpr int GL = TPV < 32 ? TPV : 32;
  __shared__ double red_g[8];
  __shared__ uint32_t red_id[8];
  __shared__ uint32_t red_self[8];
  const unsigned lane = threadIdx.x & 31u;
  const int tid = static_cast<int>(threadIdx.x % TPV);
  const int vslot = static_cast<int>(threadIdx.x / TPV);
  // every lane of the warp stays in the loop (idle vertices are masked) so that
  // the group reductions below run with the full-warp mask
  const int64_t stride = static_cast<int64_t>(gridDim.x) * VPB;
  const int64_t first_item = static_cast<int64_t>(blockIdx.x) * VPB;
  for (int64_t base_item = first_item; base_item < n; base_item += stride) {
    const int64_t item = base_item + vslot;
    const bool live = item < n;
    const vid_t v = live ? list[item] : 0;
    const eid_t e0 = live ? p.offsets[v] : 0, e1 = live ? p.offsets[v + 1] : 0;
    const vid_t self = self_of(p, v);
    const uint2 me = p.packed[self];
    const double vd = p.vdeg[v];
    // in the singleton state counter[0] == selfLoop, so eix = 0 for every gain;
    // the self-loop weight is only needed for clusterWeight and rides along
    const double eix = 0.0;
    const double ax = u32_to_f64(me.y) - vd;
    double g = -1.0;
    uint32_t id = kLabelMask;
    uint32_t selfw = 0u;
    for (eid_t eb = e0 + tid; eb < e1; eb += 4 * TPV) {  // 4 independent gathers in flight per thread
      vid_t t[4];
      uint32_t w[4];
      uint2 pk[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const eid_t e = eb + static_cast<eid_t>(k) * TPV;
        t[k] = self;
        w[k] = 0u;
        if (e < e1) {
          t[k] = __ldcs(&p.tails[e]);
          w[k] = __ldcs(&p.weights[e]);
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) pk[k] = p.packed[t[k]];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (t[k] != self) {
          const double g2 = modularity_gain(u32_to_f64(w[k]), eix, vd, u32_to_f64(pk[k].y), ax, p.constant);
          if (better_u(g2, pk[k].x, g, id)) {
            g = g2;
            id = pk[k].x;
          }
        } else {
          selfw += w[k];  // idle slots carry w = 0
        }
      }
    }
    uint32_t slu = selfw;
    if (TPV > 1) {
#pragma unroll
      for (int o = GL / 2; o > 0; o >>= 1) slu += __shfl_xor_sync(0xffffffffu, slu, o);
    }
    if (TPV > 32) {
      if (lane == 0) red_self[threadIdx.x >> 5] = slu;
      __syncthreads();
      slu = 0u;
      for (int i = 0; i < TPV / 32; ++i) slu += red_self[i];
    }
    const double counter0 = u32_to_f64(slu);
    if (TPV > 1) {
#pragma unroll
      for (int o = GL / 2; o > 0; o >>= 1) {
        const double g2 = __shfl_xor_sync(0xffffffffu, g, o);
        const uint32_t k2 = __shfl_xor_sync(0xffffffffu, id, o);
        if (better_u(g2, k2, g, id)) {
          g = g2;
          id = k2;
        }
      }
    }
    if (TPV > 32) {
      if (lane == 0) {
        red_g[threadIdx.x >> 5] = g;
        red_id[threadIdx.x >> 5] = id;
      }
      __syncthreads();
      if (threadIdx.x == 0)
        for (int i = 1; i < TPV / 32; ++i)
          if (better_u(red_g[i], red_id[i], g, id)) {
            g = red_g[i];
            id = red_id[i];
          }
    }
    if (live && tid == 0) commit_vertex_u32(p, v, me.x, g, id, vd, counter0);
    if (TPV > 32) __syncthreads();
  }
}

// Giant vertices in the first sweep: 4096-entry chunks over the whole grid, one
// (gain, key) candidate per chunk, then k_sweep_u32_first_giant_final.
__global__ void __launch_bounds__(512) k_sweep_u32_first_giant(SweepU32Params p, const vid_t *__restrict__ list, int n,
                                                              GiantTables gt, int64_t total_chunks,
                                                              double *__restrict__ chunk_gain,
                                                              uint32_t *__restrict__ chunk_key) {
  __shared__ double red_g[16];
  __shared__ uint32_t red_id[16];
  const unsigned lane = threadIdx.x & 31u;
  const int wib = threadIdx.x >> 5;
  for (int64_t chunk = blockIdx.x; chunk < total_chunks; chunk += gridDim.x) {
    const int gi = gt.chunk_owner[chunk];
    const vid_t v = list[gi];
    const eid_t ve0 = p.offsets[v], ve1 = p.offsets[v + 1];
    const eid_t e0 = ve0 + (chunk - gt.chunk_start[gi]) * kGiantChunk;
    const eid_t e1 = (e0 + kGiantChunk < ve1) ? e0 + kGiantChunk : ve1;
    const vid_t self = self_of(p, v);
    const uint2 me = p.packed[self];
    const double vd = p.vdeg[v];
    const double ax = u32_to_f64(me.y) - vd;
    double g = -1.0;
    uint32_t id = kLabelMask, selfw = 0u;
    vid_t t[8];
    uint32_t w[8];
    uint2 pk[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const eid_t e = e0 + threadIdx.x + static_cast<eid_t>(k) * 512;
      t[k] = self;
      w[k] = 0u;
      if (e < e1) {
        t[k] = __ldcs(&p.tails[e]);
        w[k] = __ldcs(&p.weights[e]);
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) pk[k] = p.packed[t[k]];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (t[k] != self) {
        const double g2 = modularity_gain(u32_to_f64(w[k]), 0.0, vd, u32_to_f64(pk[k].y), ax, p.constant);
        if (better_u(g2, pk[k].x, g, id)) {
          g = g2;
          id = pk[k].x;
        }
      } else {
        selfw += w[k];
      }
    }
    const uint32_t slu = __reduce_add_sync(0xffffffffu, selfw);
    if (lane == 0 && slu) atomicAdd(&gt.gmeta[3 * static_cast<int64_t>(gi) + 1], slu);
    double wg;
    uint32_t wid = warp_argmax(g, id, &wg);
    if (lane == 0) {
      red_g[wib] = wg;
      red_id[wib] = wid;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int i = 1; i < 16; ++i)
        if (better_u(red_g[i], red_id[i], wg, wid)) {
          wg = red_g[i];
          wid = red_id[i];
        }
      chunk_gain[chunk] = wg;
      chunk_key[chunk] = wid;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(128) k_sweep_u32_first_giant_final(SweepU32Params p, const vid_t *__restrict__ list,
                                                                    int n, GiantTables gt,
                                                                    const double *__restrict__ chunk_gain,
                                                                    const uint32_t *__restrict__ chunk_key) {
  const int gi = static_cast<int>((static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5);
  const unsigned lane = threadIdx.x & 31u;
  if (gi >= n) return;
  const vid_t v = list[gi];
  const uint2 me = p.packed[self_of(p, v)];
  double g = -1.0;
  uint32_t id = kLabelMask;
  for (int64_t q = gt.chunk_start[gi] + lane; q < gt.chunk_start[gi + 1]; q += 32)
    if (better_u(chunk_gain[q], chunk_key[q], g, id)) {
      g = chunk_gain[q];
      id = chunk_key[q];
    }
  double wg;
  const uint32_t wid = warp_argmax(g, id, &wg);
  if (lane == 0) {
    uint32_t *meta = gt.gmeta + 3 * static_cast<int64_t>(gi);
    commit_vertex_u32(p, v, me.x, wg, wid, p.vdeg[v], u32_to_f64(meta[1]));  // counter[0] = self-loop
    meta[1] = 0u;
  }
}

// rows_simple[0] = 0 when some row holds a tail twice or out of order
__global__ void k_check_simple_rows(const eid_t *__restrict__ offsets, const vid_t *__restrict__ tails, int64_t nv,
                                    int *__restrict__ not_simple) {
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const unsigned lane = threadIdx.x & 31u;
  if (row >= nv) return;
  int bad = 0;
  for (eid_t e = offsets[row] + 1 + lane; e < offsets[row + 1]; e += 32) bad |= tails[e] <= tails[e - 1];
  if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(not_simple, 1);
}

}  // namespace cvb
