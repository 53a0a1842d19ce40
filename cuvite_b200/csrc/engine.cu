// engine.cu — device-resident Louvain phase engine + the C ABI of
// include/cuvite_b200.h.  One engine = one GPU = one vertex-range partition.
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <thread>
#include <map>
#include <mutex>
#include <set>
#include <utility>
#include <vector>

#include "common.cuh"
#include "setup.cuh"
#include "sweep.cuh"
#include "sweep_u32.cuh"
#include "lm.cuh"
#include "coarsen.cuh"
#include "dist.cuh"

namespace cvb {

static thread_local std::string g_error;

void set_error(const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_error = buf;
}

template <typename T>
struct DevBuf {
  T *p = nullptr;
  size_t cap = 0;
  int ensure(size_t n) {
    if (n <= cap) return CVB_OK;
    if (p) CVB_CUDA(cudaFree(p));
    p = nullptr;
    cap = 0;
    const size_t want = n + n / 16 + 64;
    CVB_CUDA(cudaMalloc(reinterpret_cast<void **>(&p), want * sizeof(T)));
    cap = want;
    return CVB_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

struct HostBuf {  // pinned staging
  unsigned char *p = nullptr;
  size_t cap = 0;
  int ensure(size_t n) {
    if (n <= cap) return CVB_OK;
    if (p) CVB_CUDA(cudaFreeHost(p));
    p = nullptr;
    cap = 0;
    CVB_CUDA(cudaMallocHost(reinterpret_cast<void **>(&p), n + n / 16 + 64));
    cap = n + n / 16 + 64;
    return CVB_OK;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
  }
};

// Small host-side parallel loop for the "+=" contract of the plugin call
// (localCupdate / remoteCupdate are ADDED to, louvain.cpp:2307-2376).
template <typename F>
static void host_parallel_for(int64_t n, F body) {
  static int nt = [] {
    if (const char *s = getenv("CVB_HOST_THREADS")) return atoi(s) > 0 ? atoi(s) : 1;
    return 1;  // measured on the B200 host: spawning helpers per call costs more than the 2M adds
  }();
  if (n < (1 << 16) || nt <= 1) {
    body(0, n);
    return;
  }
  std::vector<std::thread> th;
  const int64_t chunk = (n + nt - 1) / nt;
  for (int t = 0; t < nt; ++t) {
    const int64_t lo = t * chunk, hi = lo + chunk < n ? lo + chunk : n;
    if (lo >= hi) break;
    th.emplace_back([=]() { body(lo, hi); });
  }
  for (auto &x : th) x.join();
}

static inline int grid_for(int64_t n, int block, int max_blocks = 148 * 16) {
  int64_t g = (n + block - 1) / block;
  if (g < 1) g = 1;
  if (g > max_blocks) g = max_blocks;
  return static_cast<int>(g);
}

}  // namespace cvb

using namespace cvb;

struct cvb_engine {
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int device = 0;
  int num_sms = 148;

  // graph (one phase)
  int64_t nvl = 0, base = 0, nvg = 0, ne = 0;
  int mode = 0;
  bool graph_ready = false, phase_ready = false;
  DevBuf<eid_t> offsets;
  DevBuf<vid_t> tails;
  DevBuf<unsigned char> weights;
  DevBuf<double> vdeg, cdeg, ddeg, cw;
  DevBuf<int32_t> csize, dsize;
  DevBuf<vid_t> lab[3], labels_global, labels_global2;  // labels_global2: gather target of the next labels
  DevBuf<double> vdeg_global;                           // replicated vertex degrees (N > 1)
  bool glabels_swapped = false;
  bool labels_only = false;  // N > 1: community info rebuilt from gathered labels (cvb_rebuild_and_reduce)
  int past = 0, curr = 1, targ = 2;
  DevBuf<vid_t> list;
  DevBuf<int32_t> block_counts;
  DevBuf<int64_t> d_class_start;
  DevBuf<unsigned long long> d_class_entries;  // [kNumClasses] + max degree at [kNumClasses]
  int64_t class_start[kNumClasses + 1] = {0};
  int64_t class_entries[kNumClasses] = {0};
  int64_t max_degree = 0;
  double local_weight = 0.0;
  double constant = 0.0;

  // reductions
  DevBuf<double> partials, d_sums;
  DevBuf<unsigned int> ticket;
  DevBuf<int> d_flags;
  HostBuf h_sums;

  // scratch
  DevBuf<unsigned char> staging;   // upload / boundary staging (device)
  DevBuf<unsigned char> staging2;  // boundary outputs (device)
  DevBuf<unsigned char> giant;     // global-memory tables
  HostBuf h_stage;                 // pinned host staging
  int giant_log2cap = 0;
  int giant_blocks = 0;
  size_t giant_stride = 0;

  // CVB_WEIGHTS_U32 fast path
  DevBuf<uint2> packed;                 // [nv_global] {label|singleton bit, community degree}
  DevBuf<int64_t> giant_layout;         // chunk_start[ng+1], table_start[ng+1]
  DevBuf<uint32_t> gmeta, gpiece_key;
  DevBuf<uint2> gslots;
  DevBuf<double> gpiece_gain, gchunk_gain;
  DevBuf<uint32_t> gchunk_key;
  DevBuf<int32_t> gchunk_owner;
  // "big" = every vertex with deg > 8192 (cluster classes + giants): chunk layout of the first-sweep kernels
  DevBuf<int64_t> big_layout;
  DevBuf<int32_t> big_owner;
  DevBuf<uint32_t> big_meta;
  int64_t big_chunks = 0;
  int64_t giant_chunks = 0, giant_slots = 0;
  // coarsening stage 2 output: distinct (src << 32 | dst) pairs of the local rows, ascending
  DevBuf<uint64_t> pair_keys;
  DevBuf<unsigned char> pair_w;
  int64_t pair_count = 0;
  // The global tables of the big vertices are processed in batches that share ONE arena about
  // the size of L2 (insert -> eval -> the next batch reuses the same lines), instead of one
  // table per vertex spread over gigabytes of HBM (3 GB at R-MAT scale 24).  Measured: the
  // arena bounds memory but buys little time (scale 24: 7.6 ms/sweep at 64 MB and unbatched,
  // 7.85 at 32 MB where batches stop filling the GPU) — the insert is bound by L2 atomics.
  struct GiantBatch { int64_t chunk_lo, chunk_hi, piece_lo, piece_hi, slot_base; int vert_lo; };
  // CVB_BIG_DENSE=1 (A/B variant, off by default): every vertex of a batch owns a DENSE uint32
  // accumulator indexed by community id (no hashing in global memory: chunk totals are added
  // with fire-and-forget L2 reductions) plus a list of the communities it met.  Measured on
  // B200: same speed at scale 20 (0.382 vs 0.388 ms/sweep), SLOWER at scale 24 (9.6 vs 7.4 ms:
  // the touched sectors of 128 concurrent accumulators no longer fit L2), so the open-addressing
  // tables stay the default.
  bool big_dense = false;
  int64_t dense_budget_bytes = 8LL << 30;  // CVB_BIG_DENSE_MB overrides
  int64_t dense_stride = 0;
  DevBuf<uint32_t> dense, kcount, klist;
  DevBuf<double> gchunk2_gain;
  DevBuf<uint32_t> gchunk2_key;
  std::vector<GiantBatch> giant_batches;
  int64_t giant_arena_slots = 16 << 20;  // 128 MB of 8-byte slots (about one L2); CVB_GIANT_ARENA_MB overrides
  bool use_tma = false;        // CVB_TMA=1: cp.async.bulk + mbarrier staging of tails/weights tiles
  bool use_cluster = false;    // CVB_CLUSTER=1: DSMEM cluster kernels for 8193..65536
  bool use_twopass = true;     // 8193..16384: two-pass shared-memory kernel (CVB_NO_TWOPASS=1: global tables)
  bool use_split = false;      // CVB_SPLIT=1: hash classes gather 4-byte labels, degrees per distinct community (unmeasured A/B)
  DevBuf<uint32_t> cdeg32;
  // second-generation kernels (lm.cuh): per-vertex records in class-list order, self-loop weights
  DevBuf<VertexRec> vrec;
  DevBuf<uint32_t> selfw;
  // "big" path (lm.cuh, partition then hash) for every vertex above 8192 entries
  DevBuf<int32_t> bg_log2p, bg_item_start, bg_item_owner;
  DevBuf<uint4> bg_records;
  DevBuf<uint32_t> bg_cursor, bg_own, bg_cand_key;
  DevBuf<double> bg_cand_gain;
  int bg_items = 0;
  bool use_big = true;         // CVB_BIG=0: first-generation two-pass / global-table kernels
  bool big_ready = false;
  DevBuf<uint2> mp_stage;      // multi-pass kernel: staged packed records of the 8193..65536 rows
  DevBuf<uint32_t> mp_off, mp_deg;
  DevBuf<int> mp_redo;
  bool use_mp = false;         // CVB_MP=1: serial multi-pass kernel for 8193..65536 (superseded by the big path)
  bool mp_ready = false;
  DevBuf<vid_t> rank, tails_r;  // degree-ordered gather layout (one GPU): rank[v], tails mapped to ranks
  bool remap = false;          // the sweep reads tails_r / rank-ordered packed[]
  bool use_remap = true;       // CVB_REMAP=0 disables it
  bool use_hash2 = true;       // CVB_HASH2=0: the round-1 hash kernels (separate key / weight arrays, no unit / singleton paths)
  bool use_lm = true;          // CVB_LM=0: first-generation hash/group kernels only
  uint32_t lm_mask = (1u << kClsG8) | (1u << kClsG16) | (1u << kClsG32) | (1u << kClsH16K);  // classes on lm.cuh
  bool lm_ready = false;       // vrec/selfw built for the current graph (U32 mode, < 2^32 entries)
  bool unit_weights = false;   // every weight == 1: the sweep never reads the weights array
  bool has_selfloops = false;
  bool l2_persist = false;     // CVB_L2_PERSIST=1: access-policy window that keeps `packed` persisting in L2 (unmeasured A/B)
  // first degree class that goes through the global tables
  int global_table_class() const {
    if (use_lm && lm_ready && use_big && big_ready) return kNumClasses;  // nothing goes through the global tables
    if (use_lm && lm_ready && use_mp) return kClsGiant;
    return use_cluster ? kClsGiant : (use_twopass ? kClsC4 : kClsC2);
  }
  bool giant_direct = false;   // CVB_GIANT_DIRECT=1: giant entries go straight to the global tables
  bool use_first_sweep = true; // cvb env CVB_NO_FIRST_SWEEP=1 disables the singleton fast path
  bool rows_simple = false;    // every row holds each tail at most once (ascending)
  int sweeps_in_phase = 0;     // 0 right after cvb_phase_begin: all vertices are singletons
  static constexpr int kAux = 12;
  cudaStream_t aux[kAux] = {};
  cudaEvent_t ev_fork = nullptr, ev_join[kAux] = {};
  int nstreams = 12;  // streams the class kernels of a sweep are spread over (CVB_STREAMS; r01 used 4)
  cudaStream_t S(int k) const { return aux[k % nstreams]; }

  DistState dist;  // N > 1: NCCL communicator and vertex ranges (dist.cuh)
  // coarsening scratch, kept between levels and runs (cudaMalloc / cudaFree of multi-GB buffers per level cost more
  // than the sorts on a box where 8 processes share the driver; CVB_KEEP_SCRATCH=0 frees it after every call)
  struct CoarsenScratch {
    DevBuf<int32_t> first, flag, rank, head, seg, vnew;
    DevBuf<unsigned char> tmp, w_out, uw;
    DevBuf<uint64_t> k_in, k_out, ukeys;
    DevBuf<int64_t> seg_start, noff;
    DevBuf<unsigned long long> row_count;
    void release() {
      first.release(); flag.release(); rank.release(); head.release(); seg.release(); vnew.release(); tmp.release();
      w_out.release(); uw.release(); k_in.release(); k_out.release(); ukeys.release(); seg_start.release();
      noff.release(); row_count.release();
    }
  } cs;
  bool keep_scratch = true;
  // colouring (cvb_phase_run_color): per-colour class lists; ovr_* redirect the sweep's class ranges
  DevBuf<int32_t> col_color, col_keys, col_keys2;
  DevBuf<vid_t> col_list, col_ids;
  DevBuf<VertexRec> col_vrec;
  DevBuf<vid_t> col_bak_t;
  DevBuf<double> col_bak_cw;
  int col_num = 0;                       // numColor passed to the Louvain loop (= colours + 1)
  std::vector<int64_t> col_start;        // [col_num * kNumClasses + 1] start of every (colour, class) range in col_list
  bool ovr = false;                      // a colour step: classes below kClsC2 come from col_list / col_vrec
  bool ovr_big = false;                  // ... and this colour has vertices above 8192 entries
  bool ovr_et = false;                   // ... with early termination (frozen vertices keep target / clusterWeight)
  int ovr_color = 0;
  // early termination (cvb_phase_run_et)
  bool et_no_deltas = false;
  DevBuf<unsigned char> et_active;
  DevBuf<vid_t> et_frozen_target;
  DevBuf<double> et_frozen_cw, et_p[2];
  DevBuf<unsigned int> et_count;

  int64_t launches = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t up_ev[8] = {};  // narrowing upload: one per staging slot
  bool sweep_timed = false;

  vid_t *glabels_curr() { return glabels_swapped ? labels_global2.p : labels_global.p; }
  vid_t *glabels_next() { return glabels_swapped ? labels_global.p : labels_global2.p; }
  const vid_t *sweep_labels() const {
    return nvl == nvg ? lab[curr].p : (glabels_swapped ? labels_global2.p : labels_global.p);
  }
};

namespace {

// cudaFuncSetAttribute is per device: remember (device, kernel) pairs, not one flag per kernel,
// so that engines on several devices of one process all get their dynamic shared memory opt-in
bool first_use_on_device(int device, const void *kernel) {
  static std::mutex m;
  static std::set<std::pair<int, const void *>> seen;
  std::lock_guard<std::mutex> lock(m);
  return seen.insert({device, kernel}).second;
}

template <typename W, int CAP>
int launch_warp(cvb_engine *e, const SweepParams &p, const vid_t *list, int64_t n) {
  const size_t smem = (sizeof(W) + sizeof(vid_t)) * CAP * 4;
  if (first_use_on_device(e->device, reinterpret_cast<const void *>(&k_sweep_warp<W, CAP>))) {
    CVB_CUDA(cudaFuncSetAttribute(k_sweep_warp<W, CAP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(smem)));
  }
  const int grid = static_cast<int>((n + 3) / 4);
  k_sweep_warp<W, CAP><<<grid, 128, smem, e->stream>>>(p, list, static_cast<int>(n));
  e->launches++;
  return CVB_OK;
}

template <typename W, int CAP, int THREADS>
int launch_block(cvb_engine *e, const SweepParams &p, const vid_t *list, int64_t n) {
  const size_t smem = (sizeof(W) + sizeof(vid_t)) * CAP;
  if (first_use_on_device(e->device, reinterpret_cast<const void *>(&k_sweep_block<W, CAP, THREADS, false>))) {
    CVB_CUDA(cudaFuncSetAttribute(k_sweep_block<W, CAP, THREADS, false>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  }
  k_sweep_block<W, CAP, THREADS, false>
      <<<static_cast<int>(n), THREADS, smem, e->stream>>>(p, list, static_cast<int>(n), nullptr, 0, 0);
  e->launches++;
  return CVB_OK;
}

template <typename W>
int launch_sweep(cvb_engine *e) {
  SweepParams p;
  p.offsets = e->offsets.p;
  p.tails = e->tails.p;
  p.weights = e->weights.p;
  p.labels = e->sweep_labels();
  p.vdeg = e->vdeg.p;
  p.cdeg = e->cdeg.p;
  p.csize = e->csize.p;
  p.constant = e->constant;
  p.base = static_cast<vid_t>(e->base);
  p.target = e->lab[e->targ].p;
  p.cw = e->cw.p;
  p.ddeg = (e->labels_only || e->et_no_deltas) ? nullptr : e->ddeg.p;
  p.dsize = (e->labels_only || e->et_no_deltas) ? nullptr : e->dsize.p;
  // (colour steps: the (colour, class) ranges of the per-colour lists — every kernel here takes a list)
  const int64_t *cs = e->ovr ? e->col_start.data() + static_cast<size_t>(e->ovr_color) * kNumClasses : e->class_start;
  const vid_t *cl = e->ovr ? e->col_list.p : e->list.p;
  auto cnt = [&](int c) { return cs[c + 1] - cs[c]; };
  auto lst = [&](int c) { return cl + cs[c]; };
  int64_t n;
  if ((n = cnt(kClsG4)) > 0) {
    k_sweep_group<4, W><<<static_cast<int>((n * 4 + 255) / 256), 256, 0, e->stream>>>(p, lst(kClsG4), (int)n);
    e->launches++;
  }
  if ((n = cnt(kClsG8)) > 0) {
    k_sweep_group<8, W><<<static_cast<int>((n * 8 + 255) / 256), 256, 0, e->stream>>>(p, lst(kClsG8), (int)n);
    e->launches++;
  }
  if ((n = cnt(kClsG16)) > 0) {
    k_sweep_group<16, W><<<static_cast<int>((n * 16 + 255) / 256), 256, 0, e->stream>>>(p, lst(kClsG16), (int)n);
    e->launches++;
  }
  if ((n = cnt(kClsG32)) > 0) {
    k_sweep_group<32, W><<<static_cast<int>((n * 32 + 255) / 256), 256, 0, e->stream>>>(p, lst(kClsG32), (int)n);
    e->launches++;
  }
  // adjacent classes are contiguous in the list, so two classes go out in one launch
  auto cnt2 = [&](int c) { return cs[c + 2] - cs[c]; };
  if ((n = cnt2(kClsH128)) > 0) CVB_TRY((launch_warp<W, 256>(e, p, lst(kClsH128), n)));
  if ((n = cnt2(kClsH512)) > 0) CVB_TRY((launch_warp<W, 1024>(e, p, lst(kClsH512), n)));
  if ((n = cnt2(kClsH2K)) > 0) CVB_TRY((launch_block<W, 4096, 256>(e, p, lst(kClsH2K), n)));
  if ((n = cnt2(kClsH8K)) > 0) CVB_TRY((launch_block<W, 16384, 1024>(e, p, lst(kClsH8K), n)));
  if ((n = cs[kClsGiant + 1] - cs[kClsC2]) > 0) {  // cluster classes + giants
    const int grid = static_cast<int>(n < e->giant_blocks ? n : e->giant_blocks);
    k_sweep_block<W, 1024, 1024, true><<<grid, 1024, 0, e->stream>>>(p, lst(kClsC2), (int)n, e->giant.p,
                                                                    e->giant_stride, e->giant_log2cap);
    e->launches++;
  }
  CVB_CUDA(cudaGetLastError());
  return CVB_OK;
}

template <int TPV, int KC, int NCH, int MINB, bool TMA, bool SPLIT = false>
int launch_hash_impl(cvb_engine *e, cudaStream_t st, const SweepU32Params &p, const vid_t *list, int64_t n) {
  constexpr int VPB = TPV <= 32 ? 256 / TPV : 1;
  constexpr int THREADS = TPV <= 32 ? 256 : TPV;
  const size_t smem = hash_block_smem<TPV, KC, NCH, TMA>();
  if (first_use_on_device(e->device, reinterpret_cast<const void *>(&k_sweep_u32_hash<TPV, KC, NCH, MINB, TMA, SPLIT>))) {
    CVB_CUDA(cudaFuncSetAttribute(k_sweep_u32_hash<TPV, KC, NCH, MINB, TMA, SPLIT>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  }
  int64_t grid = (n + VPB - 1) / VPB;
  const int64_t cap = static_cast<int64_t>(e->num_sms) * MINB;
  if (grid > cap) grid = cap;
  k_sweep_u32_hash<TPV, KC, NCH, MINB, TMA, SPLIT>
      <<<static_cast<int>(grid), THREADS, smem, st>>>(p, list, static_cast<int>(n));
  e->launches++;
  return CVB_OK;
}

template <int TPV, int KC, int NCH, int MINB, bool UNIT>
int launch_hash2_impl(cvb_engine *e, cudaStream_t st, const SweepU32Params &p, const vid_t *list, int64_t n) {
  constexpr int VPB = TPV <= 32 ? 256 / TPV : 1;
  constexpr int THREADS = TPV <= 32 ? 256 : TPV;
  const size_t smem = hash2_block_smem<TPV, KC, NCH>();
  if (first_use_on_device(e->device, reinterpret_cast<const void *>(&k_sweep_u32_hash2<TPV, KC, NCH, MINB, UNIT>))) {
    CVB_CUDA(cudaFuncSetAttribute(k_sweep_u32_hash2<TPV, KC, NCH, MINB, UNIT>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  }
  int64_t grid = (n + VPB - 1) / VPB;
  const int64_t cap = static_cast<int64_t>(e->num_sms) * MINB;
  if (grid > cap) grid = cap;
  k_sweep_u32_hash2<TPV, KC, NCH, MINB, UNIT><<<static_cast<int>(grid), THREADS, smem, st>>>(p, list, static_cast<int>(n));
  e->launches++;
  return CVB_OK;
}

template <int TPV, int KC, int NCH, int MINB>
int launch_hash(cvb_engine *e, cudaStream_t st, const SweepU32Params &p, const vid_t *list, int64_t n) {
  if (e->use_hash2 && !e->use_split && !(TPV >= 128 && e->use_tma))
    return e->unit_weights ? launch_hash2_impl<TPV, KC, NCH, MINB, true>(e, st, p, list, n)
                           : launch_hash2_impl<TPV, KC, NCH, MINB, false>(e, st, p, list, n);
  if (e->use_split) return launch_hash_impl<TPV, KC, NCH, MINB, false, true>(e, st, p, list, n);
  if (TPV >= 128 && e->use_tma) return launch_hash_impl<TPV, KC, NCH, MINB, (TPV >= 128)>(e, st, p, list, n);
  return launch_hash_impl<TPV, KC, NCH, MINB, false>(e, st, p, list, n);
}

template <int CS>
int launch_cluster(cvb_engine *e, cudaStream_t st, const SweepU32Params &p, const vid_t *list, int64_t n) {
  const size_t smem = static_cast<size_t>(kClusterSlice) * 8 + 16;
  if (first_use_on_device(e->device, reinterpret_cast<const void *>(&k_sweep_u32_cluster<CS>))) {
    CVB_CUDA(cudaFuncSetAttribute(k_sweep_u32_cluster<CS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(smem)));
  }
  int64_t nclusters = n;
  const int64_t cap = e->num_sms / CS;
  if (nclusters > cap) nclusters = cap;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(nclusters * CS));
  cfg.blockDim = dim3(1024);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CVB_CUDA(cudaLaunchKernelEx(&cfg, k_sweep_u32_cluster<CS>, p, list, static_cast<int>(n)));
  e->launches++;
  return CVB_OK;
}

// ---- second-generation kernels (lm.cuh) ----------------------------------------
// dynamic shared memory opt-in (once per device) and resident blocks per SM of a kernel
static int lm_setup(cvb_engine *e, const void *kernel, int threads, size_t smem) {
  static std::mutex m;
  static std::map<std::pair<int, const void *>, int> occ;
  std::lock_guard<std::mutex> lock(m);
  auto it = occ.find({e->device, kernel});
  if (it != occ.end()) return it->second;
  if (smem > 40 * 1024 &&  // static shared memory counts against the 48 KB default too
      cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess)
    return -1;
  int blocks = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, kernel, threads, smem) != cudaSuccess || blocks < 1)
    return -1;
  occ[{e->device, kernel}] = blocks;
  return blocks;
}

template <int G, int U, int CAPLOG2>
int launch_lm_group(cvb_engine *e, cudaStream_t st, const LmParams &p, int64_t first, int64_t n) {
  if (n <= 0) return CVB_OK;
  constexpr int NG = 32 / G;
  constexpr size_t smem = lm_group_smem<G, U, CAPLOG2>();
  auto go = [&](auto kern) -> int {
    const int occ = lm_setup(e, reinterpret_cast<const void *>(kern), 256, smem);
    CVB_CHECK(occ > 0, CVB_ERR_CUDA, "lm group kernel <%d,%d,%d>: cannot configure (%s)", G, U, CAPLOG2,
              cudaGetErrorString(cudaGetLastError()));
    int64_t grid = ((n + NG - 1) / NG + 7) / 8;
    const int64_t cap = static_cast<int64_t>(e->num_sms) * occ;
    if (grid > cap) grid = cap;
    kern<<<static_cast<int>(grid), 256, smem, st>>>(p, static_cast<int>(first), static_cast<int>(n));
    e->launches++;
    return CVB_OK;
  };
  static const bool v4 = [] { const char *s = getenv("CVB_LM4G"); return !(s && s[0] == '0'); }();
  if (v4) return e->unit_weights ? go(k_lm_group4<G, U, CAPLOG2, true>) : go(k_lm_group4<G, U, CAPLOG2, false>);
  return e->unit_weights ? go(k_lm_group<G, U, CAPLOG2, true>) : go(k_lm_group<G, U, CAPLOG2, false>);
}

template <int TPV, int CAPLOG2>
int launch_lm_block(cvb_engine *e, cudaStream_t st, const LmParams &p, int64_t first, int64_t n) {
  if (n <= 0) return CVB_OK;
  constexpr size_t smem = lm_block_smem<TPV, CAPLOG2>();
  static const bool v4 = [] { const char *s = getenv("CVB_LM4"); return !(s && s[0] == '0'); }();
  auto go = [&](auto kern) -> int {
    const int occ = lm_setup(e, reinterpret_cast<const void *>(kern), TPV, smem);
    CVB_CHECK(occ > 0, CVB_ERR_CUDA, "lm block kernel <%d,%d>: cannot configure (%s)", TPV, CAPLOG2,
              cudaGetErrorString(cudaGetLastError()));
    int64_t grid = n;
    const int64_t cap = static_cast<int64_t>(e->num_sms) * occ;
    if (grid > cap) grid = cap;
    kern<<<static_cast<int>(grid), TPV, smem, st>>>(p, static_cast<int>(first), static_cast<int>(n));
    e->launches++;
    return CVB_OK;
  };
  if (v4) return e->unit_weights ? go(k_lm_block4<TPV, CAPLOG2, true>) : go(k_lm_block4<TPV, CAPLOG2, false>);
  return e->unit_weights ? go(k_lm_block<TPV, CAPLOG2, true>) : go(k_lm_block<TPV, CAPLOG2, false>);
}

// The CVB_WEIGHTS_U32 sweep: pack labels, then the degree-class kernels forked
// over the auxiliary streams and joined back on the engine stream.
int launch_sweep_u32(cvb_engine *e) {
  cudaStream_t st = e->stream;
  const bool lm = e->use_lm && e->lm_ready;
  const bool first_sweep = e->sweeps_in_phase == 0 && e->rows_simple && e->use_first_sweep && !e->ovr;
  const bool remap = lm && e->remap && !e->use_split;  // the split-gather A/B variant indexes labels[] by vertex id
  k_pack_labels<<<grid_for(e->nvg, 256), 256, 0, st>>>(e->nvg, e->sweep_labels(), e->cdeg.p, e->csize.p,
                                                       e->packed.p, remap ? e->rank.p : nullptr);
  e->launches++;
  SweepU32Params p;
  p.offsets = e->offsets.p;
  p.tails = remap ? e->tails_r.p : e->tails.p;
  p.rank = remap ? e->rank.p : nullptr;
  p.weights = reinterpret_cast<const uint32_t *>(e->weights.p);
  p.packed = e->packed.p;
  p.cdeg = e->cdeg.p;
  p.vdeg = e->vdeg.p;
  p.constant = e->constant;
  p.base = static_cast<vid_t>(e->base);
  p.target = e->lab[e->targ].p;
  p.cw = e->cw.p;
  p.ddeg = (e->labels_only || e->et_no_deltas) ? nullptr : e->ddeg.p;  // labels-only exchange derives the deltas from labels
  p.dsize = (e->labels_only || e->et_no_deltas) ? nullptr : e->dsize.p;
  p.labels = e->sweep_labels();
  p.csize = e->csize.p;
  p.cdeg32 = nullptr;
  // singleton bypass: needs community sizes that agree with the labels; early termination rotates the labels of
  // frozen vertices WITHOUT updating the community info (louvain.cpp:174-186, mirrored), so it is off there
  // (and in colour steps: the community info is updated colour by colour while the labels stay those of the iteration)
  p.simple_rows = (e->rows_simple && !e->et_no_deltas && !e->ovr && !getenv("CVB_NO_SINGLETON_SKIP")) ? 1 : 0;
  if (e->use_split) {
    CVB_TRY(e->cdeg32.ensure(e->nvg + 1));
    k_cdeg32<<<grid_for(e->nvg, 256), 256, 0, st>>>(e->nvg, e->cdeg.p, e->cdeg32.p);
    e->launches++;
    p.cdeg32 = e->cdeg32.p;
  }
  // class ranges: the degree-class list of the graph, or — in a colour step — the range of (colour, class)
  // in the per-colour lists for the classes below kClsC2 (the big classes always run over all their vertices)
  auto cpos = [&](int c) -> int64_t {
    return (e->ovr && c < kClsC2) ? e->col_start[e->ovr_color * kNumClasses + c] : e->class_start[c];
  };
  auto cnt = [&](int c) -> int64_t {
    return (e->ovr && c < kClsC2) ? e->col_start[e->ovr_color * kNumClasses + c + 1] - e->col_start[e->ovr_color * kNumClasses + c]
                                  : e->class_start[c + 1] - e->class_start[c];
  };
  auto lst = [&](int c) { return ((e->ovr && c < kClsC2) ? e->col_list.p : e->list.p) + cpos(c); };
  int64_t n;
  LmParams q{};
  if (lm) {
    q.vrec = e->ovr ? e->col_vrec.p : e->vrec.p;
    q.tails = remap ? e->tails_r.p : e->tails.p;
    q.remap = remap ? ((e->nvl == e->nvg && !e->ovr) ? 1 : 2) : 0;  // colour lists: rank[] lookup
    q.rank = e->rank.p;
    q.rank_top = static_cast<int>(e->nvl - 1);
    q.simple_rows = p.simple_rows;
    q.weights = reinterpret_cast<const uint32_t *>(e->weights.p);
    q.packed = e->packed.p;
    q.selfw = e->has_selfloops ? e->selfw.p : nullptr;
    q.constant = e->constant;
    q.base = static_cast<vid_t>(e->base);
    q.target = p.target;
    q.cw = p.cw;
    q.ddeg = p.ddeg;
    q.dsize = p.dsize;
  }
  auto pos = [&](int c) { return cpos(c); };
  if (first_sweep) {
    // singleton start: nothing to merge, table-free kernels (one launch per width)
    auto span = [&](int c0, int c1) { return e->class_start[c1 + 1] - e->class_start[c0]; };
    CVB_CUDA(cudaEventRecord(e->ev_fork, st));
    for (int i = 0; i < e->nstreams; ++i) CVB_CUDA(cudaStreamWaitEvent(e->aux[i], e->ev_fork, 0));
    auto go = [&](auto kern, int tpv, int c0, int c1, cudaStream_t s2) {
      const int64_t m = span(c0, c1);
      if (m <= 0) return;
      const int vpb = 256 / tpv;
      int64_t grid = (m + vpb - 1) / vpb;
      const int64_t cap = static_cast<int64_t>(e->num_sms) * 32;
      if (grid > cap) grid = cap;
      kern<<<static_cast<int>(grid), 256, 0, s2>>>(p, lst(c0), static_cast<int>(m));
      e->launches++;
    };
    if (lm) {
      auto gof = [&](auto kern_unit, auto kern_w, int threads, int vpb, int c0, int c1, cudaStream_t s2) {
        const int64_t m = span(c0, c1);
        if (m <= 0) return;
        int64_t grid = (m + vpb - 1) / vpb;
        const int64_t cap = static_cast<int64_t>(e->num_sms) * (2048 / threads);
        if (grid > cap) grid = cap;
        if (e->unit_weights)
          kern_unit<<<static_cast<int>(grid), threads, 0, s2>>>(q, static_cast<int>(pos(c0)), static_cast<int>(m));
        else
          kern_w<<<static_cast<int>(grid), threads, 0, s2>>>(q, static_cast<int>(pos(c0)), static_cast<int>(m));
        e->launches++;
      };
      gof(k_lm_first_block<512, true>, k_lm_first_block<512, false>, 512, 1, kClsH8K, kClsH16K, e->S(0));
      gof(k_lm_first_block<128, true>, k_lm_first_block<128, false>, 128, 1, kClsH2K, kClsH4K, e->S(4));
      gof(k_lm_first_group<32, true>, k_lm_first_group<32, false>, 256, 8, kClsH512, kClsH1K, e->S(2));
      gof(k_lm_first_group<16, true>, k_lm_first_group<16, false>, 256, 16, kClsH256, kClsH256, e->S(5));
      gof(k_lm_first_group<8, true>, k_lm_first_group<8, false>, 256, 32, kClsG32, kClsH128, e->S(3));
      gof(k_lm_first_group<4, true>, k_lm_first_group<4, false>, 256, 64, kClsG8, kClsG16, e->S(6));
      go(k_sweep_u32_first<1>, 1, kClsG4, kClsG4, e->S(7));
    } else {
    go(k_sweep_u32_first<256>, 256, kClsH2K, kClsH16K, e->S(0));
    go(k_sweep_u32_first<32>, 32, kClsH256, kClsH1K, e->S(2));
    go(k_sweep_u32_first<8>, 8, kClsG16, kClsH128, e->S(3));
    go(k_sweep_u32_first<1>, 1, kClsG4, kClsG8, e->S(3));
    }
    if ((n = span(kClsC2, kClsGiant)) > 0) {  // every vertex above 8192 entries: chunked over the grid
      GiantTables gt{};
      gt.chunk_owner = e->big_owner.p;
      gt.chunk_start = e->big_layout.p;
      gt.table_start = e->big_layout.p + (n + 1);
      gt.gmeta = e->big_meta.p;
      int64_t g1 = e->big_chunks < 8LL * e->num_sms ? e->big_chunks : 8LL * e->num_sms;
      k_sweep_u32_first_giant<<<static_cast<int>(g1), 512, 0, e->S(1)>>>(p, lst(kClsC2), (int)n, gt, e->big_chunks,
                                                                          e->gchunk_gain.p, e->gchunk_key.p);
      k_sweep_u32_first_giant_final<<<static_cast<int>((n * 32 + 127) / 128), 128, 0, e->S(1)>>>(
          p, lst(kClsC2), (int)n, gt, e->gchunk_gain.p, e->gchunk_key.p);
      e->launches += 2;
    }
    for (int i = 0; i < e->nstreams; ++i) {
      CVB_CUDA(cudaEventRecord(e->ev_join[i], e->aux[i]));
      CVB_CUDA(cudaStreamWaitEvent(st, e->ev_join[i], 0));
    }
    CVB_CUDA(cudaGetLastError());
    return CVB_OK;
  }
  CVB_CUDA(cudaEventRecord(e->ev_fork, st));
  for (int i = 0; i < e->nstreams; ++i) CVB_CUDA(cudaStreamWaitEvent(e->aux[i], e->ev_fork, 0));
  // Which generation serves a degree class is decided per class from the measurements of round 2
  // (profiles/r02_*: R-MAT scale 20 and 24): the second-generation kernels (lm.cuh) win on the small
  // classes (5..32 entries: per-vertex records, sub-warp groups without match.any), on 4097..8192 and,
  // with the partition-then-hash path, on everything above 8192; the first-generation hash kernels
  // (eight register-resident claims per thread, two-deep header pipeline) stay faster from 33 to 4096.
  // CVB_LM_MASK overrides (bit c = class c on lm.cuh).
  const uint32_t lm_mask = lm ? e->lm_mask : 0u;
  auto on_lm = [&](int c) { return ((lm_mask >> c) & 1u) != 0u; };
  // hash classes: 8 CSR entries per thread, TPV = upper degree / 8 threads per vertex
  // stream 0: block-per-vertex classes, largest vertices first
  if (on_lm(kClsH16K)) CVB_TRY((launch_lm_block<1024, 14>(e, e->S(0), q, pos(kClsH16K), cnt(kClsH16K))));
  else if ((n = cnt(kClsH16K)) > 0) CVB_TRY((launch_hash<1024, 4, 2, 1>(e, e->S(0), p, lst(kClsH16K), n)));
  if (on_lm(kClsH8K)) CVB_TRY((launch_lm_block<512, 13>(e, e->S(4), q, pos(kClsH8K), cnt(kClsH8K))));
  else if ((n = cnt(kClsH8K)) > 0) CVB_TRY((launch_hash<512, 4, 2, 2>(e, e->S(4), p, lst(kClsH8K), n)));
  if (on_lm(kClsH4K)) CVB_TRY((launch_lm_block<256, 12>(e, e->S(5), q, pos(kClsH4K), cnt(kClsH4K))));
  else if ((n = cnt(kClsH4K)) > 0) CVB_TRY((launch_hash<256, 4, 2, 4>(e, e->S(5), p, lst(kClsH4K), n)));
  if (on_lm(kClsH2K)) CVB_TRY((launch_lm_block<128, 11>(e, e->S(6), q, pos(kClsH2K), cnt(kClsH2K))));
  else if ((n = cnt(kClsH2K)) > 0) CVB_TRY((launch_hash<128, 4, 2, 8>(e, e->S(6), p, lst(kClsH2K), n)));
  // stream 1: vertices above 8192 entries.  Default: global (L2-resident) tables for all of
  // them; CVB_CLUSTER=1 routes 8193..65536 to thread-block clusters with the table in
  // distributed shared memory (correct, but DSMEM moves ~32 B per 4-byte atomic and
  // measured slower on B200 — DESIGN.md section 5).
  const int gcls = e->global_table_class();
  const bool big = lm && e->big_ready;
  CVB_CHECK(!e->ovr || big || e->class_start[kClsGiant + 1] == e->class_start[kClsC2], CVB_ERR_STATE,
            "colouring needs the default big-vertex path (CVB_BIG / CVB_LM left at their defaults)");
  if (big && !(e->ovr && !e->ovr_big) && (n = e->class_start[kClsGiant + 1] - e->class_start[kClsC2]) > 0) {
    BigParams bp{};
    bp.chunk_owner = e->big_owner.p;
    bp.chunk_start = e->big_layout.p;
    bp.log2p = e->bg_log2p.p;
    bp.item_start = e->bg_item_start.p;
    bp.item_owner = e->bg_item_owner.p;
    bp.records = e->bg_records.p;
    bp.cursor = e->bg_cursor.p;
    bp.own = e->bg_own.p;
    bp.cand_gain = e->bg_cand_gain.p;
    bp.cand_key = e->bg_cand_key.p;
    bp.flags = e->d_flags.p;
    bp.big_first = static_cast<int>(pos(kClsC2));
    LmParams qb = q;
    if (e->ovr) {  // colour step: every big vertex is evaluated, k_color_fix_big keeps the ones of this colour
      qb.vrec = e->vrec.p;
      qb.remap = remap ? 1 : 0;
      qb.ddeg = nullptr;
      qb.dsize = nullptr;
    }
    constexpr size_t kStageSmem = size_t(8) << 13, kPassSmem = size_t(8) << 13;
    auto stage = e->unit_weights ? k_big_stage<true> : k_big_stage<false>;
    const int occ1 = lm_setup(e, reinterpret_cast<const void *>(stage), 512, kStageSmem);
    const int occ2 = lm_setup(e, reinterpret_cast<const void *>(k_big_pass), 512, kPassSmem);
    CVB_CHECK(occ1 > 0 && occ2 > 0, CVB_ERR_CUDA, "big-vertex kernels: cannot configure (%s)",
              cudaGetErrorString(cudaGetLastError()));
    const int64_t cap1 = static_cast<int64_t>(e->num_sms) * occ1, cap2 = static_cast<int64_t>(e->num_sms) * occ2;
    stage<<<static_cast<int>(e->big_chunks < cap1 ? e->big_chunks : cap1), 512, kStageSmem, e->S(1)>>>(qb, bp, e->big_chunks);
    k_big_pass<<<static_cast<int>(e->bg_items < cap2 ? e->bg_items : cap2), 512, kPassSmem, e->S(1)>>>(qb, bp, e->bg_items);
    k_big_final<<<static_cast<int>((n * 32 + 127) / 128), 128, 0, e->S(1)>>>(qb, bp, static_cast<int>(n));
    e->launches += 3;
    if (e->ovr) {
      k_color_fix_big<<<static_cast<int>((n + 127) / 128), 128, 0, e->S(1)>>>(
          e->list.p + e->class_start[kClsC2], static_cast<int>(n), e->col_color.p, e->col_num, e->ovr_color,
          e->lab[e->targ].p, e->lab[e->curr].p, e->cw.p, e->col_bak_t.p, e->col_bak_cw.p, e->vdeg.p, e->ddeg.p,
          e->dsize.p, e->ovr_et ? e->et_active.p : nullptr, e->et_frozen_target.p, e->et_frozen_cw.p);
      e->launches++;
    }
  }
  const bool mp = lm && e->mp_ready && !big;
  if (mp && (n = e->class_start[kClsGiant] - e->class_start[kClsC2]) > 0) {
    constexpr size_t smem = lm_multipass_smem();
    auto go = [&](auto kern) -> int {
      const int occ = lm_setup(e, reinterpret_cast<const void *>(kern), 1024, smem);
      CVB_CHECK(occ > 0, CVB_ERR_CUDA, "lm multipass kernel: cannot configure (%s)",
                cudaGetErrorString(cudaGetLastError()));
      const int grid = static_cast<int>(n < e->num_sms ? n : e->num_sms);
      kern<<<grid, 1024, smem, e->S(1)>>>(q, static_cast<int>(pos(kClsC2)), static_cast<int>(n), e->mp_off.p,
                                            e->mp_stage.p, e->mp_redo.p);
      e->launches++;
      return CVB_OK;
    };
    CVB_TRY(e->unit_weights ? go(k_lm_multipass<true>) : go(k_lm_multipass<false>));
  }
  if (!mp && !big && !e->use_cluster && e->use_twopass && (n = cnt(kClsC2)) > 0) {
    constexpr int kTwoPassSmem = 16384 * 8;
    if (first_use_on_device(e->device, reinterpret_cast<const void *>(&k_sweep_u32_twopass))) {
      CVB_CUDA(cudaFuncSetAttribute(k_sweep_u32_twopass, cudaFuncAttributeMaxDynamicSharedMemorySize, kTwoPassSmem));
    }
    const int grid = static_cast<int>(n < e->num_sms ? n : e->num_sms);
    k_sweep_u32_twopass<<<grid, 1024, kTwoPassSmem, e->S(1)>>>(p, lst(kClsC2), static_cast<int>(n));
    e->launches++;
  }
  if (e->use_cluster && !mp && !big) {
    if ((n = cnt(kClsC8)) > 0) CVB_TRY(launch_cluster<8>(e, e->S(1), p, lst(kClsC8), n));
    if ((n = cnt(kClsC4)) > 0) CVB_TRY(launch_cluster<4>(e, e->S(1), p, lst(kClsC4), n));
    if ((n = cnt(kClsC2)) > 0) CVB_TRY(launch_cluster<2>(e, e->S(1), p, lst(kClsC2), n));
  }
  if ((n = e->class_start[kClsGiant + 1] - e->class_start[gcls]) > 0) {
    GiantTables gt{};
    gt.chunk_owner = e->gchunk_owner.p;
    gt.chunk_start = e->giant_layout.p;
    gt.table_start = e->giant_layout.p + (n + 1);
    gt.slots = e->gslots.p;
    gt.gmeta = e->gmeta.p;
    gt.piece_gain = e->gpiece_gain.p;
    gt.piece_key = e->gpiece_key.p;
    constexpr int kGiantSmem = 2 * kGiantChunk * 8;
    if (first_use_on_device(e->device, reinterpret_cast<const void *>(&k_sweep_u32_giant_insert<false>))) {
      CVB_CUDA(cudaFuncSetAttribute(k_sweep_u32_giant_insert<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGiantSmem));
    }
    gt.dense = e->dense.p;
    gt.dense_stride = e->dense_stride;
    gt.kcount = e->kcount.p;
    gt.klist = e->klist.p;
    if (first_use_on_device(e->device, reinterpret_cast<const void *>(&k_sweep_u32_giant_insert<true>))) {
      CVB_CUDA(cudaFuncSetAttribute(k_sweep_u32_giant_insert<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGiantSmem));
    }
    for (const auto &b : e->giant_batches) {
      gt.slot_base = b.slot_base;
      gt.vert_lo = b.vert_lo;
      const int64_t nch = b.chunk_hi - b.chunk_lo, npc = b.piece_hi - b.piece_lo;
      if (e->big_dense) {
        const int64_t g1 = nch < 2LL * e->num_sms ? nch : 2LL * e->num_sms;
        k_sweep_u32_giant_insert<true><<<static_cast<int>(g1), 512, kGiantSmem, e->S(1)>>>(p, lst(gcls), (int)n, gt,
                                                                                            b.chunk_lo, b.chunk_hi);
        const int64_t g2 = nch < 4LL * e->num_sms ? nch : 4LL * e->num_sms;
        k_sweep_u32_giant_eval_dense<<<static_cast<int>(g2), 512, 0, e->S(1)>>>(
            p, lst(gcls), (int)n, gt, b.chunk_lo, b.chunk_hi, e->gchunk2_gain.p, e->gchunk2_key.p);
        e->launches += 2;
        continue;
      }
      if (e->giant_direct) {
        const int64_t g0 = nch < 4LL * e->num_sms ? nch : 4LL * e->num_sms;
        k_sweep_u32_giant_insert_direct<<<static_cast<int>(g0), 512, 0, e->S(1)>>>(p, lst(gcls), (int)n, gt,
                                                                                    b.chunk_lo, b.chunk_hi);
      } else {
        const int64_t g1 = nch < 2LL * e->num_sms ? nch : 2LL * e->num_sms;
        k_sweep_u32_giant_insert<false><<<static_cast<int>(g1), 512, kGiantSmem, e->S(1)>>>(p, lst(gcls), (int)n, gt,
                                                                                             b.chunk_lo, b.chunk_hi);
      }
      const int64_t g2 = npc < 4LL * e->num_sms ? npc : 4LL * e->num_sms;
      k_sweep_u32_giant_eval<<<static_cast<int>(g2), 512, 0, e->S(1)>>>(p, lst(gcls), (int)n, gt, b.piece_lo,
                                                                         b.piece_hi);
      e->launches += 2;
    }
    if (e->big_dense) {
      k_sweep_u32_giant_final_dense<<<static_cast<int>((n * 32 + 127) / 128), 128, 0, e->S(1)>>>(
          p, lst(gcls), (int)n, gt, e->gchunk2_gain.p, e->gchunk2_key.p);
      e->launches += 1;
    } else {
    k_sweep_u32_giant_final<<<static_cast<int>((n * 32 + 127) / 128), 128, 0, e->S(1)>>>(p, lst(gcls), (int)n,
                                                                                         gt);
    e->launches += 1;
    }
  }
  // stream 2: 129..512
  if (on_lm(kClsH1K)) CVB_TRY((launch_lm_block<64, 10>(e, e->S(2), q, pos(kClsH1K), cnt(kClsH1K))));
  else if ((n = cnt(kClsH1K)) > 0) CVB_TRY((launch_hash<64, 4, 2, 16>(e, e->S(2), p, lst(kClsH1K), n)));
  if (on_lm(kClsH512)) CVB_TRY((launch_lm_group<32, 8, 9>(e, e->S(7), q, pos(kClsH512), cnt(kClsH512))));
  else if ((n = cnt(kClsH512)) > 0) CVB_TRY((launch_hash<32, 4, 2, 4>(e, e->S(7), p, lst(kClsH512), n)));
  // stream 3: 1..128
  if (on_lm(kClsH256)) CVB_TRY((launch_lm_group<16, 8, 8>(e, e->S(3), q, pos(kClsH256), cnt(kClsH256))));
  else if ((n = cnt(kClsH256)) > 0) CVB_TRY((launch_hash<32, 4, 1, 4>(e, e->S(3), p, lst(kClsH256), n)));
  if (on_lm(kClsH128)) CVB_TRY((launch_lm_group<8, 8, 7>(e, e->S(8), q, pos(kClsH128), cnt(kClsH128))));
  else if ((n = cnt(kClsH128)) > 0) CVB_TRY((launch_hash<32, 2, 1, 4>(e, e->S(8), p, lst(kClsH128), n)));
  if (on_lm(kClsG32)) CVB_TRY((launch_lm_group<4, 8, 6>(e, e->S(9), q, pos(kClsG32), cnt(kClsG32))));
  else if ((n = cnt(kClsG32)) > 0) {
    k_sweep_u32_group<32><<<static_cast<int>((n * 32 + 255) / 256), 256, 0, e->S(9)>>>(p, lst(kClsG32), (int)n);
    e->launches++;
  }
  if (on_lm(kClsG16)) CVB_TRY((launch_lm_group<4, 4, 5>(e, e->S(10), q, pos(kClsG16), cnt(kClsG16))));
  else if ((n = cnt(kClsG16)) > 0) {
    k_sweep_u32_group<16><<<static_cast<int>((n * 16 + 255) / 256), 256, 0, e->S(10)>>>(p, lst(kClsG16), (int)n);
    e->launches++;
  }
  if (on_lm(kClsG8)) CVB_TRY((launch_lm_group<2, 4, 4>(e, e->S(11), q, pos(kClsG8), cnt(kClsG8))));
  else if ((n = cnt(kClsG8)) > 0) {
    k_sweep_u32_group<8><<<static_cast<int>((n * 8 + 255) / 256), 256, 0, e->S(11)>>>(p, lst(kClsG8), (int)n);
    e->launches++;
  }
  if ((n = cnt(kClsG4)) > 0) {
    k_sweep_u32_t4<<<static_cast<int>((n + 255) / 256), 256, 0, e->S(9)>>>(p, lst(kClsG4), (int)n);
    e->launches++;
  }
  for (int i = 0; i < e->nstreams; ++i) {
    CVB_CUDA(cudaEventRecord(e->ev_join[i], e->aux[i]));
    CVB_CUDA(cudaStreamWaitEvent(st, e->ev_join[i], 0));
  }
  CVB_CUDA(cudaGetLastError());
  return CVB_OK;
}

// The graph is on the device in device layout (offsets / int32 tails / weights
// in e->mode): allocate the per-vertex state, compute vertex degrees, build the
// degree-class lists and the giant-class tables.
int post_graph(cvb_engine *e) {
  const int64_t nvl = e->nvl;
  cudaStream_t st = e->stream;
  const size_t wbytes = e->mode == CVB_WEIGHTS_U32 ? 4 : 8;
  CVB_TRY(e->d_sums.ensure(4));
  CVB_TRY(e->h_sums.ensure(64));
  CVB_TRY(e->ticket.ensure(4));
  CVB_TRY(e->partials.ensure(2 * 148 * 16 + 16));
  CVB_CUDA(cudaMemsetAsync(e->ticket.p, 0, 4 * sizeof(unsigned int), st));
  // per-vertex / per-community state
  CVB_TRY(e->vdeg.ensure(nvl + 1));
  CVB_TRY(e->cw.ensure(nvl + 1));
  for (int k = 0; k < 3; ++k) CVB_TRY(e->lab[k].ensure(nvl + 1));
  CVB_TRY(e->cdeg.ensure(e->nvg + 1));
  CVB_TRY(e->ddeg.ensure(e->nvg + 1));
  CVB_TRY(e->csize.ensure(e->nvg + 1));
  CVB_TRY(e->dsize.ensure(e->nvg + 1));
  if (e->nvl != e->nvg) {
    CVB_TRY(e->labels_global.ensure(e->nvg + 1));
    CVB_TRY(e->labels_global2.ensure(e->nvg + 1));
    CVB_TRY(e->vdeg_global.ensure(e->nvg + 1));
  }
  if (nvl > 0) {
    if (e->mode == CVB_WEIGHTS_U32)
      k_vertex_degree<uint32_t><<<static_cast<int>((nvl * 32 + 255) / 256), 256, 0, st>>>(
          e->offsets.p, reinterpret_cast<uint32_t *>(e->weights.p), nvl, e->vdeg.p);
    else
      k_vertex_degree<double><<<static_cast<int>((nvl + 255) / 256), 256, 0, st>>>(
          e->offsets.p, reinterpret_cast<double *>(e->weights.p), nvl, e->vdeg.p);
    e->launches++;
    k_sum_f64<<<grid_for(nvl, 512, 148 * 4), 512, 0, st>>>(e->vdeg.p, nvl, e->partials.p, e->ticket.p, e->d_sums.p);
    e->launches++;
  }
  // degree classes
  const int64_t nblocks = (nvl + kClassBlock - 1) / kClassBlock;
  CVB_TRY(e->block_counts.ensure(static_cast<size_t>(nblocks + 1) * kNumClasses));
  CVB_TRY(e->d_class_start.ensure(kNumClasses + 2));
  CVB_TRY(e->d_class_entries.ensure(kNumClasses + 2));
  CVB_TRY(e->list.ensure(nvl + 1));
  CVB_CUDA(cudaMemsetAsync(e->d_class_entries.p, 0, (kNumClasses + 2) * sizeof(unsigned long long), st));
  CVB_CUDA(cudaMemsetAsync(e->d_class_start.p, 0, (kNumClasses + 2) * sizeof(int64_t), st));
  if (nvl > 0) {
    k_class_count<<<static_cast<int>(nblocks), kClassBlock, 0, st>>>(e->offsets.p, nvl, e->block_counts.p,
                                                                   e->d_class_entries.p,
                                                                   e->d_class_entries.p + kNumClasses);
    k_class_scan<<<1, 1024, 0, st>>>(e->block_counts.p, nblocks, e->d_class_start.p);
    k_class_scatter<<<static_cast<int>(nblocks), kClassBlock, 0, st>>>(e->offsets.p, nvl, e->block_counts.p,
                                                                     e->list.p);
    e->launches += 3;
  }
  unsigned long long ent[kNumClasses + 1];
  CVB_CUDA(cudaMemcpyAsync(e->class_start, e->d_class_start.p, (kNumClasses + 1) * sizeof(int64_t),
                           cudaMemcpyDeviceToHost, st));
  CVB_CUDA(cudaMemcpyAsync(ent, e->d_class_entries.p, (kNumClasses + 1) * sizeof(unsigned long long),
                           cudaMemcpyDeviceToHost, st));
  double lw = 0.0;
  if (nvl > 0) CVB_CUDA(cudaMemcpyAsync(&lw, e->d_sums.p, sizeof(double), cudaMemcpyDeviceToHost, st));
  CVB_CUDA(cudaStreamSynchronize(st));
  CVB_CUDA(cudaGetLastError());
  for (int c = 0; c < kNumClasses; ++c) e->class_entries[c] = static_cast<int64_t>(ent[c]);
  e->max_degree = static_cast<int64_t>(ent[kNumClasses]);
  e->local_weight = lw;
  // global-memory tables for the giant class
  const int64_t nbig = e->class_start[kClsGiant + 1] - e->class_start[kClsC2];
  if (nbig > 0 && e->mode == CVB_WEIGHTS_F64) {
    int lg = 14;
    while ((int64_t(1) << lg) < 2 * e->max_degree) ++lg;
    e->giant_log2cap = lg;
    e->giant_stride = ((sizeof(vid_t) + wbytes) << lg);
    e->giant_blocks = static_cast<int>(nbig < 2 * e->num_sms ? nbig : 2 * e->num_sms);
    CVB_TRY(e->giant.ensure(e->giant_stride * e->giant_blocks));
  }
  if (e->mode == CVB_WEIGHTS_U32) {
    CVB_TRY(e->packed.ensure(e->nvg + 1));
    if (e->l2_persist && e->nvg > 0) {
      // the label gather only runs at L2 speed if the gathered array lives there; for graphs
      // whose `packed` array approaches the L2 size, pin (part of) it with an access-policy window on every
      // stream the sweep kernels run on; everything else (the CSR stream) is marked streaming on a miss
      cudaDeviceProp prop;
      CVB_CUDA(cudaGetDeviceProperties(&prop, e->device));
      const size_t bytes = static_cast<size_t>(e->nvg) * sizeof(uint2);
      size_t carve = bytes < static_cast<size_t>(prop.persistingL2CacheMaxSize) ? bytes : prop.persistingL2CacheMaxSize;
      if (carve > 0 && prop.accessPolicyMaxWindowSize > 0) {
        CVB_CUDA(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve));
        cudaStreamAttrValue attr{};
        attr.accessPolicyWindow.base_ptr = e->packed.p;
        attr.accessPolicyWindow.num_bytes =
            bytes < static_cast<size_t>(prop.accessPolicyMaxWindowSize) ? bytes : prop.accessPolicyMaxWindowSize;
        const double ratio = static_cast<double>(carve) / static_cast<double>(attr.accessPolicyWindow.num_bytes);
        attr.accessPolicyWindow.hitRatio = static_cast<float>(ratio > 1.0 ? 1.0 : ratio);
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        CVB_CUDA(cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr));
        for (int i = 0; i < cvb_engine::kAux; ++i)
          CVB_CUDA(cudaStreamSetAttribute(e->aux[i], cudaStreamAttributeAccessPolicyWindow, &attr));
      }
    }
    CVB_TRY(e->d_flags.ensure(4));
    CVB_CUDA(cudaMemsetAsync(e->d_flags.p, 0, 4 * sizeof(int), st));
    if (e->nvl > 0) {
      k_check_simple_rows<<<static_cast<int>((e->nvl * 32 + 255) / 256), 256, 0, st>>>(e->offsets.p, e->tails.p, e->nvl,
                                                                                     e->d_flags.p);
      e->launches++;
    }
    // lm.cuh: self-loop weights, unit-weight flag, per-vertex records in class-list order
    e->lm_ready = false;
    if (e->use_lm && e->nvl > 0 && e->ne < (int64_t(1) << 32)) {
      CVB_TRY(e->selfw.ensure(e->nvl + 1));
      CVB_TRY(e->vrec.ensure(e->nvl + 1));
      k_self_loops<<<static_cast<int>((e->nvl * 32 + 255) / 256), 256, 0, st>>>(
          e->offsets.p, e->tails.p, reinterpret_cast<const uint32_t *>(e->weights.p), e->nvl,
          static_cast<vid_t>(e->base), e->selfw.p, e->d_flags.p);
      k_build_vrec<<<grid_for(e->nvl, 256), 256, 0, st>>>(e->list.p, e->offsets.p, e->vdeg.p, e->nvl, e->vrec.p);
      e->launches += 2;
      e->lm_ready = true;
    }
    // multi-pass kernel (8193..65536 entries): staging area + per-vertex offsets
    e->mp_ready = false;
    if (e->lm_ready && e->use_mp) {
      const int64_t nmp = e->class_start[kClsGiant] - e->class_start[kClsC2];
      int64_t ent = 0;
      for (int c = kClsC2; c < kClsGiant; ++c) ent += e->class_entries[c];
      if (nmp > 0) {
        CVB_TRY(e->mp_stage.ensure(static_cast<size_t>(ent) + 1));
        CVB_TRY(e->mp_off.ensure(nmp + 1));
        CVB_TRY(e->mp_deg.ensure(nmp + 1));
        CVB_TRY(e->mp_redo.ensure(4));
        CVB_CUDA(cudaMemsetAsync(e->mp_redo.p, 0, 4 * sizeof(int), st));
        k_mp_degrees<<<grid_for(nmp, 256), 256, 0, st>>>(e->vrec.p + e->class_start[kClsC2], static_cast<int>(nmp),
                                                        e->mp_deg.p);
        size_t tb = 0;
        CVB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, e->mp_deg.p, e->mp_off.p, static_cast<int>(nmp), st));
        CVB_TRY(e->staging2.ensure(tb + 16));
        CVB_CUDA(cub::DeviceScan::ExclusiveSum(e->staging2.p, tb, e->mp_deg.p, e->mp_off.p, static_cast<int>(nmp), st));
        e->launches += 2;
      }
      e->mp_ready = true;
    }
    // degree-ordered gather layout: one GPU only (the class list covers every global vertex)
    e->remap = false;
    if (e->lm_ready && e->use_remap && e->nvl == e->nvg) {
      CVB_TRY(e->rank.ensure(e->nvg + 1));
      CVB_TRY(e->tails_r.ensure(static_cast<size_t>(e->ne) + 1));
      k_rank_from_list<<<grid_for(e->nvl, 256), 256, 0, st>>>(e->list.p, e->nvl, e->rank.p);
      if (e->ne > 0)
        k_remap_tails<<<grid_for(e->ne, 256), 256, 0, st>>>(e->tails.p, e->rank.p, e->ne, e->tails_r.p);
      e->launches += 2;
      e->remap = true;
    }
    int hflags[4] = {0, 0, 0, 0};
    CVB_CUDA(cudaMemcpyAsync(hflags, e->d_flags.p, 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
    CVB_CUDA(cudaStreamSynchronize(st));
    e->rows_simple = hflags[0] == 0;
    e->has_selfloops = hflags[1] != 0;
    e->unit_weights = e->lm_ready && hflags[2] == 0;
    e->big_chunks = 0;
    if (nbig > 0) {  // chunk layout over every vertex above 8192 entries (first-sweep kernels)
      CVB_TRY(e->big_layout.ensure(2 * (nbig + 1)));
      k_giant_layout<<<1, 32, 0, st>>>(e->offsets.p, e->list.p + e->class_start[kClsC2], (int)nbig, e->big_layout.p,
                                       e->big_layout.p + (nbig + 1));
      e->launches++;
      int64_t totc = 0;
      CVB_CUDA(cudaMemcpyAsync(&totc, e->big_layout.p + nbig, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
      CVB_CUDA(cudaStreamSynchronize(st));
      e->big_chunks = totc;
      CVB_TRY(e->big_owner.ensure(totc + 1));
      CVB_TRY(e->big_meta.ensure(3 * nbig + 4));
      CVB_TRY(e->gchunk_gain.ensure(totc + 1));
      CVB_TRY(e->gchunk_key.ensure(totc + 1));
      CVB_CUDA(cudaMemsetAsync(e->big_meta.p, 0, (3 * nbig + 4) * sizeof(uint32_t), st));
      k_giant_chunk_owner<<<static_cast<int>(nbig), 64, 0, st>>>(e->big_layout.p, (int)nbig, e->big_owner.p);
      e->launches++;
    }
    // "big" path (partition then hash): partitions, bucket layout and scratch of every vertex above 8192 entries
    e->big_ready = false;
    e->bg_items = 0;
    if (e->lm_ready && e->use_big && nbig > 0 && e->max_degree <= (int64_t(1) << 20)) {
      std::vector<uint32_t> hdeg(nbig);
      CVB_TRY(e->mp_deg.ensure(nbig + 1));
      k_mp_degrees<<<grid_for(nbig, 256), 256, 0, st>>>(e->vrec.p + e->class_start[kClsC2], static_cast<int>(nbig),
                                                      e->mp_deg.p);
      e->launches++;
      CVB_CUDA(cudaMemcpyAsync(hdeg.data(), e->mp_deg.p, nbig * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
      CVB_CUDA(cudaStreamSynchronize(st));
      std::vector<int32_t> lp(nbig), istart(nbig + 1), iowner;
      int items = 0;
      for (int64_t i = 0; i < nbig; ++i) {
        int l = 0;
        while ((static_cast<int64_t>(kBigPartTarget) << l) < static_cast<int64_t>(hdeg[i])) ++l;
        lp[i] = l;
        istart[i] = items;
        items += 1 << l;
      }
      istart[nbig] = items;
      iowner.resize(items);
      for (int64_t i = 0; i < nbig; ++i)
        for (int q = istart[i]; q < istart[i + 1]; ++q) iowner[q] = static_cast<int32_t>(i);
      CVB_TRY(e->bg_log2p.ensure(nbig + 1));
      CVB_TRY(e->bg_item_start.ensure(nbig + 2));
      CVB_TRY(e->bg_item_owner.ensure(items + 1));
      CVB_TRY(e->bg_records.ensure(static_cast<size_t>(items) * kBigBucketCap + 1));
      CVB_TRY(e->bg_cursor.ensure(items + 1));
      CVB_TRY(e->bg_own.ensure(nbig + 1));
      CVB_TRY(e->bg_cand_key.ensure(items + 1));
      CVB_TRY(e->bg_cand_gain.ensure(items + 1));
      CVB_CUDA(cudaMemcpyAsync(e->bg_log2p.p, lp.data(), nbig * sizeof(int32_t), cudaMemcpyHostToDevice, st));
      CVB_CUDA(cudaMemcpyAsync(e->bg_item_start.p, istart.data(), (nbig + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, st));
      CVB_CUDA(cudaMemcpyAsync(e->bg_item_owner.p, iowner.data(), items * sizeof(int32_t), cudaMemcpyHostToDevice, st));
      CVB_CUDA(cudaMemsetAsync(e->bg_cursor.p, 0, (items + 1) * sizeof(uint32_t), st));
      CVB_CUDA(cudaMemsetAsync(e->bg_own.p, 0, (nbig + 1) * sizeof(uint32_t), st));
      CVB_CUDA(cudaStreamSynchronize(st));  // the host vectors go out of scope
      e->bg_items = items;
      e->big_ready = true;
    }
    const int gcls = e->global_table_class();
    const int64_t ngiant = e->class_start[kClsGiant + 1] - e->class_start[gcls];  // global-table vertices
    if (ngiant > 0) {
      CVB_TRY(e->giant_layout.ensure(2 * (ngiant + 1)));
      k_giant_layout<<<1, 32, 0, st>>>(e->offsets.p, e->list.p + e->class_start[gcls], (int)ngiant,
                                       e->giant_layout.p, e->giant_layout.p + (ngiant + 1));
      e->launches++;
      std::vector<int64_t> lay(2 * (ngiant + 1));
      CVB_CUDA(cudaMemcpyAsync(lay.data(), e->giant_layout.p, lay.size() * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
      CVB_CUDA(cudaStreamSynchronize(st));
      const int64_t *cs = lay.data(), *ts = lay.data() + (ngiant + 1);
      e->giant_chunks = cs[ngiant];
      e->giant_slots = ts[ngiant];
      e->giant_batches.clear();
      int64_t arena = 0;
      if (e->big_dense) {
        // batches of A consecutive big vertices, one dense accumulator each
        e->dense_stride = (e->nvg + 31) & ~int64_t(31);
        int64_t A = e->dense_budget_bytes / (e->dense_stride * 4);
        A = A < 1 ? 1 : (A > ngiant ? ngiant : A);
        for (int64_t i = 0; i < ngiant; i += A) {
          const int64_t j = i + A < ngiant ? i + A : ngiant;
          e->giant_batches.push_back({cs[i], cs[j], 0, 0, 0, static_cast<int>(i)});
        }
        CVB_TRY(e->dense.ensure(static_cast<size_t>(A) * e->dense_stride));
        CVB_CUDA(cudaMemsetAsync(e->dense.p, 0, static_cast<size_t>(A) * e->dense_stride * sizeof(uint32_t), st));
        CVB_TRY(e->kcount.ensure(ngiant + 1));
        CVB_CUDA(cudaMemsetAsync(e->kcount.p, 0, (ngiant + 1) * sizeof(uint32_t), st));
        CVB_TRY(e->klist.ensure(static_cast<size_t>(e->giant_chunks) * kGiantChunk + 1));
        CVB_TRY(e->gchunk2_gain.ensure(e->giant_chunks + 1));
        CVB_TRY(e->gchunk2_key.ensure(e->giant_chunks + 1));
        arena = kGiantPiece;  // the hash-table arena is unused
      } else {
        // batches of consecutive big vertices whose tables fit the arena together
        for (int64_t i = 0; i < ngiant;) {
          int64_t j = i + 1;
          while (j < ngiant && ts[j + 1] - ts[i] <= e->giant_arena_slots) ++j;
          e->giant_batches.push_back({cs[i], cs[j], ts[i] / kGiantPiece, ts[j] / kGiantPiece, ts[i], static_cast<int>(i)});
          if (ts[j] - ts[i] > arena) arena = ts[j] - ts[i];
          i = j;
        }
      }
      CVB_TRY(e->gslots.ensure(arena));
      CVB_TRY(e->gmeta.ensure(3 * ngiant + 4));
      CVB_TRY(e->gpiece_gain.ensure(e->giant_slots / kGiantPiece + 1));
      CVB_TRY(e->gpiece_key.ensure(e->giant_slots / kGiantPiece + 1));
      CVB_TRY(e->gchunk_owner.ensure(e->giant_chunks + 1));
      k_giant_chunk_owner<<<static_cast<int>(ngiant), 64, 0, st>>>(e->giant_layout.p, (int)ngiant, e->gchunk_owner.p);
      e->launches++;
      k_giant_init_slots<<<grid_for(arena, 256), 256, 0, st>>>(e->gslots.p, arena);
      e->launches++;
      CVB_CUDA(cudaMemsetAsync(e->gmeta.p, 0, (3 * ngiant + 4) * sizeof(uint32_t), st));
    }
  }
  e->graph_ready = true;
  e->phase_ready = false;
  return CVB_OK;
}

// Everything after the CSR is on the device in boundary layout: pick the
// weight mode, narrow, compute vertex degrees, build the degree-class lists.
int finish_graph(cvb_engine *e, const void *d_aos, const int64_t *d_tails64, const double *d_w64) {
  const int64_t ne = e->ne, nvl = e->nvl;
  cudaStream_t st = e->stream;
  CVB_TRY(e->d_flags.ensure(4));
  CVB_TRY(e->d_sums.ensure(4));
  CVB_TRY(e->h_sums.ensure(64));
  CVB_TRY(e->ticket.ensure(4));
  CVB_TRY(e->partials.ensure(2 * 148 * 16 + 16));
  CVB_CUDA(cudaMemsetAsync(e->d_flags.p, 0, 4 * sizeof(int), st));
  CVB_CUDA(cudaMemsetAsync(e->d_sums.p, 0, 4 * sizeof(double), st));
  CVB_CUDA(cudaMemsetAsync(e->ticket.p, 0, 4 * sizeof(unsigned int), st));
  if (ne > 0) {
    if (d_aos)
      k_weight_stats<true><<<grid_for(ne, 256), 256, 0, st>>>(d_aos, ne, e->d_flags.p, e->d_sums.p);
    else
      k_weight_stats<false><<<grid_for(ne, 256), 256, 0, st>>>(d_w64, ne, e->d_flags.p, e->d_sums.p);
    e->launches++;
  }
  int flags = 0;
  double wsum = 0.0;
  CVB_CUDA(cudaMemcpyAsync(&flags, e->d_flags.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  CVB_CUDA(cudaMemcpyAsync(&wsum, e->d_sums.p, sizeof(double), cudaMemcpyDeviceToHost, st));
  CVB_CUDA(cudaStreamSynchronize(st));
  e->mode = (flags == 0 && wsum < 4294967296.0) ? CVB_WEIGHTS_U32 : CVB_WEIGHTS_F64;
  const size_t wbytes = e->mode == CVB_WEIGHTS_U32 ? 4 : 8;
  CVB_TRY(e->tails.ensure(static_cast<size_t>(ne) + 1));
  CVB_TRY(e->weights.ensure((static_cast<size_t>(ne) + 1) * wbytes));
  if (ne > 0) {
    const int g = grid_for(ne, 256);
    if (e->mode == CVB_WEIGHTS_U32) {
      if (d_aos)
        k_convert_edges<true, uint32_t><<<g, 256, 0, st>>>(d_aos, nullptr, nullptr, ne, e->tails.p,
                                                           reinterpret_cast<uint32_t *>(e->weights.p));
      else
        k_convert_edges<false, uint32_t><<<g, 256, 0, st>>>(nullptr, d_tails64, d_w64, ne, e->tails.p,
                                                            reinterpret_cast<uint32_t *>(e->weights.p));
    } else {
      if (d_aos)
        k_convert_edges<true, double><<<g, 256, 0, st>>>(d_aos, nullptr, nullptr, ne, e->tails.p,
                                                         reinterpret_cast<double *>(e->weights.p));
      else
        k_convert_edges<false, double><<<g, 256, 0, st>>>(nullptr, d_tails64, d_w64, ne, e->tails.p,
                                                          reinterpret_cast<double *>(e->weights.p));
    }
    e->launches++;
  }
  return post_graph(e);
}

int check_sizes(int64_t nv_local, int64_t base, int64_t nv_global) {
  CVB_CHECK(nv_local >= 0 && base >= 0 && nv_global >= nv_local && base + nv_local <= nv_global, CVB_ERR_ARG,
            "bad partition: nv_local=%lld base=%lld nv_global=%lld", (long long)nv_local, (long long)base,
            (long long)nv_global);
  CVB_CHECK(nv_global < 2147483000LL, CVB_ERR_CAPACITY,
            "nv_global=%lld exceeds the 32-bit device id range", (long long)nv_global);
  return CVB_OK;
}

}  // namespace

extern "C" {

const char *cvb_last_error(void) { return g_error.c_str(); }
const char *cvb_version(void) { return "cuvite_b200 0.1.0 (sm_100a)"; }

int cvb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int cvb_set_device(int local_rank) {
  int n = 0;
  CVB_CUDA(cudaGetDeviceCount(&n));
  CVB_CHECK(n > 0, CVB_ERR_CUDA, "no CUDA device visible");
  CVB_CUDA(cudaSetDevice(local_rank % n));
  return CVB_OK;
}

cvb_engine *cvb_create(void *stream) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    set_error("cvb_create: no usable CUDA device (%s)", cudaGetErrorString(cudaGetLastError()));
    return nullptr;
  }
  cvb_engine *e = new cvb_engine;
  e->device = dev;
  if (const char *s = getenv("CVB_NO_FIRST_SWEEP")) e->use_first_sweep = !(s[0] == '1');
  if (const char *s = getenv("CVB_GIANT_DIRECT")) e->giant_direct = (s[0] == '1');
  if (const char *s = getenv("CVB_GIANT_ARENA_MB")) {
    const long mb = atol(s);
    if (mb > 0) e->giant_arena_slots = static_cast<int64_t>(mb) << 17;
  }
  if (const char *s = getenv("CVB_CLUSTER")) e->use_cluster = (s[0] == '1');
  if (const char *s = getenv("CVB_NO_TWOPASS")) e->use_twopass = (s[0] != '1');
  if (const char *s = getenv("CVB_L2_PERSIST")) e->l2_persist = (s[0] == '1');
  if (const char *s = getenv("CVB_SPLIT")) e->use_split = (s[0] == '1');
  if (const char *s = getenv("CVB_LM")) e->use_lm = (s[0] != '0');
  if (const char *s = getenv("CVB_HASH2")) e->use_hash2 = (s[0] != '0');
  if (const char *s = getenv("CVB_STREAMS")) {
    const int k = atoi(s);
    if (k >= 1 && k <= cvb_engine::kAux) e->nstreams = k;
  }
  if (const char *s = getenv("CVB_LM_MASK")) e->lm_mask = static_cast<uint32_t>(strtoul(s, nullptr, 0));
  if (const char *s = getenv("CVB_REMAP")) e->use_remap = (s[0] != '0');
  if (const char *s = getenv("CVB_KEEP_SCRATCH")) e->keep_scratch = (s[0] != '0');
  if (const char *s = getenv("CVB_MP")) e->use_mp = (s[0] == '1');
  if (const char *s = getenv("CVB_BIG")) e->use_big = (s[0] != '0');
  if (const char *s = getenv("CVB_BIG_DENSE")) e->big_dense = (s[0] != '0');
  if (const char *s = getenv("CVB_BIG_DENSE_MB")) {
    const long mb = atol(s);
    if (mb > 0) e->dense_budget_bytes = static_cast<int64_t>(mb) << 20;
  }
  if (const char *s = getenv("CVB_TMA")) e->use_tma = (s[0] == '1');
  int sms = 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0) e->num_sms = sms;
  if (stream) {
    e->stream = static_cast<cudaStream_t>(stream);
  } else {
    if (cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess) {
      set_error("cvb_create: cudaStreamCreate failed (%s)", cudaGetErrorString(cudaGetLastError()));
      delete e;
      return nullptr;
    }
    e->own_stream = true;
  }
  if (cudaEventCreate(&e->ev0) != cudaSuccess || cudaEventCreate(&e->ev1) != cudaSuccess ||
      cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming) != cudaSuccess) {
    set_error("cvb_create: cudaEventCreate failed");
    delete e;
    return nullptr;
  }
  for (int i = 0; i < cvb_engine::kAux; ++i) {
    if (cudaStreamCreateWithFlags(&e->aux[i], cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&e->ev_join[i], cudaEventDisableTiming) != cudaSuccess) {
      set_error("cvb_create: auxiliary stream/event creation failed");
      delete e;
      return nullptr;
    }
  }
  return e;
}

void cvb_destroy(cvb_engine *e) {
  if (!e) return;
  cudaStreamSynchronize(e->stream);
  e->offsets.release(); e->tails.release(); e->weights.release();
  e->vdeg.release(); e->cdeg.release(); e->ddeg.release(); e->cw.release();
  e->csize.release(); e->dsize.release();
  for (auto &l : e->lab) l.release();
  e->labels_global.release(); e->labels_global2.release(); e->vdeg_global.release(); e->list.release(); e->block_counts.release();
  e->d_class_start.release(); e->d_class_entries.release();
  e->partials.release(); e->d_sums.release(); e->ticket.release(); e->d_flags.release();
  e->h_sums.release(); e->staging.release(); e->staging2.release(); e->giant.release(); e->h_stage.release();
  e->packed.release(); e->giant_layout.release(); e->gslots.release(); e->pair_keys.release(); e->pair_w.release();
  e->cs.release();
  e->col_color.release(); e->col_keys.release(); e->col_keys2.release(); e->col_list.release(); e->col_ids.release();
  e->col_vrec.release(); e->col_bak_t.release(); e->col_bak_cw.release();
  e->cdeg32.release(); e->vrec.release(); e->selfw.release(); e->rank.release(); e->tails_r.release();
  e->bg_log2p.release(); e->bg_item_start.release(); e->bg_item_owner.release(); e->bg_records.release();
  e->bg_cursor.release(); e->bg_own.release(); e->bg_cand_key.release(); e->bg_cand_gain.release();
  e->mp_stage.release(); e->mp_off.release(); e->mp_deg.release(); e->mp_redo.release();
  e->dense.release(); e->kcount.release(); e->klist.release(); e->gchunk2_gain.release(); e->gchunk2_key.release();
  e->gmeta.release(); e->gpiece_key.release(); e->gpiece_gain.release(); e->gchunk_gain.release();
  e->gchunk_key.release(); e->gchunk_owner.release(); e->big_layout.release(); e->big_owner.release();
  e->big_meta.release();
  for (int i = 0; i < cvb_engine::kAux; ++i) {
    if (e->aux[i]) cudaStreamDestroy(e->aux[i]);
    if (e->ev_join[i]) cudaEventDestroy(e->ev_join[i]);
  }
  if (e->ev_fork) cudaEventDestroy(e->ev_fork);
  if (e->ev0) cudaEventDestroy(e->ev0);
  if (e->ev1) cudaEventDestroy(e->ev1);
  for (auto &ev : e->up_ev)
    if (ev) cudaEventDestroy(ev);
  if (e->own_stream) cudaStreamDestroy(e->stream);
  delete e;
}

// Host-side narrowing of the boundary layout (the reference's 64-bit Edge{tail, weight}, 16 bytes per
// entry) to what the device keeps — int32 tails, and uint32 weights only when some weight is not 1 — so
// that the PCIe bill of an upload is 4 (or 8) bytes per entry instead of 16: r01's end-to-end number
// was bound by exactly that copy.  Worker threads narrow fixed-size chunks into pinned staging while the
// copies of the finished chunks are already in flight.  Returns 1 when the device arrays are filled and
// e->mode is set, 0 when the graph needs the general path (a non-integer / negative / huge weight).
// Host worker pool of the narrowing upload: created once per process (thread creation costs more than
// narrowing a 4 M-entry chunk), workers sleep between uploads.  A job is `npieces` independent pieces handed
// out through an atomic counter; the caller learns about finished pieces through `done`.
namespace {
struct NarrowPool {
  std::mutex m;
  std::condition_variable cv;
  std::vector<std::thread> th;
  uint64_t gen = 0;
  std::function<void(int64_t)> fn;
  int64_t npieces = 0;
  std::atomic<int64_t> next{0};
  std::atomic<int> active{0};

  explicit NarrowPool(int nt) {
    for (int t = 0; t < nt; ++t) {
      th.emplace_back([this]() {
        uint64_t seen = 0;
        while (true) {
          std::unique_lock<std::mutex> lk(m);
          cv.wait(lk, [&] { return gen != seen; });
          seen = gen;
          lk.unlock();
          for (int64_t p = next.fetch_add(1); p < npieces; p = next.fetch_add(1)) fn(p);
          active.fetch_sub(1);
        }
      });
      th.back().detach();  // the pool lives until the process ends
    }
  }
  // starts the job and returns; the workers are busy until active == 0
  void start(int64_t n, std::function<void(int64_t)> f) {
    std::lock_guard<std::mutex> lk(m);
    fn = std::move(f);
    npieces = n;
    next.store(0);
    active.store(static_cast<int>(th.size()));
    ++gen;
    cv.notify_all();
  }
  void wait() {
    while (active.load() != 0) std::this_thread::yield();
  }
};

int narrow_threads() {
  static const int nt = [] {
    if (const char *s = getenv("CVB_UPLOAD_THREADS")) return atoi(s) > 0 ? atoi(s) : 1;
    const unsigned hw = std::thread::hardware_concurrency();
    // measured on the round-2 host (2 x 32 cores, SMT): e2e 3.9 / 4.9 / 2.3 / 1.7 G edges/s with 16 / 32 / 64 /
    // 128 workers (profiles/r02_e2e_narrowing.txt) — one socket's worth of cores, never more than 32
    const int n = static_cast<int>(hw / 4);
    return n < 1 ? 1 : (n > 32 ? 32 : n);
  }();
  return nt;
}
NarrowPool &narrow_pool() {
  static NarrowPool *pool = new NarrowPool(narrow_threads());  // intentionally never destroyed
  return *pool;
}
}  // namespace

// Host-side narrowing of the boundary arrays (int64 tail + double weight, 16 B per entry) to the device layout
// (int32 tail, 4 B per entry; uint32 weights only when some weight differs from 1) BEFORE the PCIe transfer: the
// worker pool converts 256 K-entry pieces into pinned staging, and every finished 4 M-entry chunk goes out with
// an asynchronous copy while the next ones are converted.  Returns 1 when the graph was uploaded this way, 0 when
// the caller should use the plain path (real weights, ids beyond int32, no staging memory, small graph).
static int upload_narrowed(cvb_engine *e, const void *edges_aos, const int64_t *tails, const double *weights,
                           int64_t ne) {
  if (ne < (1 << 20)) return 0;
  {
    // Default: pageable sources only (what the reference's std::vector storage is) — there the driver would
    // stage the 16-byte entries through its own bounce buffers.  CVB_NARROW_UPLOAD=1 / 0 forces it on / off.
    const char *s = getenv("CVB_NARROW_UPLOAD");
    bool want;
    if (s && (s[0] == '0' || s[0] == '1')) {
      want = s[0] == '1';
    } else {
      cudaPointerAttributes at{};
      const void *src = edges_aos ? edges_aos : static_cast<const void *>(tails);
      want = !(cudaPointerGetAttributes(&at, src) == cudaSuccess && at.type == cudaMemoryTypeHost);
      cudaGetLastError();
    }
    if (!want) return 0;
  }
  // Staging: a ring of kRing pinned slots of one chunk each (tails + weights halves) — pinning memory for the whole
  // edge list cost more than the transfer on a first call (0.16 s for the 31 M entries of R-MAT scale 20 against
  // 0.02 s steady state).  Workers may run at most kRing chunks ahead of the oldest unfinished copy.
  constexpr int64_t kPiece = int64_t(1) << 18, kPiecesPerChunk = 4, kChunk = kPiece * kPiecesPerChunk;  // 1 M entries
  constexpr int kRing = 8;
  if (e->h_stage.ensure(static_cast<size_t>(kRing) * kChunk * 8 + 64) != CVB_OK) { cudaGetLastError(); return 0; }
  if (e->tails.ensure(static_cast<size_t>(ne) + 1) != CVB_OK) return 0;
  if (e->weights.ensure((static_cast<size_t>(ne) + 1) * 4) != CVB_OK) return 0;
  for (auto &ev : e->up_ev)
    if (!ev && cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return 0; }
  unsigned char *stage = e->h_stage.p;
  const EdgeAoS *aos = static_cast<const EdgeAoS *>(edges_aos);
  const int64_t npieces = (ne + kPiece - 1) / kPiece;
  const int64_t nchunks = (npieces + kPiecesPerChunk - 1) / kPiecesPerChunk;
  struct Stat { std::atomic<int> bad{0}, nonunit{0}; std::atomic<uint64_t> sum{0}; std::atomic<int64_t> released{0}; } stat;
  std::vector<std::atomic<int>> done(nchunks), chunk_nonunit(nchunks);
  for (auto &d : done) d.store(0);
  for (auto &d : chunk_nonunit) d.store(0);
  // unit weights everywhere first; the chunks that hold other weights are copied over it (stream order)
  k_fill_u32<<<grid_for(ne, 256), 256, 0, e->stream>>>(reinterpret_cast<uint32_t *>(e->weights.p), ne, 1u);
  e->launches++;
  static std::mutex one_job;  // the pool is process-wide: uploads of several engines take turns
  std::lock_guard<std::mutex> turn(one_job);
  NarrowPool &pool = narrow_pool();
  pool.start(npieces, [=, &stat, &done, &chunk_nonunit](int64_t p) {
    const int64_t c = p / kPiecesPerChunk;
    while (c >= stat.released.load() + kRing) std::this_thread::yield();  // the slot is still being copied out
    int32_t *ht = reinterpret_cast<int32_t *>(stage + static_cast<size_t>(c % kRing) * kChunk * 8);
    uint32_t *hw = reinterpret_cast<uint32_t *>(ht + kChunk);
    const int64_t lo = p * kPiece, hi = lo + kPiece < ne ? lo + kPiece : ne, c0 = c * kChunk;
    // branch-free body: range / integrality violations only set bits, sums are exact integers
    int64_t bad = 0, nonunit = 0;
    uint64_t sum = 0;
    for (int64_t i = lo; i < hi; ++i) {
      const int64_t tl = aos ? aos[i].tail : tails[i];
      const double w = aos ? aos[i].weight : weights[i];
      const int64_t wi = static_cast<int64_t>(w);  // NaN / out of range -> INT64_MIN: caught by the range test
      bad |= (wi >> 32) | (static_cast<double>(wi) != w) | (tl < 0) | (tl > 2147483000LL);
      nonunit |= wi ^ 1;
      sum += static_cast<uint64_t>(wi) & 0xffffffffu;
      ht[i - c0] = static_cast<int32_t>(tl);
      hw[i - c0] = static_cast<uint32_t>(wi);
    }
    if (bad) stat.bad.store(1);
    if (nonunit) {
      stat.nonunit.store(1);
      chunk_nonunit[c].store(1);
    }
    stat.sum.fetch_add(sum);
    done[c].fetch_add(1);
  });
  bool ok = true;
  for (int64_t c = 0; c < nchunks && ok; ++c) {
    const int64_t p0 = c * kPiecesPerChunk, p1 = p0 + kPiecesPerChunk < npieces ? p0 + kPiecesPerChunk : npieces;
    while (done[c].load() < static_cast<int>(p1 - p0)) std::this_thread::yield();
    if (stat.bad.load()) { ok = false; break; }
    const int64_t c0 = c * kChunk, c1 = c0 + kChunk < ne ? c0 + kChunk : ne;
    const int32_t *ht = reinterpret_cast<const int32_t *>(stage + static_cast<size_t>(c % kRing) * kChunk * 8);
    ok = cudaMemcpyAsync(e->tails.p + c0, ht, (c1 - c0) * sizeof(int32_t), cudaMemcpyHostToDevice, e->stream) == cudaSuccess;
    if (ok && chunk_nonunit[c].load())
      ok = cudaMemcpyAsync(reinterpret_cast<uint32_t *>(e->weights.p) + c0, ht + kChunk, (c1 - c0) * sizeof(uint32_t),
                           cudaMemcpyHostToDevice, e->stream) == cudaSuccess;
    ok = ok && cudaEventRecord(e->up_ev[c % kRing], e->stream) == cudaSuccess;
    // free the slots whose copies are complete, keeping at most kRing - 2 copies in flight
    while (ok && stat.released.load() + (kRing - 2) <= c) {
      const int64_t r = stat.released.load();
      ok = cudaEventSynchronize(e->up_ev[r % kRing]) == cudaSuccess;
      stat.released.store(r + 1);
    }
  }
  stat.released.store(nchunks + kRing);  // error path: let waiting workers run to the end of the job
  pool.wait();  // the job's captures (stat, done) must outlive the workers' last piece
  if (!ok || stat.bad.load()) {
    cudaStreamSynchronize(e->stream);
    cudaGetLastError();
    return 0;
  }
  if (static_cast<double>(stat.sum.load()) >= 4294967296.0) return 0;
  e->mode = CVB_WEIGHTS_U32;
  return 1;
}

int cvb_graph_upload_host(cvb_engine *e, int64_t nv_local, int64_t base, int64_t nv_global,
                          const int64_t *offsets, const void *edges_aos, const int64_t *tails,
                          const double *weights) {
  CVB_CHECK(e, CVB_ERR_ARG, "null engine");
  CVB_TRY(check_sizes(nv_local, base, nv_global));
  CVB_CHECK(offsets && (edges_aos || (tails && weights) || offsets[nv_local] == offsets[0]), CVB_ERR_ARG,
            "cvb_graph_upload_host: missing arrays");
  const int64_t ne = offsets[nv_local] - offsets[0];
  CVB_CHECK(ne >= 0 && offsets[0] == 0, CVB_ERR_ARG, "offsets must be rebased to 0 and non-decreasing");
  e->nvl = nv_local; e->base = base; e->nvg = nv_global; e->ne = ne;
  e->graph_ready = false;
  CVB_TRY(e->offsets.ensure(nv_local + 2));
  CVB_CUDA(cudaMemcpyAsync(e->offsets.p, offsets, (nv_local + 1) * sizeof(int64_t), cudaMemcpyHostToDevice,
                           e->stream));
  if (upload_narrowed(e, edges_aos, tails, weights, ne)) {
    CVB_TRY(e->d_flags.ensure(4));
    CVB_TRY(e->d_sums.ensure(4));
    CVB_TRY(e->h_sums.ensure(64));
    CVB_TRY(e->ticket.ensure(4));
    CVB_TRY(e->partials.ensure(2 * 148 * 16 + 16));
    CVB_CUDA(cudaMemsetAsync(e->ticket.p, 0, 4 * sizeof(unsigned int), e->stream));
    return post_graph(e);
  }
  CVB_TRY(e->staging.ensure(static_cast<size_t>(ne) * 16 + 16));
  if (edges_aos) {
    CVB_CUDA(cudaMemcpyAsync(e->staging.p, edges_aos, static_cast<size_t>(ne) * 16, cudaMemcpyHostToDevice,
                             e->stream));
    return finish_graph(e, e->staging.p, nullptr, nullptr);
  }
  int64_t *dt = reinterpret_cast<int64_t *>(e->staging.p);
  double *dw = reinterpret_cast<double *>(e->staging.p + static_cast<size_t>(ne) * 8);
  if (ne > 0) {
    CVB_CUDA(cudaMemcpyAsync(dt, tails, static_cast<size_t>(ne) * 8, cudaMemcpyHostToDevice, e->stream));
    CVB_CUDA(cudaMemcpyAsync(dw, weights, static_cast<size_t>(ne) * 8, cudaMemcpyHostToDevice, e->stream));
  }
  return finish_graph(e, nullptr, dt, dw);
}

int cvb_graph_upload_device(cvb_engine *e, int64_t nv_local, int64_t base, int64_t nv_global,
                            const int64_t *d_offsets, const int64_t *d_tails, const double *d_weights) {
  CVB_CHECK(e && d_offsets, CVB_ERR_ARG, "null argument");
  CVB_TRY(check_sizes(nv_local, base, nv_global));
  int64_t first = 0, last = 0;
  CVB_CUDA(cudaMemcpyAsync(&first, d_offsets, sizeof(int64_t), cudaMemcpyDeviceToHost, e->stream));
  CVB_CUDA(cudaMemcpyAsync(&last, d_offsets + nv_local, sizeof(int64_t), cudaMemcpyDeviceToHost, e->stream));
  CVB_CUDA(cudaStreamSynchronize(e->stream));
  CVB_CHECK(first == 0 && last >= 0, CVB_ERR_ARG, "offsets must be rebased to 0");
  e->nvl = nv_local; e->base = base; e->nvg = nv_global; e->ne = last;
  e->graph_ready = false;
  CVB_TRY(e->offsets.ensure(nv_local + 2));
  CVB_CUDA(cudaMemcpyAsync(e->offsets.p, d_offsets, (nv_local + 1) * sizeof(int64_t), cudaMemcpyDeviceToDevice,
                           e->stream));
  return finish_graph(e, nullptr, d_tails, d_weights);
}

int cvb_graph_get_info(cvb_engine *e, cvb_graph_info *out) {
  CVB_CHECK(e && out && e->graph_ready, CVB_ERR_STATE, "no graph uploaded");
  memset(out, 0, sizeof *out);
  out->nv_local = e->nvl; out->base = e->base; out->nv_global = e->nvg; out->ne_local = e->ne;
  out->weight_mode = e->mode;
  out->num_classes = kNumClasses;
  for (int c = 0; c < kNumClasses; ++c) {
    out->class_count[c] = e->class_start[c + 1] - e->class_start[c];
    out->class_entries[c] = e->class_entries[c];
  }
  out->max_degree = e->max_degree;
  out->local_weight = e->local_weight;
  return CVB_OK;
}

int cvb_phase_begin(cvb_engine *e, double total_weight) {
  CVB_CHECK(e && e->graph_ready, CVB_ERR_STATE, "cvb_phase_begin: no graph uploaded");
  const double tot = total_weight > 0.0 ? total_weight : e->local_weight;
  e->constant = 1.0 / tot;  // louvain.cpp:2182
  e->past = 0; e->curr = 1; e->targ = 2;
  e->glabels_swapped = false;
  e->labels_only = false;
  cudaStream_t st = e->stream;
  if (e->nvl > 0) {
    k_phase_init_local<<<grid_for(e->nvl, 256), 256, 0, st>>>(e->nvl, static_cast<vid_t>(e->base), e->lab[0].p,
                                                             e->lab[1].p, e->lab[2].p, e->cw.p);
    e->launches++;
  }
  if (e->nvg > 0) {
    const bool single = e->nvl == e->nvg;
    k_phase_init_global<<<grid_for(e->nvg, 256), 256, 0, st>>>(e->nvg, single ? e->vdeg.p : nullptr,
                                                              single ? nullptr : e->labels_global.p, e->cdeg.p,
                                                              e->csize.p, e->ddeg.p, e->dsize.p);
    e->launches++;
    if (!single && e->nvl > 0) {
      CVB_CUDA(cudaMemcpyAsync(e->cdeg.p + e->base, e->vdeg.p, e->nvl * sizeof(double), cudaMemcpyDeviceToDevice,
                               st));
      CVB_CUDA(cudaMemcpyAsync(e->vdeg_global.p + e->base, e->vdeg.p, e->nvl * sizeof(double),
                               cudaMemcpyDeviceToDevice, st));
    }
  }
  CVB_CUDA(cudaGetLastError());
  e->phase_ready = true;
  e->sweep_timed = false;
  e->sweeps_in_phase = 0;
  return CVB_OK;
}

int cvb_phase_reset(cvb_engine *e) { return cvb_phase_begin(e, e ? 1.0 / e->constant : 0.0); }

int cvb_sweep(cvb_engine *e) {
  CVB_CHECK(e && e->phase_ready, CVB_ERR_STATE, "cvb_sweep: call cvb_phase_begin first");
  CVB_CUDA(cudaEventRecord(e->ev0, e->stream));
  if (e->mode == CVB_WEIGHTS_U32)
    CVB_TRY(launch_sweep_u32(e));
  else
    CVB_TRY(launch_sweep<double>(e));
  CVB_CUDA(cudaEventRecord(e->ev1, e->stream));
  e->sweep_timed = true;
  e->sweeps_in_phase++;
  return CVB_OK;
}

static int apply_reduce_async(cvb_engine *e) {
  const int64_t n = e->nvg > e->nvl ? e->nvg : e->nvl;
  const int grid = grid_for(n, kReduceBlock, e->num_sms * 4);
  k_apply_reduce<<<grid, kReduceBlock, 0, e->stream>>>(e->nvg, e->base, e->base + e->nvl, e->nvl, e->cdeg.p,
                                                      e->csize.p, e->ddeg.p, e->dsize.p, e->cw.p, e->partials.p,
                                                      e->ticket.p, e->d_sums.p);
  e->launches++;
  CVB_CUDA(cudaGetLastError());
  k_flag_to_f64<<<1, 1, 0, e->stream>>>(e->d_flags.p + 3, e->d_sums.p + 2);  // big-path bucket overflow (never expected)
  CVB_CUDA(cudaMemcpyAsync(e->h_sums.p, e->d_sums.p, 3 * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
  return CVB_OK;
}

// labels-only mode: Cinfo was rebuilt from labels; only zero the local deltas and reduce
static int apply_reduce_zero_deltas(cvb_engine *e) {
  cudaStream_t st = e->stream;
  if (!e->labels_only) {  // first use in this phase: the first sweep still wrote deltas
    CVB_CUDA(cudaMemsetAsync(e->ddeg.p, 0, e->nvg * sizeof(double), st));
    CVB_CUDA(cudaMemsetAsync(e->dsize.p, 0, e->nvg * sizeof(int32_t), st));
    e->labels_only = true;
  }
  const int grid = grid_for(e->nvl, kReduceBlock, e->num_sms * 4);
  k_reduce_only<<<grid, kReduceBlock, 0, st>>>(e->base, e->base + e->nvl, e->nvl, e->cdeg.p, e->cw.p, e->partials.p,
                                               e->ticket.p, e->d_sums.p);
  e->launches++;
  CVB_CUDA(cudaGetLastError());
  CVB_CUDA(cudaMemcpyAsync(e->h_sums.p, e->d_sums.p, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
  return CVB_OK;
}

int cvb_apply_and_reduce(cvb_engine *e, double sums[2]) {
  CVB_CHECK(e && e->phase_ready && sums, CVB_ERR_STATE, "cvb_apply_and_reduce: no phase");
  CVB_TRY(apply_reduce_async(e));
  CVB_CUDA(cudaStreamSynchronize(e->stream));
  memcpy(sums, e->h_sums.p, 2 * sizeof(double));
  CVB_CHECK(reinterpret_cast<const double *>(e->h_sums.p)[2] == 0.0, CVB_ERR_CAPACITY,
            "a partition bucket of the big-vertex path overflowed (community ids colliding in the partition "
            "hash); rerun with CVB_BIG=0");
  return CVB_OK;
}

double cvb_modularity_from_sums(const cvb_engine *e, const double sums[2]) {
  const double c = e->constant;
  return (sums[0] * c) - (sums[1] * c * c);  // louvain.cpp:2472-2473
}

int cvb_rebuild_and_reduce_async(cvb_engine *e) {
  CVB_CHECK(e && e->phase_ready, CVB_ERR_STATE, "cvb_rebuild_and_reduce: no phase");
  CVB_CHECK(e->nvl != e->nvg, CVB_ERR_STATE, "cvb_rebuild_and_reduce is the N > 1 path");
  cudaStream_t st = e->stream;
  k_apply_label_diff<<<grid_for(e->nvg, 256), 256, 0, st>>>(e->nvg, e->glabels_curr(), e->glabels_next(),
                                                            e->vdeg_global.p, e->cdeg.p, e->csize.p);
  e->launches++;
  return apply_reduce_zero_deltas(e);
}

int cvb_rebuild_and_reduce(cvb_engine *e, double sums[2]) {
  CVB_CHECK(e && e->phase_ready && sums, CVB_ERR_STATE, "cvb_rebuild_and_reduce: no phase");
  CVB_CHECK(e->nvl != e->nvg, CVB_ERR_STATE, "cvb_rebuild_and_reduce is the N > 1 path");
  cudaStream_t st = e->stream;
  k_apply_label_diff<<<grid_for(e->nvg, 256), 256, 0, st>>>(e->nvg, e->glabels_curr(), e->glabels_next(),
                                                            e->vdeg_global.p, e->cdeg.p, e->csize.p);
  e->launches++;
  // the sweep's own delta arrays are not used in this mode (kernels skip them once labels_only is set)
  CVB_TRY(apply_reduce_zero_deltas(e));
  CVB_CUDA(cudaStreamSynchronize(st));
  memcpy(sums, e->h_sums.p, 2 * sizeof(double));
  return CVB_OK;
}

int cvb_rotate_labels(cvb_engine *e) {
  CVB_CHECK(e && e->phase_ready, CVB_ERR_STATE, "no phase");
  if (e->nvl != e->nvg && e->labels_only) e->glabels_swapped = !e->glabels_swapped;
  const int oldpast = e->past;  // louvain.cpp:562-567
  e->past = e->curr;
  e->curr = e->targ;
  e->targ = oldpast;
  return CVB_OK;
}

int cvb_phase_run(cvb_engine *e, double lower, double thresh, int max_iters, double *q_out, int *iters_out) {
  CVB_CHECK(e && e->graph_ready, CVB_ERR_STATE, "cvb_phase_run: no graph uploaded");
  CVB_CHECK(e->nvl == e->nvg, CVB_ERR_STATE, "cvb_phase_run is the one-GPU driver; use the stepwise calls");
  CVB_TRY(cvb_phase_begin(e, 0.0));
  double prev = lower, cur = -1.0;
  int iters = 0;
  while (true) {  // louvain.cpp:638-739
    iters++;
    CVB_TRY(cvb_sweep(e));
    double sums[2];
    CVB_TRY(cvb_apply_and_reduce(e, sums));
    cur = cvb_modularity_from_sums(e, sums);
    if ((cur - prev) < thresh) break;
    prev = cur;
    if (prev < lower) prev = lower;
    CVB_TRY(cvb_rotate_labels(e));  // also on the max_iters exit: the returned Q belongs to the labels in LABELS_PAST
    if (max_iters > 0 && iters >= max_iters) break;
  }
  if (q_out) *q_out = prev;
  if (iters_out) *iters_out = iters;
  return CVB_OK;
}

static int label_index(const cvb_engine *e, int which) {
  return which == CVB_LABELS_PAST ? e->past : which == CVB_LABELS_CURR ? e->curr : e->targ;
}

int cvb_get_labels(cvb_engine *e, int which, int64_t *out) {
  CVB_CHECK(e && e->phase_ready && out && which >= 0 && which <= 2, CVB_ERR_STATE, "cvb_get_labels: no phase");
  if (e->nvl == 0) return CVB_OK;
  CVB_TRY(e->staging2.ensure(static_cast<size_t>(e->nvl) * 8));
  int64_t *d64 = reinterpret_cast<int64_t *>(e->staging2.p);
  k_widen<vid_t, int64_t><<<grid_for(e->nvl, 256), 256, 0, e->stream>>>(e->lab[label_index(e, which)].p, d64, e->nvl);
  e->launches++;
  CVB_CUDA(cudaMemcpyAsync(out, d64, e->nvl * sizeof(int64_t), cudaMemcpyDeviceToHost, e->stream));
  CVB_CUDA(cudaStreamSynchronize(e->stream));
  return CVB_OK;
}

int cvb_get_cluster_weight(cvb_engine *e, double *out) {
  CVB_CHECK(e && e->phase_ready && out, CVB_ERR_STATE, "no phase");
  CVB_CUDA(cudaMemcpyAsync(out, e->cw.p, e->nvl * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
  CVB_CUDA(cudaStreamSynchronize(e->stream));
  return CVB_OK;
}

int cvb_get_colors(cvb_engine *e, int32_t *out) {
  CVB_CHECK(e && e->graph_ready && out && e->col_color.p, CVB_ERR_STATE, "cvb_get_colors: no colouring was run");
  CVB_CUDA(cudaMemcpyAsync(out, e->col_color.p, e->nvl * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
  CVB_CUDA(cudaStreamSynchronize(e->stream));
  return CVB_OK;
}

int cvb_get_cinfo(cvb_engine *e, int64_t *size_out, double *degree_out) {
  CVB_CHECK(e && e->phase_ready, CVB_ERR_STATE, "no phase");
  if (degree_out)
    CVB_CUDA(cudaMemcpyAsync(degree_out, e->cdeg.p, e->nvg * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
  if (size_out && e->nvg > 0) {
    CVB_TRY(e->staging2.ensure(static_cast<size_t>(e->nvg) * 8));
    int64_t *d64 = reinterpret_cast<int64_t *>(e->staging2.p);
    k_widen<int32_t, int64_t><<<grid_for(e->nvg, 256), 256, 0, e->stream>>>(e->csize.p, d64, e->nvg);
    e->launches++;
    CVB_CUDA(cudaMemcpyAsync(size_out, d64, e->nvg * sizeof(int64_t), cudaMemcpyDeviceToHost, e->stream));
  }
  CVB_CUDA(cudaStreamSynchronize(e->stream));
  return CVB_OK;
}

void *cvb_buffer(cvb_engine *e, int which) {
  if (!e) return nullptr;
  switch (which) {
    case CVB_BUF_LABELS_GLOBAL: return const_cast<vid_t *>(e->sweep_labels());
    case CVB_BUF_LABELS_GLOBAL_NEXT: return e->nvl == e->nvg ? e->lab[e->targ].p : e->glabels_next();
    case CVB_BUF_VDEG_GLOBAL: return e->nvl == e->nvg ? e->vdeg.p : e->vdeg_global.p;
    case CVB_BUF_SUMS: return e->d_sums.p;
    case CVB_BUF_LABELS_CURR: return e->lab[e->curr].p;
    case CVB_BUF_LABELS_TARGET: return e->lab[e->targ].p;
    case CVB_BUF_LABELS_PAST: return e->lab[e->past].p;
    case CVB_BUF_CDEG: return e->cdeg.p;
    case CVB_BUF_CSIZE: return e->csize.p;
    case CVB_BUF_DDEG: return e->ddeg.p;
    case CVB_BUF_DSIZE: return e->dsize.p;
    case CVB_BUF_VDEG: return e->vdeg.p;
    case CVB_BUF_CW: return e->cw.p;
    case CVB_BUF_OFFSETS: return e->offsets.p;
    case CVB_BUF_TAILS: return e->tails.p;
    case CVB_BUF_WEIGHTS: return e->weights.p;
    case CVB_BUF_PAIR_KEYS: return e->pair_keys.p;
    case CVB_BUF_PAIR_WEIGHTS: return e->pair_w.p;
    default: return nullptr;
  }
}

void *cvb_stream(cvb_engine *e) { return e ? e->stream : nullptr; }

int cvb_sync(cvb_engine *e) {
  CVB_CHECK(e, CVB_ERR_ARG, "null engine");
  CVB_CUDA(cudaStreamSynchronize(e->stream));
  return CVB_OK;
}

int64_t cvb_launch_count(const cvb_engine *e) { return e ? e->launches : 0; }

float cvb_last_sweep_ms(cvb_engine *e) {
  if (!e || !e->sweep_timed) return -1.f;
  if (cudaEventSynchronize(e->ev1) != cudaSuccess) return -1.f;
  float ms = -1.f;
  if (cudaEventElapsedTime(&ms, e->ev0, e->ev1) != cudaSuccess) return -1.f;
  return ms;
}

// ---- per-iteration boundary call (gpu_for_louvain_iteration flattened) ------
int cvb_louvain_iteration(cvb_engine *e, int64_t nv, int64_t base, int64_t nv_global, const int64_t *offsets,
                          const void *edges_aos, const int64_t *curr_comm, const double *v_degree,
                          const int64_t *local_csize, const double *local_cdeg, int64_t n_ghost,
                          const int64_t *ghost_vertex, const int64_t *ghost_comm, int64_t n_rcomm,
                          const int64_t *rcomm_id, const int64_t *rcomm_size, const double *rcomm_degree,
                          double constant, int num_iters, int64_t *target_comm, double *cluster_weight,
                          int64_t *local_cupd_size, double *local_cupd_deg, int64_t *rcomm_upd_size,
                          double *rcomm_upd_deg) {
  if (!e) { set_error("null engine"); return CVB_ERR_ARG; }
  if (nv <= 0) return 1;  // louvain_cuda.cu:1685
  cudaStream_t st = e->stream;
  if (num_iters == 1) {
    CVB_TRY(cvb_graph_upload_host(e, nv, base, nv_global, offsets, edges_aos, nullptr, nullptr));
    CVB_TRY(cvb_phase_begin(e, 1.0));
    if (v_degree) CVB_CUDA(cudaMemcpyAsync(e->vdeg.p, v_degree, nv * sizeof(double), cudaMemcpyHostToDevice, st));
  }
  CVB_CHECK(e->phase_ready && e->nvl == nv && e->base == base && e->nvg == nv_global, CVB_ERR_STATE,
            "cvb_louvain_iteration: graph changed without num_iters == 1");
  e->constant = constant;
  // staging layout (device): [curr 8nv][csize 8nv][cdeg 8nv][ghost 16ng][rcomm 24nr]
  const size_t need = static_cast<size_t>(nv) * 24 + static_cast<size_t>(n_ghost) * 16 +
                      static_cast<size_t>(n_rcomm) * 24 + 64;
  CVB_TRY(e->staging.ensure(need));
  unsigned char *sp = e->staging.p;
  int64_t *d_curr = reinterpret_cast<int64_t *>(sp); sp += nv * 8;
  int64_t *d_csz = reinterpret_cast<int64_t *>(sp); sp += nv * 8;
  double *d_cdg = reinterpret_cast<double *>(sp); sp += nv * 8;
  int64_t *d_gv = reinterpret_cast<int64_t *>(sp); sp += n_ghost * 8;
  int64_t *d_gc = reinterpret_cast<int64_t *>(sp); sp += n_ghost * 8;
  int64_t *d_rid = reinterpret_cast<int64_t *>(sp); sp += n_rcomm * 8;
  int64_t *d_rsz = reinterpret_cast<int64_t *>(sp); sp += n_rcomm * 8;
  double *d_rdg = reinterpret_cast<double *>(sp);
  CVB_CUDA(cudaMemcpyAsync(d_curr, curr_comm, nv * 8, cudaMemcpyHostToDevice, st));
  CVB_CUDA(cudaMemcpyAsync(d_csz, local_csize, nv * 8, cudaMemcpyHostToDevice, st));
  CVB_CUDA(cudaMemcpyAsync(d_cdg, local_cdeg, nv * 8, cudaMemcpyHostToDevice, st));
  vid_t *lab_local = e->lab[e->curr].p;
  k_narrow<int64_t, vid_t><<<grid_for(nv, 256), 256, 0, st>>>(d_curr, lab_local, nv, 0);
  k_scatter_cinfo<<<grid_for(nv, 256), 256, 0, st>>>(nv, nullptr, d_csz, d_cdg, base, e->csize.p, e->cdeg.p);
  e->launches += 2;
  if (e->nvl != e->nvg) {
    CVB_CUDA(cudaMemcpyAsync(e->labels_global.p + base, lab_local, nv * sizeof(vid_t), cudaMemcpyDeviceToDevice, st));
    if (n_ghost > 0) {
      CVB_CUDA(cudaMemcpyAsync(d_gv, ghost_vertex, n_ghost * 8, cudaMemcpyHostToDevice, st));
      CVB_CUDA(cudaMemcpyAsync(d_gc, ghost_comm, n_ghost * 8, cudaMemcpyHostToDevice, st));
      k_scatter_labels<<<grid_for(n_ghost, 256), 256, 0, st>>>(n_ghost, d_gv, d_gc, e->labels_global.p);
      e->launches++;
    }
    if (n_rcomm > 0) {
      CVB_CUDA(cudaMemcpyAsync(d_rid, rcomm_id, n_rcomm * 8, cudaMemcpyHostToDevice, st));
      CVB_CUDA(cudaMemcpyAsync(d_rsz, rcomm_size, n_rcomm * 8, cudaMemcpyHostToDevice, st));
      CVB_CUDA(cudaMemcpyAsync(d_rdg, rcomm_degree, n_rcomm * 8, cudaMemcpyHostToDevice, st));
      k_scatter_cinfo<<<grid_for(n_rcomm, 256), 256, 0, st>>>(n_rcomm, d_rid, d_rsz, d_rdg, 0, e->csize.p,
                                                             e->cdeg.p);
      e->launches++;
    }
  }
  e->sweeps_in_phase = 1;  // caller-provided labels: never the singleton fast path
  CVB_TRY(cvb_sweep(e));
  // outputs: [target 8nv][dsize 8nv][ddeg 8nv][rsize 8nr][rdeg 8nr]
  const size_t need2 = static_cast<size_t>(nv) * 24 + static_cast<size_t>(n_rcomm) * 16 + 64;
  CVB_TRY(e->staging2.ensure(need2));
  CVB_TRY(e->h_stage.ensure(need2));
  unsigned char *op = e->staging2.p;
  int64_t *o_tgt = reinterpret_cast<int64_t *>(op); op += nv * 8;
  int64_t *o_dsz = reinterpret_cast<int64_t *>(op); op += nv * 8;
  double *o_ddg = reinterpret_cast<double *>(op); op += nv * 8;
  int64_t *o_rsz = reinterpret_cast<int64_t *>(op); op += n_rcomm * 8;
  double *o_rdg = reinterpret_cast<double *>(op);
  k_widen<vid_t, int64_t><<<grid_for(nv, 256), 256, 0, st>>>(e->lab[e->targ].p, o_tgt, nv);
  k_gather_deltas<<<grid_for(nv, 256), 256, 0, st>>>(nv, nullptr, base, e->ddeg.p, e->dsize.p, o_dsz, o_ddg);
  e->launches += 2;
  if (n_rcomm > 0 && e->nvl != e->nvg) {
    k_gather_deltas<<<grid_for(n_rcomm, 256), 256, 0, st>>>(n_rcomm, d_rid, 0, e->ddeg.p, e->dsize.p, o_rsz, o_rdg);
    e->launches++;
  }
  CVB_CUDA(cudaGetLastError());
  CVB_CUDA(cudaMemcpyAsync(target_comm, o_tgt, nv * 8, cudaMemcpyDeviceToHost, st));
  CVB_CUDA(cudaMemcpyAsync(cluster_weight, e->cw.p, nv * 8, cudaMemcpyDeviceToHost, st));
  unsigned char *hp = e->h_stage.p;
  CVB_CUDA(cudaMemcpyAsync(hp, o_dsz, static_cast<size_t>(nv) * 16 + static_cast<size_t>(n_rcomm) * 16,
                           cudaMemcpyDeviceToHost, st));
  CVB_CUDA(cudaStreamSynchronize(st));
  const int64_t *h_dsz = reinterpret_cast<const int64_t *>(hp);
  const double *h_ddg = reinterpret_cast<const double *>(hp + nv * 8);
  host_parallel_for(nv, [=](int64_t lo, int64_t hi) {
    for (int64_t i = lo; i < hi; ++i) {
      local_cupd_size[i] += h_dsz[i];
      local_cupd_deg[i] += h_ddg[i];
    }
  });
  if (n_rcomm > 0 && e->nvl != e->nvg) {
    const int64_t *h_rsz = reinterpret_cast<const int64_t *>(hp + nv * 16);
    const double *h_rdg = reinterpret_cast<const double *>(hp + nv * 16 + n_rcomm * 8);
    for (int64_t i = 0; i < n_rcomm; ++i) {
      rcomm_upd_size[i] += h_rsz[i];
      rcomm_upd_deg[i] += h_rdg[i];
    }
  }
  // this call's labels become the engine's "current"; the caller rotates its own vectors
  return 1;
}

int cvb_reduce_mod(cvb_engine *e, int64_t nv, const double *local_cdeg, const double *cluster_weight,
                   double le_la_xx[2]) {
  CVB_CHECK(e && le_la_xx, CVB_ERR_ARG, "null argument");
  le_la_xx[0] = le_la_xx[1] = 0.0;
  if (nv <= 0) return CVB_OK;
  cudaStream_t st = e->stream;
  CVB_TRY(e->staging.ensure(static_cast<size_t>(nv) * 16 + 64));
  CVB_TRY(e->d_sums.ensure(4));
  CVB_TRY(e->h_sums.ensure(64));
  CVB_TRY(e->ticket.ensure(4));
  CVB_TRY(e->partials.ensure(2 * 148 * 16 + 16));
  double *d_cw = reinterpret_cast<double *>(e->staging.p);
  double *d_cd = d_cw + nv;
  CVB_CUDA(cudaMemcpyAsync(d_cw, cluster_weight, nv * 8, cudaMemcpyHostToDevice, st));
  CVB_CUDA(cudaMemcpyAsync(d_cd, local_cdeg, nv * 8, cudaMemcpyHostToDevice, st));
  CVB_TRY(e->staging2.ensure(static_cast<size_t>(nv) * 12 + 64));  // zero deltas for the fused kernel
  double *z_dd = reinterpret_cast<double *>(e->staging2.p);
  int32_t *z_ds = reinterpret_cast<int32_t *>(e->staging2.p + nv * 8);
  CVB_CUDA(cudaMemsetAsync(e->staging2.p, 0, static_cast<size_t>(nv) * 12, st));
  CVB_CUDA(cudaMemsetAsync(e->ticket.p, 0, sizeof(unsigned int), st));
  // a throwaway size array is not needed: deltas are zero so csize is never touched
  k_apply_reduce<<<grid_for(nv, kReduceBlock, e->num_sms * 4), kReduceBlock, 0, st>>>(
      nv, 0, nv, nv, d_cd, nullptr, z_dd, z_ds, d_cw, e->partials.p, e->ticket.p, e->d_sums.p);
  e->launches++;
  CVB_CUDA(cudaGetLastError());
  CVB_CUDA(cudaMemcpyAsync(e->h_sums.p, e->d_sums.p, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
  CVB_CUDA(cudaStreamSynchronize(st));
  memcpy(le_la_xx, e->h_sums.p, 2 * sizeof(double));
  return CVB_OK;
}

// ---- coarsening + multi-phase driver -------------------------------------------
// ---- coarsening in three stages (one rank runs them back to back; N ranks exchange the
// ---- contracted pairs between stage 2 and stage 3, cuvite_b200/dist.py) --------------

// Stage 1 — distReNumber (rebuild.cpp:72-85): vnew[i] = rank, in order of first appearance
// scanning i ascending, of labels[i]'s community.  labels are ids in [0, n).
static int renumber_impl(cvb_engine *e, const vid_t *labels, int64_t n, int32_t *vnew, int64_t *nnv_out) {
  cudaStream_t st = e->stream;
  DevBuf<int32_t> &first = e->cs.first, &flag = e->cs.flag, &rank = e->cs.rank;
  DevBuf<unsigned char> &tmp = e->cs.tmp;
  auto release = [&]() { if (!e->keep_scratch) { first.release(); flag.release(); rank.release(); tmp.release(); } };
  struct Guard { decltype(release) &f; ~Guard() { f(); } } guard{release};
  *nnv_out = 0;
  if (n == 0) return CVB_OK;
  CVB_TRY(first.ensure(n + 1)); CVB_TRY(flag.ensure(n + 1)); CVB_TRY(rank.ensure(n + 2));
  k_fill_i32<<<grid_for(n, 256), 256, 0, st>>>(first.p, n, 0x7FFFFFFF);
  k_first_appearance<<<grid_for(n, 256), 256, 0, st>>>(labels, n, first.p);
  k_flag_first<<<grid_for(n, 256), 256, 0, st>>>(labels, first.p, n, flag.p);
  size_t tb = 0;
  CVB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, flag.p, rank.p, static_cast<int>(n), st));
  CVB_TRY(tmp.ensure(tb + 1024));
  size_t tcap = tmp.cap;
  CVB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tcap, flag.p, rank.p, static_cast<int>(n), st));
  k_vertex_newid<<<grid_for(n, 256), 256, 0, st>>>(labels, first.p, rank.p, n, vnew);
  e->launches += 5;
  int32_t last_rank = 0, last_flag = 0;
  CVB_CUDA(cudaMemcpyAsync(&last_rank, rank.p + (n - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  CVB_CUDA(cudaMemcpyAsync(&last_flag, flag.p + (n - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  CVB_CUDA(cudaStreamSynchronize(st));
  *nnv_out = static_cast<int64_t>(last_rank) + last_flag;
  return CVB_OK;
}

// Sort (key, weight) pairs by key (stable) and merge equal keys: ukeys / uw get the distinct
// keys ascending and their summed weights (F64: summed sequentially in sorted order — the
// reference's order on one rank; U32: exact in any order).
static int sort_reduce_pairs(cvb_engine *e, const uint64_t *k_in, const void *w_in, int64_t m, int bits,
                             DevBuf<uint64_t> &ukeys, DevBuf<unsigned char> &uw, int64_t *mu_out) {
  cudaStream_t st = e->stream;
  const bool u32 = e->mode == CVB_WEIGHTS_U32;
  const size_t wbytes = u32 ? 4 : 8;
  *mu_out = 0;
  if (m == 0) return CVB_OK;
  CVB_CHECK(m < 2147483000LL, CVB_ERR_CAPACITY, "coarsening: more than 2^31 entries on one GPU");
  DevBuf<uint64_t> &k_out = e->cs.k_out;
  DevBuf<unsigned char> &w_out = e->cs.w_out, &tmp = e->cs.tmp;
  DevBuf<int32_t> &head = e->cs.head, &seg = e->cs.seg;
  DevBuf<int64_t> &seg_start = e->cs.seg_start;
  auto release = [&]() {
    if (!e->keep_scratch) { k_out.release(); w_out.release(); tmp.release(); head.release(); seg.release(); seg_start.release(); }
  };
  struct Guard { decltype(release) &f; ~Guard() { f(); } } guard{release};
  size_t tb2 = 0, tb4 = 0;
  if (u32)
    CVB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb2, (uint64_t *)nullptr, (uint64_t *)nullptr,
                                             (uint32_t *)nullptr, (uint32_t *)nullptr, static_cast<int>(m), 0, 64, st));
  else
    CVB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb2, (uint64_t *)nullptr, (uint64_t *)nullptr,
                                             (double *)nullptr, (double *)nullptr, static_cast<int>(m), 0, 64, st));
  CVB_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tb4, (int32_t *)nullptr, (int32_t *)nullptr, static_cast<int>(m), st));
  CVB_TRY(tmp.ensure((tb2 > tb4 ? tb2 : tb4) + 1024));
  CVB_TRY(k_out.ensure(m)); CVB_TRY(w_out.ensure(m * wbytes)); CVB_TRY(head.ensure(m)); CVB_TRY(seg.ensure(m));
  size_t tcap = tmp.cap;
  if (u32)
    CVB_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tcap, k_in, k_out.p, reinterpret_cast<const uint32_t *>(w_in),
                                             reinterpret_cast<uint32_t *>(w_out.p), static_cast<int>(m), 0, bits, st));
  else
    CVB_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tcap, k_in, k_out.p, reinterpret_cast<const double *>(w_in),
                                             reinterpret_cast<double *>(w_out.p), static_cast<int>(m), 0, bits, st));
  k_heads<<<grid_for(m, 256), 256, 0, st>>>(k_out.p, m, head.p);
  tcap = tmp.cap;
  CVB_CUDA(cub::DeviceScan::InclusiveSum(tmp.p, tcap, head.p, seg.p, static_cast<int>(m), st));
  e->launches += 3;
  int32_t nseg32 = 0;
  CVB_CUDA(cudaMemcpyAsync(&nseg32, seg.p + (m - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  CVB_CUDA(cudaStreamSynchronize(st));
  const int64_t mu = nseg32;
  CVB_TRY(ukeys.ensure(mu + 1)); CVB_TRY(uw.ensure((mu + 1) * wbytes)); CVB_TRY(seg_start.ensure(mu + 1));
  k_emit_pairs<<<grid_for(m, 256), 256, 0, st>>>(k_out.p, head.p, seg.p, m, ukeys.p, seg_start.p);
  if (u32) {
    CVB_CUDA(cudaMemsetAsync(uw.p, 0, static_cast<size_t>(mu) * sizeof(uint32_t), st));
    k_sum_segments_u32<<<static_cast<int>((m + 255) / 256), 256, 0, st>>>(
        reinterpret_cast<uint32_t *>(w_out.p), seg.p, m, reinterpret_cast<uint32_t *>(uw.p));
  } else
    k_sum_segments_f64<<<grid_for(mu, 256), 256, 0, st>>>(reinterpret_cast<double *>(w_out.p), seg_start.p, mu, m,
                                                         reinterpret_cast<double *>(uw.p));
  e->launches += 2;
  CVB_CUDA(cudaGetLastError());
  *mu_out = mu;
  return CVB_OK;
}

// Stage 2 — fill_newEdgesMap (rebuild.cpp:244-279) for the rows this engine holds: entry
// (i -> t, w) becomes (vnew[base+i] -> vnew[t], w); parallel entries are merged.  The distinct
// pairs (key = src << 32 | dst, ascending) land in e->pair_keys / e->pair_w.
static int contract_impl(cvb_engine *e, const int32_t *vnew_global, int64_t nnv, int64_t *m_out) {
  cudaStream_t st = e->stream;
  const int64_t nv = e->nvl, ne = e->ne;
  *m_out = 0;
  e->pair_count = 0;
  if (ne == 0) return CVB_OK;
  DevBuf<uint64_t> &k_in = e->cs.k_in;
  struct Guard { cvb_engine *e; ~Guard() { if (!e->keep_scratch) e->cs.k_in.release(); } } guard{e};
  CVB_TRY(k_in.ensure(ne));
  k_edge_keys<<<static_cast<int>((nv * 32 + 255) / 256), 256, 0, st>>>(e->offsets.p, e->tails.p, vnew_global, e->base, nv,
                                                                      k_in.p);
  e->launches++;
  int bits = 33;  // high word holds src < nnv
  while ((int64_t(1) << (bits - 32)) < nnv && bits < 64) ++bits;
  CVB_TRY(sort_reduce_pairs(e, k_in.p, e->weights.p, ne, bits, e->pair_keys, e->pair_w, m_out));
  e->pair_count = *m_out;
  return CVB_OK;
}

// Stage 3 — send_newEdges (rebuild.cpp:405-427): the pairs whose src lies in
// [base, base + nv_local) become the engine's graph, adjacency sorted by tail.  `unique`
// says the pairs are already distinct and ascending (one rank); otherwise (pairs received
// from several ranks) they are sorted and merged first.
static int install_pairs_impl(cvb_engine *e, int64_t nv_local, int64_t base, int64_t nv_global, const uint64_t *keys,
                              const void *w, int64_t m, bool unique) {
  cudaStream_t st = e->stream;
  const size_t wbytes = e->mode == CVB_WEIGHTS_U32 ? 4 : 8;
  CVB_TRY(check_sizes(nv_local, base, nv_global));
  DevBuf<uint64_t> &ukeys = e->cs.ukeys;
  DevBuf<unsigned char> &uw = e->cs.uw, &tmp = e->cs.tmp;
  DevBuf<unsigned long long> &row_count = e->cs.row_count;
  DevBuf<int64_t> &noff = e->cs.noff;
  auto release = [&]() {
    if (!e->keep_scratch) { ukeys.release(); uw.release(); tmp.release(); row_count.release(); noff.release(); }
  };
  struct Guard { decltype(release) &f; ~Guard() { f(); } } guard{release};
  int64_t mu = m;
  if (!unique && m > 0) {
    int bits = 33;
    while ((int64_t(1) << (bits - 32)) < nv_global && bits < 64) ++bits;
    CVB_TRY(sort_reduce_pairs(e, keys, w, m, bits, ukeys, uw, &mu));
    keys = ukeys.p;
    w = uw.p;
  }
  CVB_TRY(noff.ensure(nv_local + 2));
  CVB_TRY(row_count.ensure(nv_local + 2));
  CVB_CUDA(cudaMemsetAsync(row_count.p, 0, (nv_local + 2) * sizeof(unsigned long long), st));
  CVB_TRY(e->tails.ensure(mu + 1));
  CVB_TRY(e->d_flags.ensure(4));
  CVB_CUDA(cudaMemsetAsync(e->d_flags.p + 2, 0, sizeof(int), st));
  if (mu > 0) {
    k_pairs_to_rows<<<static_cast<int>((mu + 255) / 256), 256, 0, st>>>(keys, mu, base, nv_local, e->tails.p, row_count.p,
                                                                    e->d_flags.p + 2);
    e->launches++;
  }
  size_t tb5 = 0;
  CVB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb5, (unsigned long long *)nullptr, (int64_t *)nullptr,
                                         static_cast<int>(nv_local + 1), st));
  CVB_TRY(tmp.ensure(tb5 + 1024));
  size_t tcap = tmp.cap;
  CVB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tcap, row_count.p, noff.p, static_cast<int>(nv_local + 1), st));
  e->launches++;
  e->nvl = nv_local;
  e->nvg = nv_global;
  e->base = base;
  e->ne = mu;
  e->graph_ready = false;
  CVB_TRY(e->offsets.ensure(nv_local + 2));
  CVB_CUDA(cudaMemcpyAsync(e->offsets.p, noff.p, (nv_local + 1) * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
  if (w != e->weights.p) {  // (one rank: pair_w is a different buffer from the old graph's weights)
    CVB_TRY(e->weights.ensure((mu + 1) * wbytes));
    if (mu > 0) CVB_CUDA(cudaMemcpyAsync(e->weights.p, w, mu * wbytes, cudaMemcpyDeviceToDevice, st));
  }
  int bad = 0;
  CVB_CUDA(cudaMemcpyAsync(&bad, e->d_flags.p + 2, sizeof(int), cudaMemcpyDeviceToHost, st));
  CVB_CUDA(cudaStreamSynchronize(st));
  CVB_CHECK(bad == 0, CVB_ERR_ARG, "cvb_graph_from_pairs: a pair's source lies outside [base, base + nv_local)");
  CVB_CUDA(cudaGetLastError());
  return post_graph(e);
}

static int coarsen_impl(cvb_engine *e, int64_t *new_nv_out, int32_t *d_compose, int64_t n_compose) {
  CVB_CHECK(e && e->phase_ready, CVB_ERR_STATE, "cvb_coarsen: run a phase first");
  CVB_CHECK(e->nvl == e->nvg, CVB_ERR_STATE, "cvb_coarsen is the one-GPU path (N ranks: cvb_renumber + "
            "cvb_contract_local + cvb_graph_from_pairs)");
  cudaStream_t st = e->stream;
  const int64_t nv = e->nvl;
  DevBuf<int32_t> &vnew = e->cs.vnew;
  struct Guard { cvb_engine *e; ~Guard() { if (!e->keep_scratch) e->cs.vnew.release(); } } guard{e};
  CVB_TRY(vnew.ensure(nv + 1));
  int64_t nnv = 0, m = 0;
  CVB_TRY(renumber_impl(e, e->lab[e->past].p, nv, vnew.p, &nnv));  // cvect = pastComm, louvain.cpp:741
  if (d_compose) {
    k_compose<<<grid_for(n_compose, 256), 256, 0, st>>>(d_compose, vnew.p, n_compose);
    e->launches++;
  }
  CVB_TRY(contract_impl(e, vnew.p, nnv, &m));
  CVB_TRY(install_pairs_impl(e, nnv, 0, nnv, e->pair_keys.p, e->pair_w.p, m, true));
  if (new_nv_out) *new_nv_out = nnv;
  return CVB_OK;
}

int cvb_renumber(cvb_engine *e, const int32_t *d_labels, int64_t n, int32_t *d_newid, int64_t *n_comm_out) {
  CVB_CHECK(e && d_labels && d_newid && n_comm_out && n >= 0, CVB_ERR_ARG, "cvb_renumber: bad argument");
  return renumber_impl(e, d_labels, n, d_newid, n_comm_out);
}

int cvb_contract_local(cvb_engine *e, const int32_t *d_newid_global, int64_t n_comm, int64_t *n_pairs_out) {
  CVB_CHECK(e && e->graph_ready && d_newid_global && n_pairs_out, CVB_ERR_ARG, "cvb_contract_local: bad argument");
  return contract_impl(e, d_newid_global, n_comm, n_pairs_out);
}

int cvb_graph_from_pairs(cvb_engine *e, int64_t nv_local, int64_t base, int64_t nv_global, const uint64_t *d_keys,
                         const void *d_weights, int64_t n_pairs) {
  CVB_CHECK(e && n_pairs >= 0 && (n_pairs == 0 || (d_keys && d_weights)), CVB_ERR_ARG, "cvb_graph_from_pairs: bad argument");
  CVB_TRY(e->d_flags.ensure(4));
  return install_pairs_impl(e, nv_local, base, nv_global, d_keys, d_weights, n_pairs, false);
}

int cvb_coarsen(cvb_engine *e, int64_t *new_nv_out) { return coarsen_impl(e, new_nv_out, nullptr, 0); }

int cvb_graph_download(cvb_engine *e, int64_t *offsets, int64_t *tails, double *weights) {
  CVB_CHECK(e && e->graph_ready, CVB_ERR_STATE, "no graph");
  cudaStream_t st = e->stream;
  if (offsets) CVB_CUDA(cudaMemcpyAsync(offsets, e->offsets.p, (e->nvl + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  if (e->ne > 0 && (tails || weights)) {
    CVB_TRY(e->staging2.ensure(static_cast<size_t>(e->ne) * 8));
    if (tails) {
      int64_t *d64 = reinterpret_cast<int64_t *>(e->staging2.p);
      k_widen<vid_t, int64_t><<<grid_for(e->ne, 256), 256, 0, st>>>(e->tails.p, d64, e->ne);
      CVB_CUDA(cudaMemcpyAsync(tails, d64, e->ne * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
      CVB_CUDA(cudaStreamSynchronize(st));
      e->launches++;
    }
    if (weights) {
      if (e->mode == CVB_WEIGHTS_U32) {
        double *d64 = reinterpret_cast<double *>(e->staging2.p);
        k_widen<uint32_t, double><<<grid_for(e->ne, 256), 256, 0, st>>>(reinterpret_cast<uint32_t *>(e->weights.p), d64, e->ne);
        e->launches++;
        CVB_CUDA(cudaMemcpyAsync(weights, d64, e->ne * sizeof(double), cudaMemcpyDeviceToHost, st));
      } else {
        CVB_CUDA(cudaMemcpyAsync(weights, e->weights.p, e->ne * sizeof(double), cudaMemcpyDeviceToHost, st));
      }
    }
  }
  CVB_CUDA(cudaStreamSynchronize(st));
  return CVB_OK;
}

int cvb_louvain_run(cvb_engine *e, int max_levels, double *level_q, int *level_iters, int64_t *level_nv,
                    int64_t *comm_out) {
  return cvb_louvain_run_ex(e, max_levels, 0, level_q, level_iters, level_nv, comm_out);
}

// distLouvainMethod with early termination (louvain.cpp:7-204 for et_type 1 / 3, 206-423 for 2 / 4) on one
// GPU.  The sweep kernels are the ordinary ones; they run with the community-delta atomics switched
// off and k_et_fix adds the deltas of the ACTIVE vertices only and gives the frozen ones their frozen
// target / clusterWeight back, so every kernel generation serves this mode unchanged.  Labels rotate
// per vertex (k_et_rotate), not by pointer.
int cvb_phase_run_et(cvb_engine *e, double lower, double thresh, int et_type, double et_delta, double *q_out,
                     int *iters_out) {
  CVB_CHECK(e && e->graph_ready, CVB_ERR_STATE, "cvb_phase_run_et: no graph uploaded");
  CVB_CHECK(e->nvl == e->nvg, CVB_ERR_STATE, "cvb_phase_run_et is the one-GPU driver");
  CVB_CHECK(et_type >= 1 && et_type <= 4, CVB_ERR_ARG, "early-termination type must be 1..4 (option -t)");
  CVB_CHECK(!(et_type == 2 || et_type == 4) || (et_delta >= 0.0 && et_delta <= 1.0), CVB_ERR_ARG,
            "early-termination delta must be in [0, 1] (option -a)");
  const bool prob = et_type == 2 || et_type == 4, cutoff = et_type == 3 || et_type == 4;
  cudaStream_t st = e->stream;
  const int64_t nvl = e->nvl;
  CVB_TRY(cvb_phase_begin(e, 0.0));
  CVB_TRY(e->et_active.ensure(nvl + 1));
  CVB_TRY(e->et_frozen_target.ensure(nvl + 1));
  CVB_TRY(e->et_frozen_cw.ensure(nvl + 1));
  CVB_TRY(e->et_count.ensure(4));
  CVB_CUDA(cudaMemsetAsync(e->et_active.p, 1, nvl + 1, st));
  CVB_CUDA(cudaMemsetAsync(e->et_count.p, 0, 4 * sizeof(unsigned int), st));
  if (prob) {
    for (auto &b : e->et_p) CVB_TRY(b.ensure(nvl + 1));
    if (nvl > 0) {
      k_fill_f64<<<grid_for(nvl, 256), 256, 0, st>>>(e->et_p[0].p, nvl, 1.0);  // p_prev = 1 (louvain.cpp:250)
      CVB_CUDA(cudaMemsetAsync(e->et_p[1].p, 0, nvl * sizeof(double), st));
    }
  }
  struct Restore { cvb_engine *e; ~Restore() { e->et_no_deltas = false; } } restore{e};
  e->et_no_deltas = true;
  double prev = lower, cur = -1.0;
  int iters = 0, pp = 0;
  unsigned int frozen = 0;  // vc_count of the iteration about to run = vertices frozen so far
  while (true) {
    iters++;
    CVB_TRY(cvb_sweep(e));
    if (nvl > 0)
      k_et_fix<<<grid_for(nvl, 256), 256, 0, st>>>(nvl, e->et_active.p, e->lab[e->targ].p, e->lab[e->curr].p,
                                                   e->et_frozen_target.p, e->cw.p, e->et_frozen_cw.p, e->vdeg.p,
                                                   e->ddeg.p, e->dsize.p);
    e->launches++;
    if (cutoff && frozen >= 90u) {  // ET_CUTOFF, louvain.hpp:76; louvain.cpp:114-120
      // the reference leaves the loop BEFORE distUpdateLocalCinfo: drop this sweep's deltas
      CVB_CUDA(cudaMemsetAsync(e->ddeg.p, 0, e->nvg * sizeof(double), st));
      CVB_CUDA(cudaMemsetAsync(e->dsize.p, 0, e->nvg * sizeof(int32_t), st));
      cur = -1.0;
      break;
    }
    double sums[2];
    CVB_TRY(cvb_apply_and_reduce(e, sums));
    cur = cvb_modularity_from_sums(e, sums);
    if ((cur - prev) < thresh) break;
    prev = cur;
    if (prev < lower) prev = lower;
    if (nvl > 0)
      k_et_rotate<<<grid_for(nvl, 256), 256, 0, st>>>(nvl, iters, prob ? 1 : 0, 1.0 - et_delta, e->lab[e->past].p,
                                                      e->lab[e->curr].p, e->lab[e->targ].p, e->et_active.p,
                                                      e->et_frozen_target.p, prob ? e->et_p[pp].p : nullptr,
                                                      prob ? e->et_p[pp ^ 1].p : nullptr, e->et_count.p);
    e->launches++;
    pp ^= 1;  // std::swap(p_curr, p_prev), louvain.cpp:401
    CVB_CUDA(cudaMemcpyAsync(&frozen, e->et_count.p, sizeof frozen, cudaMemcpyDeviceToHost, st));
    CVB_CUDA(cudaStreamSynchronize(st));
  }
  if (q_out) *q_out = prev;
  if (iters_out) *iters_out = iters;
  return CVB_OK;
}

// distColoringMultiHashMinMax (coloring.cpp:3-72) on one GPU: rounds of k_color_round with nHash =
// max_colors / 2 hash functions until 70 % (MAX_COVG, main.cpp:26) of the vertices are coloured or a round
// colours nothing; the colour ids of round r start at r * 2 * nHash and the seed is re-hashed per round.
// Returns the reference's numColors (= rounds * 2 * nHash) and leaves the colours in e->col_color.
static int color_graph(cvb_engine *e, int max_colors, int *num_colors_out) {
  const int n_hash = max_colors / 2;  // main.cpp:245
  CVB_CHECK(n_hash >= 0 && n_hash <= 16, CVB_ERR_ARG, "max_colors must be in 0..32 (option -c)");
  cudaStream_t st = e->stream;
  const int64_t nvl = e->nvl;
  CVB_TRY(e->col_color.ensure(nvl + 1));
  CVB_TRY(e->et_count.ensure(4));
  CVB_CUDA(cudaMemsetAsync(e->col_color.p, 0xff, (nvl + 1) * sizeof(int32_t), st));  // -1 = uncoloured
  CVB_CUDA(cudaMemsetAsync(e->et_count.p, 0, 4 * sizeof(unsigned int), st));
  unsigned int seed = 1012u, coloured = 0, last = 0;
  int next_color = 0;
  const int64_t target = (e->nvg * 70) / 100;
  while (true) {
    if (nvl > 0 && n_hash > 0) {
      k_color_round<<<static_cast<int>((nvl * 32 + 255) / 256), 256, 0, st>>>(
          e->offsets.p, e->tails.p, nvl, static_cast<vid_t>(e->base), n_hash, next_color, seed, e->col_color.p,
          e->et_count.p);
      e->launches++;
    }
    CVB_CUDA(cudaMemcpyAsync(&coloured, e->et_count.p, sizeof coloured, cudaMemcpyDeviceToHost, st));
    CVB_CUDA(cudaStreamSynchronize(st));
    const bool finished = static_cast<int64_t>(coloured) >= target || coloured == last;  // coloring.cpp:50
    next_color += 2 * n_hash;
    seed = color_hash(seed, 0u);
    last = coloured;
    if (finished) break;
  }
  *num_colors_out = next_color;
  return CVB_OK;
}

// per-colour degree-class lists: vertices sorted by (colour bucket, degree class), uncoloured in the last
// bucket (louvain.cpp:806-842), with the per-vertex records of lm.cuh in the same order
static int color_lists(cvb_engine *e, int num_color) {
  cudaStream_t st = e->stream;
  const int64_t nvl = e->nvl;
  const int nkeys = num_color * kNumClasses;
  e->col_num = num_color;
  CVB_TRY(e->col_keys.ensure(nvl + 1));
  CVB_TRY(e->col_keys2.ensure(nvl + 1));
  CVB_TRY(e->col_ids.ensure(nvl + 1));
  CVB_TRY(e->col_list.ensure(nvl + 1));
  e->col_start.assign(static_cast<size_t>(nkeys) + 1, 0);
  if (nvl == 0) return CVB_OK;
  k_color_keys<<<grid_for(nvl, 256), 256, 0, st>>>(e->offsets.p, e->col_color.p, nvl, num_color, e->col_keys.p,
                                                   e->col_ids.p);
  int bits = 1;
  while ((1 << bits) < nkeys) bits++;
  size_t tb = 0;
  CVB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, e->col_keys.p, e->col_keys2.p, e->col_ids.p, e->col_list.p,
                                           static_cast<int>(nvl), 0, bits, st));
  CVB_TRY(e->staging2.ensure(tb + static_cast<size_t>(nkeys) * sizeof(unsigned int) + 64));
  CVB_CUDA(cub::DeviceRadixSort::SortPairs(e->staging2.p, tb, e->col_keys.p, e->col_keys2.p, e->col_ids.p,
                                           e->col_list.p, static_cast<int>(nvl), 0, bits, st));
  unsigned int *hist = reinterpret_cast<unsigned int *>(e->staging2.p + ((tb + 15) / 16) * 16);
  CVB_CUDA(cudaMemsetAsync(hist, 0, static_cast<size_t>(nkeys) * sizeof(unsigned int), st));
  k_key_histogram<<<grid_for(nvl, 256), 256, 0, st>>>(e->col_keys.p, nvl, hist);
  e->launches += 3;
  std::vector<unsigned int> h(nkeys);
  CVB_CUDA(cudaMemcpyAsync(h.data(), hist, static_cast<size_t>(nkeys) * sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
  CVB_CUDA(cudaStreamSynchronize(st));
  for (int k = 0; k < nkeys; ++k) e->col_start[k + 1] = e->col_start[k] + h[k];
  if (e->lm_ready) {
    CVB_TRY(e->col_vrec.ensure(nvl + 1));
    k_build_vrec<<<grid_for(nvl, 256), 256, 0, st>>>(e->col_list.p, e->offsets.p, e->vdeg.p, nvl, e->col_vrec.p);
    e->launches++;
  }
  return CVB_OK;
}

// distLouvainMethodWithColoring (louvain.cpp:756-949) on one GPU.  Inside an iteration the colour
// buckets are swept one after the other: every step reads the labels of the iteration's start and the
// community info as the earlier colours of the same iteration left it (louvain.cpp:864-899); the new labels
// take effect together at the end of the iteration.  The sweep kernels are the ordinary ones, launched over
// the (colour, class) ranges of the per-colour lists.
int cvb_color_graph(cvb_engine *e, int max_colors, int *num_colors_out) {
  CVB_CHECK(e && e->graph_ready && num_colors_out, CVB_ERR_STATE, "cvb_color_graph: no graph uploaded");
  CVB_CHECK(e->nvl == e->nvg, CVB_ERR_STATE, "cvb_color_graph is the one-GPU colouring");
  return color_graph(e, max_colors, num_colors_out);
}

int cvb_phase_run_color(cvb_engine *e, double lower, double thresh, int max_colors, double *q_out, int *iters_out,
                        int *num_colors_out) {
  return cvb_phase_run_color_et(e, lower, thresh, max_colors, 0, 1.0, q_out, iters_out, num_colors_out);
}

// et_type 0: louvain.cpp:756-949.  et_type 1 / 3: louvain.cpp:951-1178, et_type 2 / 4: louvain.cpp:1180-1426 — the
// coloured phase with the per-vertex freezing of cvb_phase_run_et: every kernel of a colour step runs with the
// delta atomics off, k_et_fix_list / k_color_fix_big add the moves of the ACTIVE vertices of the colour and give the
// frozen ones their frozen target / clusterWeight back; labels rotate per vertex (k_et_rotate).
int cvb_phase_run_color_et(cvb_engine *e, double lower, double thresh, int max_colors, int et_type, double et_delta,
                           double *q_out, int *iters_out, int *num_colors_out) {
  CVB_CHECK(e && e->graph_ready, CVB_ERR_STATE, "cvb_phase_run_color: no graph uploaded");
  CVB_CHECK(e->nvl == e->nvg, CVB_ERR_STATE, "cvb_phase_run_color is the one-GPU driver");
  CVB_CHECK(et_type >= 0 && et_type <= 4, CVB_ERR_ARG, "early-termination type must be 0..4 (option -t)");
  CVB_CHECK(!(et_type == 2 || et_type == 4) || (et_delta >= 0.0 && et_delta <= 1.0), CVB_ERR_ARG,
            "early-termination delta must be in [0, 1] (option -a)");
  const bool et = et_type != 0, prob = et_type == 2 || et_type == 4, cutoff = et_type == 3 || et_type == 4;
  int num_colors = 0;
  CVB_TRY(color_graph(e, max_colors, &num_colors));
  if (num_colors_out) *num_colors_out = num_colors;
  const int num_color = num_colors + 1;  // main.cpp:259: the last bucket holds the uncoloured vertices
  if (num_color == 1 || e->nvl == 0)  // louvain.cpp:764-768, 958-962 (and the empty graph: no colour step would run)
    return cvb_phase_run(e, lower, thresh, 0, q_out, iters_out);
  CVB_TRY(color_lists(e, num_color));
  cudaStream_t st = e->stream;
  const int64_t nvl = e->nvl;
  const int64_t nbig = e->mode == CVB_WEIGHTS_U32 ? e->class_start[kClsGiant + 1] - e->class_start[kClsC2] : 0;
  const vid_t *big_list = e->list.p + e->class_start[kClsC2];
  if (nbig > 0) {
    CVB_TRY(e->col_bak_t.ensure(nbig));
    CVB_TRY(e->col_bak_cw.ensure(nbig));
  }
  CVB_TRY(cvb_phase_begin(e, 0.0));
  if (et) {
    CVB_TRY(e->et_active.ensure(nvl + 1));
    CVB_TRY(e->et_frozen_target.ensure(nvl + 1));
    CVB_TRY(e->et_frozen_cw.ensure(nvl + 1));
    CVB_TRY(e->et_count.ensure(4));
    CVB_CUDA(cudaMemsetAsync(e->et_active.p, 1, nvl + 1, st));
    CVB_CUDA(cudaMemsetAsync(e->et_frozen_cw.p, 0, (nvl + 1) * sizeof(double), st));
    CVB_CUDA(cudaMemsetAsync(e->et_count.p, 0, 4 * sizeof(unsigned int), st));
    if (prob) {
      for (auto &b : e->et_p) CVB_TRY(b.ensure(nvl + 1));
      k_fill_f64<<<grid_for(nvl, 256), 256, 0, st>>>(e->et_p[0].p, nvl, 1.0);  // p_prev = 1
      CVB_CUDA(cudaMemsetAsync(e->et_p[1].p, 0, nvl * sizeof(double), st));
    }
  }
  struct Restore { cvb_engine *e; ~Restore() { e->ovr = false; e->ovr_et = false; e->et_no_deltas = false; } } restore{e};
  e->ovr_et = et;
  e->et_no_deltas = et;
  double prev = lower, cur = -1.0;
  int iters = 0, pp = 0;
  unsigned int frozen = 0;  // vc_count of the iteration about to run = vertices frozen so far
  while (true) {
    iters++;
    CVB_CUDA(cudaMemsetAsync(e->cw.p, 0, nvl * sizeof(double), st));  // louvain.cpp:857-858
    if (nbig > 0) {
      k_color_bak_init<<<static_cast<int>((nbig + 127) / 128), 128, 0, st>>>(big_list, static_cast<int>(nbig),
                                                                            e->lab[e->targ].p, e->col_bak_t.p,
                                                                            e->col_bak_cw.p);
      e->launches++;
    }
    for (int ci = 0; ci < num_color; ++ci) {
      const int64_t *cs = e->col_start.data() + static_cast<size_t>(ci) * kNumClasses;
      if (cs[kNumClasses] == cs[0]) continue;  // no vertex of this colour
      e->ovr = true;
      e->ovr_color = ci;
      e->ovr_big = cs[kNumClasses] > cs[kClsC2];
      const int rc = cvb_sweep(e);
      e->ovr = false;
      CVB_TRY(rc);
      if (et) {  // the listed vertices of the colour (U32 mode: the big ones are handled by k_color_fix_big)
        const int64_t n = (e->mode == CVB_WEIGHTS_U32 ? cs[kClsC2] : cs[kNumClasses]) - cs[0];
        if (n > 0) {
          k_et_fix_list<<<grid_for(n, 256), 256, 0, st>>>(e->col_list.p + cs[0], n, e->et_active.p, e->lab[e->targ].p,
                                                         e->lab[e->curr].p, e->et_frozen_target.p, e->cw.p,
                                                         e->et_frozen_cw.p, e->vdeg.p, e->ddeg.p, e->dsize.p);
          e->launches++;
        }
      }
      CVB_TRY(apply_reduce_async(e));  // louvain.cpp:887-893; the sums of the last colour are the iteration's
    }
    CVB_CUDA(cudaStreamSynchronize(st));
    double sums[3];
    memcpy(sums, e->h_sums.p, sizeof sums);
    CVB_CHECK(sums[2] == 0.0, CVB_ERR_CAPACITY, "a partition bucket of the big-vertex path overflowed; rerun with CVB_BIG=0");
    if (cutoff && frozen >= 90u) {  // ET_CUTOFF; louvain.cpp:1131-1138 (here AFTER the community updates of the colours)
      cur = -1.0;
      break;
    }
    cur = cvb_modularity_from_sums(e, sums);
    if ((cur - prev) < thresh) break;
    prev = cur;
    if (prev < lower) prev = lower;
    if (!et) {
      CVB_TRY(cvb_rotate_labels(e));
      continue;
    }
    k_et_rotate<<<grid_for(nvl, 256), 256, 0, st>>>(nvl, iters, prob ? 1 : 0, 1.0 - et_delta, e->lab[e->past].p,
                                                    e->lab[e->curr].p, e->lab[e->targ].p, e->et_active.p,
                                                    e->et_frozen_target.p, prob ? e->et_p[pp].p : nullptr,
                                                    prob ? e->et_p[pp ^ 1].p : nullptr, e->et_count.p);
    e->launches++;
    pp ^= 1;
    CVB_CUDA(cudaMemcpyAsync(&frozen, e->et_count.p, sizeof frozen, cudaMemcpyDeviceToHost, st));
    CVB_CUDA(cudaStreamSynchronize(st));
  }
  if (q_out) *q_out = prev;
  if (iters_out) *iters_out = iters;
  return CVB_OK;
}

static int louvain_run_impl(cvb_engine *e, int max_levels, int flags, int et_type, double et_delta, int max_colors,
                            double *level_q, int *level_iters, int64_t *level_nv, int64_t *comm_out);

int cvb_louvain_run_ex(cvb_engine *e, int max_levels, int flags, double *level_q, int *level_iters,
                       int64_t *level_nv, int64_t *comm_out) {
  return louvain_run_impl(e, max_levels, flags, 0, 1.0, -1, level_q, level_iters, level_nv, comm_out);
}

int cvb_louvain_run_et(cvb_engine *e, int max_levels, int flags, int et_type, double et_delta, double *level_q,
                       int *level_iters, int64_t *level_nv, int64_t *comm_out) {
  return louvain_run_impl(e, max_levels, flags, et_type, et_delta, -1, level_q, level_iters, level_nv, comm_out);
}

int cvb_louvain_run_color(cvb_engine *e, int max_levels, int flags, int max_colors, double *level_q,
                          int *level_iters, int64_t *level_nv, int64_t *comm_out) {
  return louvain_run_impl(e, max_levels, flags, 0, 1.0, max_colors, level_q, level_iters, level_nv, comm_out);
}

int cvb_louvain_run_color_et(cvb_engine *e, int max_levels, int flags, int max_colors, int et_type, double et_delta,
                             double *level_q, int *level_iters, int64_t *level_nv, int64_t *comm_out) {
  return louvain_run_impl(e, max_levels, flags, et_type, et_delta, max_colors, level_q, level_iters, level_nv, comm_out);
}

static int louvain_run_impl(cvb_engine *e, int max_levels, int flags, int et_type, double et_delta, int max_colors,
                            double *level_q, int *level_iters, int64_t *level_nv, int64_t *comm_out) {
  if (!e || !e->graph_ready) { set_error("cvb_louvain_run: no graph uploaded"); return CVB_ERR_STATE; }
  if (e->nvl != e->nvg) { set_error("cvb_louvain_run is the one-GPU driver"); return CVB_ERR_STATE; }
  cudaStream_t st = e->stream;
  const int64_t nv0 = e->nvl;
  DevBuf<int32_t> compose;
  struct Guard { DevBuf<int32_t> &b; ~Guard() { b.release(); } } guard{compose};
  if (comm_out && nv0 > 0) {
    if (compose.ensure(nv0) != CVB_OK) return CVB_ERR_CUDA;
    k_iota_i32<<<grid_for(nv0, 256), 256, 0, st>>>(compose.p, nv0);
    e->launches++;
  }
  double curr = -1.0, prev = -1.0;  // main.cpp:177
  int phase = 0, tot_iters = 0;
  double threshold = 1.0e-6;  // main.cpp:239
  while (true) {
    const int64_t nv = e->nvl;
    int iters = 0;
    if (flags & CVB_RUN_THRESHOLD_CYCLING) {  // option -i, main.cpp:225-236 (short_phase == phase below the cap)
      const int sp = phase % 13;
      threshold = sp <= 2 ? 1.0e-3 : sp <= 6 ? 1.0e-4 : sp <= 9 ? 1.0e-5 : 1.0e-6;
    }
    int rc = (max_colors >= 0 && phase == 0)  // option -c: the first phase only, main.cpp:242-282
                 ? cvb_phase_run_color_et(e, curr, threshold, max_colors, et_type, et_delta, &curr, &iters, nullptr)
                 : et_type ? cvb_phase_run_et(e, curr, threshold, et_type, et_delta, &curr, &iters)  // main.cpp:326-341
                           : cvb_phase_run(e, curr, threshold, 0, &curr, &iters);                    // main.cpp:344
    if (rc != CVB_OK) return rc;
    tot_iters += iters;
    if ((curr - prev) > threshold) {  // main.cpp:361
      rc = coarsen_impl(e, nullptr, comm_out ? compose.p : nullptr, nv0);  // main.cpp:413
      if (rc != CVB_OK) return rc;
    } else {
      break;
    }
    if (phase < max_levels) {  // main.cpp:448-450
      if (level_q) level_q[phase] = curr;
      if (level_iters) level_iters[phase] = tot_iters;
      if (level_nv) level_nv[phase] = nv;
    }
    prev = curr;
    curr = -1.0;
    phase++;
    if (phase == 200 || tot_iters > 10000) break;  // main.cpp:486
  }
  if (comm_out && nv0 > 0) {
    DevBuf<int64_t> wide;
    if (wide.ensure(nv0) != CVB_OK) return CVB_ERR_CUDA;
    k_widen<int32_t, int64_t><<<grid_for(nv0, 256), 256, 0, st>>>(compose.p, wide.p, nv0);
    e->launches++;
    cudaError_t err = cudaMemcpyAsync(comm_out, wide.p, nv0 * sizeof(int64_t), cudaMemcpyDeviceToHost, st);
    if (err == cudaSuccess) err = cudaStreamSynchronize(st);
    wide.release();
    if (err != cudaSuccess) { set_error("cvb_louvain_run: D2H failed: %s", cudaGetErrorString(err)); return CVB_ERR_CUDA; }
  }
  return phase < max_levels ? phase : max_levels;  // the level arrays hold max_levels entries
}

// ---------------------------------------------------------------------------
// N > 1 ranks: NCCL phase driver (dist.cuh).  Every function is collective: all
// ranks call it in the same order.
// ---------------------------------------------------------------------------
int cvb_dist_unique_id(void *out, int nbytes) {
  CVB_CHECK(out && nbytes >= static_cast<int>(sizeof(ncclUniqueId)), CVB_ERR_ARG, "cvb_dist_unique_id: need %d bytes",
            static_cast<int>(sizeof(ncclUniqueId)));
  NcclApi *n = nccl_api();
  CVB_CHECK(n, CVB_ERR_STATE, "NCCL not available (dlopen libnccl.so.2 failed; set CVB_NCCL_LIB)");
  ncclUniqueId id;
  CVB_NCCL(n->GetUniqueId(&id));
  memcpy(out, &id, sizeof id);
  return CVB_OK;
}

int cvb_dist_init(cvb_engine *e, const void *id, int nbytes, int rank, int world) {
  CVB_CHECK(e && world >= 1 && rank >= 0 && rank < world, CVB_ERR_ARG, "cvb_dist_init: bad rank/world");
  DistState &d = e->dist;
  d.rank = rank;
  d.world = world;
  d.timing = getenv("CVB_DIST_TIMING") && getenv("CVB_DIST_TIMING")[0] == '1';
  if (d.timing && !d.ev[0])
    for (auto &x : d.ev) CVB_CUDA(cudaEventCreate(&x));
  if (world == 1) return CVB_OK;
  CVB_CHECK(id && nbytes >= static_cast<int>(sizeof(ncclUniqueId)), CVB_ERR_ARG, "cvb_dist_init: missing unique id");
  NcclApi *n = nccl_api();
  CVB_CHECK(n, CVB_ERR_STATE, "NCCL not available (dlopen libnccl.so.2 failed; set CVB_NCCL_LIB)");
  ncclUniqueId uid;
  memcpy(&uid, id, sizeof uid);
  CVB_CUDA(cudaSetDevice(e->device));
  // This communicator carries one 4 B/vertex all-gather and one 16-byte all-reduce per sweep.  Measured on 8
  // B200s (profiles/r02_bench_n8*.json): with NVLink-SHARP (NVLS) channels set up, the label all-gather of a
  // sweep took 0.70 ms against 0.25 ms without them (0.09 ms in isolation either way) — so NVLS is off unless
  // the host program asked for it.  NCCL reads the variable when the first communicator of the process
  // initialises; a host that created one earlier (torch.distributed) sets it itself (cuvite_b200/dist.py does).
  setenv("NCCL_NVLS_ENABLE", "0", /*overwrite=*/0);
  CVB_NCCL(n->CommInitRank(&d.comm, world, uid, rank));
  return CVB_OK;
}

int cvb_dist_finalize(cvb_engine *e) {
  CVB_CHECK(e, CVB_ERR_ARG, "null engine");
  if (e->dist.comm) {
    cudaStreamSynchronize(e->stream);
    nccl_api()->CommDestroy(e->dist.comm);
    e->dist.comm = nullptr;
  }
  return CVB_OK;
}

int cvb_dist_set_parts(cvb_engine *e, const int64_t *parts) {
  CVB_CHECK(e && e->graph_ready, CVB_ERR_STATE, "cvb_dist_set_parts: no graph uploaded");
  DistState &d = e->dist;
  d.parts.resize(d.world + 1);
  for (int p = 0; p <= d.world; ++p)
    d.parts[p] = parts ? parts[p] : (e->nvg * p) / d.world;  // distgraph.cpp:118-119
  CVB_CHECK(d.parts[0] == 0 && d.parts[d.world] == e->nvg && d.parts[d.rank] == e->base &&
                d.parts[d.rank + 1] == e->base + e->nvl,
            CVB_ERR_ARG, "cvb_dist_set_parts: ranges do not match the uploaded rows [%lld, %lld)", (long long)e->base,
            (long long)(e->base + e->nvl));
  d.even = true;
  for (int p = 0; p < d.world; ++p)
    if (d.parts[p + 1] - d.parts[p] != d.parts[1] - d.parts[0]) d.even = false;
  return CVB_OK;
}

namespace {
// global[parts[p] .. parts[p+1]) <- rank p's local rows, for every p (device buffers, `elem` bytes per row)
int dist_allgather_rows(cvb_engine *e, const void *local, void *global, int dtype, size_t elem) {
  DistState &d = e->dist;
  if (d.world == 1) {
    if (local != global)
      CVB_CUDA(cudaMemcpyAsync(static_cast<char *>(global) + e->base * elem, local, e->nvl * elem,
                               cudaMemcpyDeviceToDevice, e->stream));
    return CVB_OK;
  }
  NcclApi *n = nccl_api();
  if (d.even) {
    CVB_NCCL(n->AllGather(local, global, static_cast<size_t>(e->nvl), dtype, d.comm, e->stream));
    return CVB_OK;
  }
  CVB_NCCL(n->GroupStart());  // uneven ranges (coarsened levels, -b): one broadcast per owner, fused in a group
  for (int p = 0; p < d.world; ++p) {
    const size_t cnt = static_cast<size_t>(d.parts[p + 1] - d.parts[p]);
    if (cnt == 0) continue;
    char *dst = static_cast<char *>(global) + d.parts[p] * elem;
    n->Broadcast(p == d.rank ? local : dst, dst, cnt, dtype, p, d.comm, e->stream);
  }
  CVB_NCCL(n->GroupEnd());
  return CVB_OK;
}

int dist_allreduce(cvb_engine *e, void *buf, size_t count, int dtype, int op) {
  if (e->dist.world == 1) return CVB_OK;
  CVB_NCCL(nccl_api()->AllReduce(buf, buf, count, dtype, op, e->dist.comm, e->stream));
  return CVB_OK;
}
}  // namespace

int cvb_dist_agree(cvb_engine *e) {
  CVB_CHECK(e && e->graph_ready, CVB_ERR_STATE, "cvb_dist_agree: no graph uploaded");
  DistState &d = e->dist;
  if (d.parts.empty()) CVB_TRY(cvb_dist_set_parts(e, nullptr));
  if (d.world == 1) return CVB_OK;
  // the exact-integer mode is only valid when EVERY rank's weights are integers and the GLOBAL total
  // stays below 2^32 (the replicated community degrees are global sums): agree on it
  double h[2] = {e->mode == CVB_WEIGHTS_U32 ? 1.0 : 0.0, e->local_weight};
  cudaStream_t st = e->stream;
  CVB_CUDA(cudaMemcpyAsync(e->d_sums.p, h, sizeof h, cudaMemcpyHostToDevice, st));
  CVB_TRY(dist_allreduce(e, e->d_sums.p, 1, ncclFloat64, ncclMin));
  CVB_TRY(dist_allreduce(e, e->d_sums.p + 1, 1, ncclFloat64, ncclSum));
  CVB_CUDA(cudaMemcpyAsync(h, e->d_sums.p, sizeof h, cudaMemcpyDeviceToHost, st));
  CVB_CUDA(cudaStreamSynchronize(st));
  const bool all_u32 = h[0] > 0.5 && h[1] < 4294967296.0;
  if (!all_u32 && e->mode == CVB_WEIGHTS_U32) {
    DevBuf<unsigned char> w2;
    CVB_TRY(w2.ensure((static_cast<size_t>(e->ne) + 1) * 8));
    if (e->ne > 0)
      k_weights_u32_to_f64<<<grid_for(e->ne, 256), 256, 0, st>>>(reinterpret_cast<const uint32_t *>(e->weights.p),
                                                                 reinterpret_cast<double *>(w2.p), e->ne);
    CVB_CUDA(cudaStreamSynchronize(st));
    std::swap(e->weights.p, w2.p);
    std::swap(e->weights.cap, w2.cap);
    w2.release();
    e->mode = CVB_WEIGHTS_F64;
    CVB_TRY(post_graph(e));
  }
  return CVB_OK;
}

int cvb_dist_phase_begin(cvb_engine *e) {
  CVB_CHECK(e && e->graph_ready, CVB_ERR_STATE, "cvb_dist_phase_begin: no graph uploaded");
  DistState &d = e->dist;
  if (d.parts.empty() || d.parts[d.world] != e->nvg) CVB_TRY(cvb_dist_set_parts(e, nullptr));
  cudaStream_t st = e->stream;
  double tot = e->local_weight;  // distCalcConstantForSecondTerm's Allreduce, louvain.cpp:2175
  if (d.world > 1) {
    CVB_CUDA(cudaMemcpyAsync(e->d_sums.p, &tot, sizeof tot, cudaMemcpyHostToDevice, st));
    CVB_TRY(dist_allreduce(e, e->d_sums.p, 1, ncclFloat64, ncclSum));
    CVB_CUDA(cudaMemcpyAsync(&tot, e->d_sums.p, sizeof tot, cudaMemcpyDeviceToHost, st));
    CVB_CUDA(cudaStreamSynchronize(st));
  }
  d.total_weight = tot;
  CVB_TRY(cvb_phase_begin(e, tot));
  if (d.world > 1) {  // community degrees of every vertex: all-gather of the vertex degrees
    CVB_TRY(dist_allgather_rows(e, e->vdeg.p, e->cdeg.p, ncclFloat64, sizeof(double)));
    CVB_CUDA(cudaMemcpyAsync(e->vdeg_global.p, e->cdeg.p, e->nvg * sizeof(double), cudaMemcpyDeviceToDevice, st));
    // degree-ordered gather layout over the GLOBAL vertices (once per graph): the replicated packed[]
    // array is as large as the whole graph on every rank, and the partitioned inputs carry permuted
    // ids — without it the records of the hubs are scattered over all of it
    if (e->lm_ready && e->use_remap && !e->remap && e->mode == CVB_WEIGHTS_U32) {
      CVB_TRY(e->rank.ensure(e->nvg + 1));
      CVB_TRY(e->tails_r.ensure(static_cast<size_t>(e->ne) + 1));
      CVB_TRY(e->staging2.ensure(64 * sizeof(unsigned int)));
      unsigned int *hist = reinterpret_cast<unsigned int *>(e->staging2.p);
      CVB_CUDA(cudaMemsetAsync(hist, 0, 64 * sizeof(unsigned int), st));
      k_degree_hist<<<grid_for(e->nvg, 256), 256, 0, st>>>(e->vdeg_global.p, e->nvg, hist);
      k_degree_prefix<<<1, 32, 0, st>>>(hist);
      k_degree_rank<<<grid_for(e->nvg, 256), 256, 0, st>>>(e->vdeg_global.p, e->nvg, hist, e->rank.p);
      if (e->ne > 0)
        k_remap_tails<<<grid_for(e->ne, 256), 256, 0, st>>>(e->tails.p, e->rank.p, e->ne, e->tails_r.p);
      e->launches += 4;
      e->remap = true;
    }
  }
  return CVB_OK;
}

int cvb_dist_sweep(cvb_engine *e, double *q_out) {
  CVB_CHECK(e && e->phase_ready && q_out, CVB_ERR_STATE, "cvb_dist_sweep: call cvb_dist_phase_begin first");
  DistState &d = e->dist;
  cudaStream_t st = e->stream;
  double sums[2];
  if (d.world == 1) {
    CVB_TRY(cvb_sweep(e));
    CVB_TRY(cvb_apply_and_reduce(e, sums));
    *q_out = cvb_modularity_from_sums(e, sums);
    return CVB_OK;
  }
  if (d.timing) CVB_CUDA(cudaEventRecord(d.ev[0], st));
  CVB_TRY(cvb_sweep(e));
  if (d.timing) CVB_CUDA(cudaEventRecord(d.ev[1], st));
  // ghost labels: ONE all-gather of the new labels (replaces fillRemoteCommunities' label pull)
  CVB_TRY(dist_allgather_rows(e, e->lab[e->targ].p, e->glabels_next(), ncclInt32, sizeof(vid_t)));
  if (d.timing) CVB_CUDA(cudaEventRecord(d.ev[2], st));
  // community info of ALL ranks' moves from the label diff, locally (replaces updateRemoteCommunities)
  CVB_TRY(cvb_rebuild_and_reduce_async(e));
  if (d.timing) CVB_CUDA(cudaEventRecord(d.ev[3], st));
  CVB_TRY(dist_allreduce(e, e->d_sums.p, 2, ncclFloat64, ncclSum));  // louvain.cpp:2522
  if (d.timing) CVB_CUDA(cudaEventRecord(d.ev[4], st));
  CVB_CUDA(cudaMemcpyAsync(e->h_sums.p, e->d_sums.p, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (d.timing) CVB_CUDA(cudaEventRecord(d.ev[5], st));
  CVB_CUDA(cudaStreamSynchronize(st));
  memcpy(sums, e->h_sums.p, sizeof sums);
  if (d.timing) {
    for (int i = 0; i < 5; ++i) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, d.ev[i], d.ev[i + 1]);
      d.stage_ms[i] += ms;
    }
    d.stage_n++;
  }
  *q_out = cvb_modularity_from_sums(e, sums);
  return CVB_OK;
}

int cvb_dist_advance(cvb_engine *e) { return cvb_rotate_labels(e); }

int cvb_dist_phase_run(cvb_engine *e, double lower, double thresh, int max_iters, double *q_out, int *iters_out) {
  CVB_TRY(cvb_dist_phase_begin(e));
  double prev = lower, cur = -1.0;
  int iters = 0;
  while (true) {  // louvain.cpp:638-739
    iters++;
    CVB_TRY(cvb_dist_sweep(e, &cur));
    if ((cur - prev) < thresh) break;
    prev = cur;
    if (prev < lower) prev = lower;
    CVB_TRY(cvb_dist_advance(e));
    if (max_iters > 0 && iters >= max_iters) break;
  }
  if (q_out) *q_out = prev;
  if (iters_out) *iters_out = iters;
  return CVB_OK;
}

int cvb_dist_stage_ms(cvb_engine *e, double out[5]) {
  CVB_CHECK(e && out, CVB_ERR_ARG, "null argument");
  for (int i = 0; i < 5; ++i) out[i] = e->dist.stage_n ? e->dist.stage_ms[i] / e->dist.stage_n : 0.0;
  return CVB_OK;
}

}  // extern "C"
