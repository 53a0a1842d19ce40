/* cuvite_b200.h — C ABI of the B200-native Louvain local-move engine.
 *
 * Drop-in scope: the GPU offload library of pnnl/cuVite ("cuda_louvain":
 * louvain_cuda.cu, modularity.cu, GpuGraph.cu) and the GPU overload of the
 * phase loop that calls it.  Each entry point cites the reference interface it
 * replaces (paths relative to the reference tree).  Plain pointers and sizes
 * only; ids are int64 and weights are double at this boundary, exactly the
 * reference's 64-bit build (edge.hpp:15-19).  The C++ source-level adapter that
 * re-creates the reference's own symbols (CuVite::GpuGraph,
 * gpu_for_louvain_iteration, gpu_reduce_mod, set_gpuDevices,
 * distLouvainMethod(..., GpuGraph&)) on top of this ABI is
 * include/GpuGraph.cuh + cuvite_b200/csrc/adapter.cpp; see INTEGRATION.md.
 *
 * Error model: every int-returning function returns 0 on success and a
 * negative cvb_status otherwise; cvb_last_error() gives the message.  (The
 * reference printf()s and exit(1)s on any CUDA error, GpuGraph.cuh:23; the
 * C++ adapter reproduces that on a non-zero status.)  There is NO CPU
 * fallback: without a CUDA device every compute entry point fails.
 *
 * Threading: an engine is used from one host thread at a time (the reference
 * calls its GPU library from the rank's main thread, louvain.cpp:665).
 */
#ifndef CUVITE_B200_H
#define CUVITE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cvb_engine cvb_engine;

typedef enum {
  CVB_OK = 0,
  CVB_ERR_CUDA = -1,      /* a CUDA runtime call failed */
  CVB_ERR_ARG = -2,       /* invalid argument / call order */
  CVB_ERR_CAPACITY = -3,  /* graph exceeds a 32-bit device id / offset limit */
  CVB_ERR_STATE = -4      /* engine not in the state the call needs */
} cvb_status;

/* Arithmetic mode chosen per graph at upload (cvb_graph_info.weight_mode). */
enum { CVB_WEIGHTS_U32 = 1 /* all weights non-negative integers, total < 2^32: exact */,
       CVB_WEIGHTS_F64 = 2 /* arbitrary doubles */ };

const char *cvb_last_error(void);
const char *cvb_version(void);

/* Number of visible CUDA devices (0 when there is none; never fails). */
int cvb_device_count(void);

/* Replaces set_gpuDevices(int *me) — louvain_cuda.cu:1634-1669, declared
 * louvain.hpp:249, called main.cpp:73: bind this process to device
 * local_rank % ndev.  (The reference derives the node-local rank from
 * MPI_Comm_split_type; here the launcher's LOCAL_RANK is passed in.) */
int cvb_set_device(int local_rank);

/* Replaces the CuVite::GpuGraph constructor + devMemAllocate() +
 * pinnedAndDevMemAllocate() — GpuGraph.cuh:38-45, GpuGraph.cu:9-101, called
 * main.cpp:209-211.  Buffers are sized lazily per graph, grown never shrunk,
 * and freed by cvb_destroy (the reference never frees).  `stream` is a
 * cudaStream_t (0 = the engine creates its own). */
cvb_engine *cvb_create(void *stream);
void cvb_destroy(cvb_engine *e);

/* ---- graph upload (once per phase; the reference re-uploads the whole edge
 * list every iteration, louvain_cuda.cu:2403-2438) ------------------------- */

/* Rows [base, base+nv_local) of a graph with nv_global vertices, Vite CSR
 * layout in HOST memory: offsets[nv_local+1] rebased to 0 and tails/weights
 * either as the reference's AoS `Edge{int64 tail; double weight}` array
 * (edges_aos != NULL; graph.hpp:35-36, edge.hpp:22-30) or as two SoA arrays.
 * Replaces the flatten + H2D part of gpu_for_louvain_iteration,
 * louvain_cuda.cu:2244-2438. */
int cvb_graph_upload_host(cvb_engine *e, int64_t nv_local, int64_t base, int64_t nv_global,
                          const int64_t *offsets, const void *edges_aos,
                          const int64_t *tails, const double *weights);

/* Same, from DEVICE memory (direct-to-GPU load path): offsets int64, tails
 * int64, weights double, all device pointers. */
int cvb_graph_upload_device(cvb_engine *e, int64_t nv_local, int64_t base, int64_t nv_global,
                            const int64_t *d_offsets, const int64_t *d_tails,
                            const double *d_weights);

typedef struct {
  int64_t nv_local, base, nv_global, ne_local;
  int32_t weight_mode;        /* CVB_WEIGHTS_* */
  int32_t num_classes;        /* degree classes in use */
  int64_t class_count[24];    /* vertices per degree class */
  int64_t class_entries[24];  /* CSR entries per degree class */
  int64_t max_degree;
  double local_weight;        /* sum of local vertex degrees (the 2m share) */
} cvb_graph_info;
int cvb_graph_get_info(cvb_engine *e, cvb_graph_info *out);

/* ---- one phase, resident on the device ---------------------------------- */

/* Replaces distInitLouvain — louvain.cpp:2103-2124 (distSumVertexDegree
 * 2126-2151, distCalcConstantForSecondTerm 2153-2183, distInitComm
 * 2569-2586): labels = own id, Cinfo = {1, vDegree}, constant = 1/total.
 * total_weight = the global sum of vertex degrees (pass <= 0 on one GPU to use
 * the local sum).  On >1 GPU the caller then fills the replicated community
 * degree array through cvb_buffer (see cuvite_b200/dist.py). */
int cvb_phase_begin(cvb_engine *e, double total_weight);

/* One synchronous local-move sweep over the local rows: replaces
 * distCleanCWandCU + the kernels of gpu_for_louvain_iteration
 * (louvain_cuda.cu:2463-2699) + updateLocalTarget_gpu (49-173); semantics of
 * distExecuteLouvainIteration (louvain.cpp:2246-2382).  Reads labels/Cinfo of
 * iteration k, writes target[], clusterWeight[], and ADDS the community
 * deltas into the (zeroed) delta arrays.  Asynchronous on the engine stream. */
int cvb_sweep(cvb_engine *e);

/* Replaces distUpdateLocalCinfo (louvain.cpp:2539-2552) + gpu_reduce_mod /
 * compute_modularity (modularity.cu:36-141): Cinfo += delta (all nv_global
 * communities), delta = 0, then sums[0] = sum of clusterWeight over local
 * vertices, sums[1] = sum of degree^2 over the communities this rank owns
 * ([base, base+nv_local)).  Synchronises the stream; sums is a HOST pointer. */
int cvb_apply_and_reduce(cvb_engine *e, double sums[2]);

/* N > 1, labels-only exchange.  After the sweep the ranks all-gather their
 * target labels into CVB_BUF_LABELS_GLOBAL_NEXT; this call then REBUILDS the
 * replicated community info from those labels and the replicated vertex
 * degrees (CVB_BUF_VDEG_GLOBAL) instead of reducing per-sweep deltas across
 * ranks — it replaces updateRemoteCommunities (louvain.cpp:2983-3116) and the
 * remote half of fillRemoteCommunities (2696-2954) with zero bytes on the wire —
 * and returns the same two sums as cvb_apply_and_reduce.  cvb_rotate_labels
 * then also swaps the two global label arrays. */
int cvb_rebuild_and_reduce(cvb_engine *e, double sums[2]);
/* Same, asynchronous: the sums stay in CVB_BUF_SUMS on the device (so that the
 * host side can all-reduce them with NCCL and read back once). */
int cvb_rebuild_and_reduce_async(cvb_engine *e);

/* Q = e*c - a2*c*c (louvain.cpp:2472-2473) from globally reduced sums. */
double cvb_modularity_from_sums(const cvb_engine *e, const double sums[2]);

/* past <- curr <- target <- past (louvain.cpp:562-567); O(1) pointer rotation. */
int cvb_rotate_labels(cvb_engine *e);

/* Whole phase on one GPU: replaces distLouvainMethod(..., GpuGraph&) —
 * louvain.hpp:93-97, louvain.cpp:591-753, called main.cpp:344.  Returns the
 * accepted modularity in *q_out, iterations in *iters_out; labels of the last
 * accepted iteration (cvect = pastComm, louvain.cpp:741) stay on the device
 * (cvb_get_labels).  max_iters <= 0 means unlimited. */
int cvb_phase_run(cvb_engine *e, double lower, double thresh, int max_iters,
                  double *q_out, int *iters_out);

enum { CVB_LABELS_PAST = 0, CVB_LABELS_CURR = 1, CVB_LABELS_TARGET = 2 };
/* D2H of a local label vector as int64 global ids (host pointer, nv_local). */
int cvb_get_labels(cvb_engine *e, int which, int64_t *out);
/* D2H of clusterWeight[nv_local] / community {size,degree}[nv_global]. */
int cvb_get_cluster_weight(cvb_engine *e, double *out);
int cvb_get_cinfo(cvb_engine *e, int64_t *size_out, double *degree_out);

/* Reset labels/Cinfo to the start-of-phase state without re-uploading. */
int cvb_phase_reset(cvb_engine *e);

/* ---- raw device buffers, for NCCL collectives issued by the host side ---- */
enum {
  CVB_BUF_LABELS_GLOBAL = 0, /* int32[nv_global] labels every sweep reads          */
  CVB_BUF_LABELS_CURR = 1,   /* int32[nv_local]                                     */
  CVB_BUF_LABELS_TARGET = 2, /* int32[nv_local]                                     */
  CVB_BUF_LABELS_PAST = 3,   /* int32[nv_local]                                     */
  CVB_BUF_CDEG = 4,          /* double[nv_global] community degree (replicated)     */
  CVB_BUF_CSIZE = 5,         /* int32[nv_global]  community size   (replicated)     */
  CVB_BUF_DDEG = 6,          /* double[nv_global] degree deltas of this sweep       */
  CVB_BUF_DSIZE = 7,         /* int32[nv_global]  size deltas of this sweep         */
  CVB_BUF_VDEG = 8,          /* double[nv_local]                                    */
  CVB_BUF_CW = 9,            /* double[nv_local]  clusterWeight                     */
  CVB_BUF_OFFSETS = 10,      /* int64[nv_local+1]                                   */
  CVB_BUF_TAILS = 11,        /* int32[ne_local]                                     */
  CVB_BUF_WEIGHTS = 12,      /* uint32 or double [ne_local], see weight_mode        */
  CVB_BUF_LABELS_GLOBAL_NEXT = 13, /* int32[nv_global] gather target for the NEXT labels (N>1) */
  CVB_BUF_VDEG_GLOBAL = 14,  /* double[nv_global] every vertex's degree (N>1, replicated)   */
  CVB_BUF_SUMS = 15,         /* double[2] device copy of the last {sum clusterWeight, sum degree^2} */
  CVB_BUF_PAIR_KEYS = 16,    /* uint64[n_pairs] (src << 32 | dst) of cvb_contract_local, ascending */
  CVB_BUF_PAIR_WEIGHTS = 17  /* uint32 or double [n_pairs], see weight_mode                        */
};
void *cvb_buffer(cvb_engine *e, int which);
void *cvb_stream(cvb_engine *e);
int cvb_sync(cvb_engine *e);

/* ---- the reference's per-iteration plugin call, flattened to C ----------- */

/* Replaces gpu_for_louvain_iteration — louvain.hpp:232-245,
 * louvain_cuda.cu:1671-1686 / 2235-2735, called louvain.cpp:665-668 — with the
 * STL containers flattened (HOST pointers):
 *   curr_comm[nv]            currComm (global ids)
 *   v_degree[nv]             vDegree (read when num_iters == 1)
 *   local_csize/cdeg[nv]     localCinfo
 *   n_ghost, ghost_vertex[], ghost_comm[]     remoteComm  (vertex -> community)
 *   n_rcomm, rcomm_id[], rcomm_size[], rcomm_degree[]   remoteCinfo (ascending id,
 *                            i.e. std::map order, louvain_cuda.cu:2260-2290)
 *   constant                 constantForSecondTerm
 *   num_iters                1-based; 1 => `graph` (offsets/edges) is new and uploaded
 * outputs:
 *   target_comm[nv], cluster_weight[nv] (= counter[0], louvain.cpp:2294),
 *   local_cupd_size/deg[nv]  += deltas (localCupdate),
 *   rcomm_upd_size/deg[n_rcomm] += deltas (remoteCupdate, same order as rcomm_id).
 * Returns 1 on success like the reference (louvain_cuda.cu:1685), <= 0 on error. */
int cvb_louvain_iteration(cvb_engine *e, int64_t nv, int64_t base, int64_t nv_global,
                          const int64_t *offsets, const void *edges_aos,
                          const int64_t *curr_comm, const double *v_degree,
                          const int64_t *local_csize, const double *local_cdeg,
                          int64_t n_ghost, const int64_t *ghost_vertex, const int64_t *ghost_comm,
                          int64_t n_rcomm, const int64_t *rcomm_id, const int64_t *rcomm_size,
                          const double *rcomm_degree, double constant, int num_iters,
                          int64_t *target_comm, double *cluster_weight,
                          int64_t *local_cupd_size, double *local_cupd_deg,
                          int64_t *rcomm_upd_size, double *rcomm_upd_deg);

/* Replaces gpu_reduce_mod — louvain.hpp:247-248, modularity.cu:102-141, called
 * louvain.cpp:2496: le_la_xx[0] = sum clusterWeight, [1] = sum degree^2 over nv
 * local entries (HOST pointers). */
int cvb_reduce_mod(cvb_engine *e, int64_t nv, const double *local_cdeg,
                   const double *cluster_weight, double le_la_xx[2]);

/* ---- coarsening (rebuild.cpp) and the multi-phase driver (main.cpp) ------ */

/* Replaces distbuildNextLevelGraph on one rank — rebuild.hpp:59,
 * rebuild.cpp:430-454: renumber the communities of the PAST labels in order of
 * first appearance (rebuild.cpp:72-85), aggregate edges, emit the CSR sorted by
 * tail (rebuild.cpp:413-426) and make it the engine's graph.  new_nv_out gets
 * the number of communities. */
int cvb_coarsen(cvb_engine *e, int64_t *new_nv_out);

/* distbuildNextLevelGraph across N ranks (rebuild.cpp:27-454) in three device stages; the
 * caller (cuvite_b200/dist.py) moves the contracted pairs between ranks with an NCCL
 * all-to-all in place of the reference's MPI exchange (rebuild.cpp:281-403).
 *
 * 1. cvb_renumber       — distReNumber, rebuild.cpp:72-85 with the ONE-rank numbering (SURVEY
 *    section 8c: the N-rank runs are compared with the 1-rank reference): d_newid[i] = rank,
 *    in order of first appearance scanning i ascending, of d_labels[i].  d_labels is the
 *    all-gathered cvect (DEVICE, int32[n], values in [0,n)); *n_comm_out = communities.
 * 2. cvb_contract_local — fill_newEdgesMap, rebuild.cpp:244-279: every entry (i -> t, w) of
 *    this engine's rows becomes (newid[base+i] -> newid[t], w), parallel entries merged;
 *    the distinct pairs are left ascending in CVB_BUF_PAIR_KEYS / CVB_BUF_PAIR_WEIGHTS.
 * 3. cvb_graph_from_pairs — send_newEdges, rebuild.cpp:405-427: the (DEVICE) pairs routed to
 *    this rank — sources in [base, base + nv_local), any order, duplicates allowed — are
 *    merged, emitted as a CSR sorted by tail and become the engine's graph.  Weights are in
 *    the engine's weight mode (uint32 or double). */
int cvb_renumber(cvb_engine *e, const int32_t *d_labels, int64_t n, int32_t *d_newid, int64_t *n_comm_out);
int cvb_contract_local(cvb_engine *e, const int32_t *d_newid_global, int64_t n_comm, int64_t *n_pairs_out);
int cvb_graph_from_pairs(cvb_engine *e, int64_t nv_local, int64_t base, int64_t nv_global,
                         const uint64_t *d_keys, const void *d_weights, int64_t n_pairs);

/* ---- Vite binary files (distgraph.cpp) --------------------------------------- */

/* Which rows of a Vite file rank `rank` of `nparts` owns: the reference's vertex ranges
 * parts[p] = nv*p / P (loadDistGraphMPIIO, distgraph.cpp:69-203, :118-119) or, with
 * `balanced`, its -b edge-balanced ranges (balanceEdges, distgraph.cpp:22-66 — restated
 * loop for loop, incl. its off-by-one charging).  parts_out (nparts+1, may be NULL) gets
 * all range starts.  Pure host code. */
typedef struct cvb_vite_part {
  int64_t nv_global, ne_global; /* file header                                   */
  int64_t base, nv_local;       /* this rank's rows [base, base + nv_local)       */
  int64_t first_entry, ne_local;/* its slice of the edge list                     */
} cvb_vite_part;
int cvb_vite_partition(const char *path, int rank, int nparts, int balanced, cvb_vite_part *out,
                       int64_t *parts_out);
/* Read those rows: offsets[nv_local+1] rebased to 0 (distgraph.cpp:194-195) and the AoS
 * Edge{int64 tail; double weight} slice (HOST buffers of the caller).  Pure host code. */
int cvb_vite_read_rows(const char *path, const cvb_vite_part *part, int64_t *offsets, void *edges_aos);
/* File -> engine: partition, read into pinned staging, cvb_graph_upload_host.  Replaces
 * loadDistGraphMPIIO[Balanced] + the first gpu_for_louvain_iteration upload. */
int cvb_graph_load_vite(cvb_engine *e, const char *path, int rank, int nparts, int balanced,
                        cvb_vite_part *out);

/* D2H of the engine's current graph in the boundary layout (for writing Vite
 * files / handing the next level back to the reference's DistGraph). */
int cvb_graph_download(cvb_engine *e, int64_t *offsets, int64_t *tails, double *weights);

/* The default-flag phase driver of main.cpp:218-495 on one GPU: phases with
 * threshold 1e-6 until a phase gains no more, coarsening in between.
 * level_q/level_iters (cumulative)/level_nv hold up to max_levels accepted
 * levels; comm_out[nv0] (may be NULL) receives the composed community of
 * every original vertex.  Returns the number of accepted levels (>= 0). */
int cvb_louvain_run(cvb_engine *e, int max_levels, double *level_q, int *level_iters,
                    int64_t *level_nv, int64_t *comm_out);

/* Same with the reference's driver options.  CVB_RUN_THRESHOLD_CYCLING = option -i
 * (main.cpp:225-239, 482-483): the phase threshold (sweep-loop exit and level acceptance)
 * cycles 1e-3 x3, 1e-4 x4, 1e-5 x3, 1e-6 x3.  (The extra 1e-6 pass of main.cpp:436-441 right
 * before exiting changes no reported level and no output, and is not run.) */
enum { CVB_RUN_THRESHOLD_CYCLING = 1 };
int cvb_louvain_run_ex(cvb_engine *e, int max_levels, int flags, double *level_q, int *level_iters,
                       int64_t *level_nv, int64_t *comm_out);

/* Early termination, options -t 1..4 and -a (main.cpp:613-618, 326-341): the phase with per-vertex
 * freezing — louvain.cpp:7-204 for et_type 1 (local) and 3 (stops once 90 vertices are frozen),
 * louvain.cpp:206-423 for et_type 2 and 4 (probabilistic: a vertex whose label did not change
 * multiplies its probability by 1 - et_delta and freezes at p <= 0.02).  The freeze rules are
 * mirrored bug for bug (see setup.cuh, k_et_rotate).  One GPU. */
int cvb_phase_run_et(cvb_engine *e, double lower, double thresh, int et_type, double et_delta,
                     double *q_out, int *iters_out);
/* cvb_louvain_run_ex with every phase run by cvb_phase_run_et (et_type 0 = no early termination). */
int cvb_louvain_run_et(cvb_engine *e, int max_levels, int flags, int et_type, double et_delta,
                       double *level_q, int *level_iters, int64_t *level_nv, int64_t *comm_out);

/* Distance-1 colouring, option -c max_colors (main.cpp:242-282, 600-601): distColoringMultiHashMinMax
 * (coloring.cpp:3-200, nHash = max_colors / 2, 70 % coverage target) followed by
 * distLouvainMethodWithColoring (louvain.cpp:756-949) — inside an iteration the colour buckets are swept
 * one after the other against the community info the earlier buckets left.  *num_colors_out = the
 * reference's "Number of colors (2*nHash)".  max_colors < 2 runs the plain phase, as the reference does.
 * One GPU. */
int cvb_phase_run_color(cvb_engine *e, double lower, double thresh, int max_colors, double *q_out,
                        int *iters_out, int *num_colors_out);
/* cvb_louvain_run_ex with the FIRST phase run by cvb_phase_run_color (the reference colours phase 0 only). */
int cvb_louvain_run_color(cvb_engine *e, int max_levels, int flags, int max_colors, double *level_q,
                          int *level_iters, int64_t *level_nv, int64_t *comm_out);
/* -c together with -t / -a (main.cpp:258-277): the coloured phase with early termination, louvain.cpp:951-1178
 * (et_type 1, 3) and 1180-1426 (et_type 2, 4); et_type 0 = cvb_phase_run_color.  In cvb_louvain_run_color_et the
 * first phase is coloured with early termination and the later ones run cvb_phase_run_et, as main.cpp does. */
int cvb_phase_run_color_et(cvb_engine *e, double lower, double thresh, int max_colors, int et_type, double et_delta,
                           double *q_out, int *iters_out, int *num_colors_out);
int cvb_louvain_run_color_et(cvb_engine *e, int max_levels, int flags, int max_colors, int et_type, double et_delta,
                             double *level_q, int *level_iters, int64_t *level_nv, int64_t *comm_out);
/* distColoringMultiHashMinMax alone (coloring.cpp:3-72): colours the uploaded graph, *num_colors_out = the
 * reference's "Number of colors (2*nHash)"; read the colours with cvb_get_colors.  This is all option -d
 * (vertex ordering, main.cpp:283-324, louvain.cpp:1433-1625) adds on one rank: distLouvainMethodVertexOrder visits
 * the vertices colour by colour but applies the community updates once per iteration, so every vertex still reads
 * the iteration's starting state and the phase equals cvb_phase_run exactly — likewise `-d` with `-t` equals
 * cvb_phase_run_et (tests/test_coloring.py pins this on the reference CLI: `-d 8` prints the plain run's level
 * lines). */
int cvb_color_graph(cvb_engine *e, int max_colors, int *num_colors_out);
/* The colours of the last cvb_phase_run_color / cvb_color_graph (-1 = uncoloured), nv_local values. */
int cvb_get_colors(cvb_engine *e, int32_t *out);

/* ---- instrumentation ------------------------------------------------------ */
/* Kernels launched by this engine since creation (bench.py "gpu_launches"). */
int64_t cvb_launch_count(const cvb_engine *e);
/* Device time of the last cvb_sweep's kernels in ms (CUDA events on the engine
 * stream; valid after a sync).  Returns < 0 when no sweep ran. */
float cvb_last_sweep_ms(cvb_engine *e);

/* ---- N > 1 ranks: the phase over NCCL ------------------------------------------------
 * One engine per GPU / process holding the rows [parts[rank], parts[rank+1]) of the reference's
 * vertex-range partition (distgraph.cpp:118-119, or the -b ranges).  These calls are collective
 * (every rank, same order) and replace distLouvainMethod(..., GpuGraph&) at any rank count —
 * louvain.hpp:93-97, louvain.cpp:591-753 — with its MPI ghost exchange: exchangeVertexReqs
 * (3118-3264), fillRemoteCommunities (2588-2959), updateRemoteCommunities (2983-3116) and the
 * Allreduces of louvain.cpp:2175 and 2522.  NCCL is resolved with dlopen at the first call.
 *
 * cvb_dist_unique_id : rank 0 makes the 128-byte NCCL id; the host program ships it to the other
 *                      ranks by whatever it already has (MPI_Bcast, torch.distributed, a file).
 * cvb_dist_init      : create the communicator on the engine's device.
 * cvb_dist_set_parts : vertex ranges of the uploaded graph (parts[world+1]); NULL = nv*p/P.
 * cvb_dist_agree     : after every upload: the exact-integer weight mode is kept only if ALL ranks
 *                      have integer weights and the GLOBAL total stays below 2^32.
 * cvb_dist_phase_begin / _sweep / _advance : distInitLouvain, one iteration (sweep + label
 *                      all-gather + community rebuild + reduction; *q_out = this iteration's
 *                      modularity, louvain.cpp:2472-2473), label rotation.
 * cvb_dist_phase_run : the whole loop (same contract as cvb_phase_run).
 * cvb_dist_stage_ms  : with CVB_DIST_TIMING=1, mean CUDA-event milliseconds of
 *                      {sweep kernels, label all-gather, label diff + reduce, sums all-reduce, read-back}. */
int cvb_dist_unique_id(void *out, int nbytes);
int cvb_dist_init(cvb_engine *e, const void *id, int nbytes, int rank, int world);
int cvb_dist_finalize(cvb_engine *e);
int cvb_dist_set_parts(cvb_engine *e, const int64_t *parts);
int cvb_dist_agree(cvb_engine *e);
int cvb_dist_phase_begin(cvb_engine *e);
int cvb_dist_sweep(cvb_engine *e, double *q_out);
int cvb_dist_advance(cvb_engine *e);
int cvb_dist_phase_run(cvb_engine *e, double lower, double thresh, int max_iters,
                       double *q_out, int *iters_out);
int cvb_dist_stage_ms(cvb_engine *e, double out[5]);

#ifdef __cplusplus
}
#endif
#endif /* CUVITE_B200_H */
