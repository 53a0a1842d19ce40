/* Single-rank MPI shim — TEST INFRASTRUCTURE ONLY (part of oracle/).
 *
 * The reference (pnnl/cuVite) is an MPI program and this image ships no MPI.
 * The shim lets the reference's *unmodified* CPU sources (louvain.cpp,
 * rebuild.cpp, distgraph.cpp, main.cpp ...) compile and run as exactly one
 * rank so that they can serve as the parity oracle.
 *
 * Semantics with one rank: every collective is a self-copy, point-to-point
 * traffic never happens (the reference guards all Isend/Irecv with
 * `i != me`; Sendrecv is only ever issued to MPI_PROC_NULL), and MPI-IO maps
 * onto stdio.  Anything that would need a second rank aborts loudly.
 *
 * Nothing in the product (cuvite_b200/, include/) includes this header.
 */
#ifndef CUVITE_B200_ORACLE_MPI_SHIM_H
#define CUVITE_B200_ORACLE_MPI_SHIM_H

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

typedef int MPI_Comm;
typedef int MPI_Datatype; /* value = extent in bytes */
typedef int MPI_Op;
typedef int MPI_Request;
typedef int MPI_Info;
typedef long MPI_Aint;
typedef long long MPI_Offset;
typedef FILE *MPI_File;
struct MPI_Status { int MPI_SOURCE, MPI_TAG, MPI_ERROR; };

#define MPI_COMM_WORLD 0
#define MPI_SUCCESS 0
#define MPI_REQUEST_NULL (-1)
#define MPI_STATUS_IGNORE ((MPI_Status *)0)
#define MPI_STATUSES_IGNORE ((MPI_Status *)0)
#define MPI_IN_PLACE ((void *)-1)
#define MPI_PROC_NULL (-2)
#define MPI_INFO_NULL 0
#define MPI_UNDEFINED (-32766)
#define MPI_COMM_TYPE_SHARED 1
#define MPI_MODE_RDONLY 1
#define MPI_MODE_WRONLY 2
#define MPI_MODE_CREATE 4

#define MPI_BYTE 1
#define MPI_INT 4
#define MPI_FLOAT 4
#define MPI_INT32_T 4
#define MPI_LONG 8
#define MPI_DOUBLE 8
#define MPI_INT64_T 8

#define MPI_SUM 1
#define MPI_MIN 2
#define MPI_MAX 3

static inline void shim_need_peer_(const char *what) {
  std::fprintf(stderr, "mpi shim: %s needs a second rank\n", what);
  std::abort();
}
static inline void shim_copy_(const void *s, void *r, long n, MPI_Datatype t) {
  if (s != MPI_IN_PLACE && s != r && n > 0) std::memcpy(r, s, (size_t)n * (size_t)t);
}

static inline int MPI_Init(int *, char ***) { return MPI_SUCCESS; }
static inline int MPI_Finalize() { return MPI_SUCCESS; }
static inline int MPI_Comm_size(MPI_Comm, int *n) { *n = 1; return MPI_SUCCESS; }
static inline int MPI_Comm_rank(MPI_Comm, int *r) { *r = 0; return MPI_SUCCESS; }
static inline int MPI_Comm_split_type(MPI_Comm, int, int, MPI_Info, MPI_Comm *c) { *c = 0; return MPI_SUCCESS; }
static inline int MPI_Comm_free(MPI_Comm *) { return MPI_SUCCESS; }
static inline int MPI_Barrier(MPI_Comm) { return MPI_SUCCESS; }
static inline int MPI_Abort(MPI_Comm, int code) { std::fprintf(stderr, "MPI_Abort(%d)\n", code); std::exit(code ? code : 1); }
static inline double MPI_Wtime() {
  using namespace std::chrono;
  return duration<double>(steady_clock::now().time_since_epoch()).count();
}

static inline int MPI_Allreduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op, MPI_Comm) { shim_copy_(s, r, n, t); return MPI_SUCCESS; }
static inline int MPI_Reduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op, int, MPI_Comm) { shim_copy_(s, r, n, t); return MPI_SUCCESS; }
static inline int MPI_Bcast(void *, int, MPI_Datatype, int, MPI_Comm) { return MPI_SUCCESS; }
static inline int MPI_Alltoall(const void *s, int sn, MPI_Datatype st, void *r, int, MPI_Datatype, MPI_Comm) { shim_copy_(s, r, sn, st); return MPI_SUCCESS; }
static inline int MPI_Ialltoall(const void *s, int sn, MPI_Datatype st, void *r, int, MPI_Datatype, MPI_Comm, MPI_Request *q) { shim_copy_(s, r, sn, st); *q = MPI_REQUEST_NULL; return MPI_SUCCESS; }
static inline int MPI_Alltoallv(const void *s, const int *sc, const int *sd, MPI_Datatype st, void *r, const int *, const int *rd, MPI_Datatype, MPI_Comm) {
  shim_copy_((const char *)s + (size_t)sd[0] * st, (char *)r + (size_t)rd[0] * st, sc[0], st);
  return MPI_SUCCESS;
}
static inline int MPI_Allgather(const void *s, int sn, MPI_Datatype st, void *r, int, MPI_Datatype, MPI_Comm) { shim_copy_(s, r, sn, st); return MPI_SUCCESS; }
static inline int MPI_Gather(const void *s, int sn, MPI_Datatype st, void *r, int, MPI_Datatype, int, MPI_Comm) { shim_copy_(s, r, sn, st); return MPI_SUCCESS; }
static inline int MPI_Gatherv(const void *s, int sn, MPI_Datatype st, void *r, const int *, const int *rd, MPI_Datatype, int, MPI_Comm) {
  shim_copy_(s, (char *)r + (size_t)rd[0] * st, sn, st);
  return MPI_SUCCESS;
}

static inline int MPI_Isend(const void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request *) { shim_need_peer_("MPI_Isend"); return 1; }
static inline int MPI_Irecv(void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request *) { shim_need_peer_("MPI_Irecv"); return 1; }
static inline int MPI_Sendrecv(const void *, int, MPI_Datatype, int dst, int, void *, int, MPI_Datatype, int src, int, MPI_Comm, MPI_Status *) {
  if (dst != MPI_PROC_NULL || src != MPI_PROC_NULL) shim_need_peer_("MPI_Sendrecv");
  return MPI_SUCCESS;
}
static inline int MPI_Wait(MPI_Request *, MPI_Status *) { return MPI_SUCCESS; }
static inline int MPI_Waitall(int, MPI_Request *, MPI_Status *) { return MPI_SUCCESS; }
static inline int MPI_Test(MPI_Request *, int *flag, MPI_Status *) { *flag = 1; return MPI_SUCCESS; }
static inline int MPI_Testany(int, MPI_Request *, int *idx, int *flag, MPI_Status *) { *idx = MPI_UNDEFINED; *flag = 1; return MPI_SUCCESS; }

static inline int MPI_Get_address(const void *p, MPI_Aint *a) { *a = (MPI_Aint)(intptr_t)p; return MPI_SUCCESS; }
static inline int MPI_Type_create_struct(int n, const int *bl, const MPI_Aint *d, const MPI_Datatype *t, MPI_Datatype *nt) {
  long ext = 0;
  for (int i = 0; i < n; i++) { long e = d[i] + (long)bl[i] * t[i]; if (e > ext) ext = e; }
  *nt = (int)((ext + 7) / 8 * 8);
  return MPI_SUCCESS;
}
static inline int MPI_Type_commit(MPI_Datatype *) { return MPI_SUCCESS; }
static inline int MPI_Type_free(MPI_Datatype *) { return MPI_SUCCESS; }
static inline int MPI_Info_create(MPI_Info *i) { *i = 0; return MPI_SUCCESS; }
static inline int MPI_Info_set(MPI_Info, const char *, const char *) { return MPI_SUCCESS; }
static inline int MPI_Info_free(MPI_Info *) { return MPI_SUCCESS; }

static inline int MPI_File_open(MPI_Comm, const char *name, int mode, MPI_Info, MPI_File *fh) {
  *fh = std::fopen(name, (mode & MPI_MODE_RDONLY) ? "rb" : "wb+");
  return *fh ? MPI_SUCCESS : 1;
}
static inline int MPI_File_close(MPI_File *fh) { if (*fh) std::fclose(*fh); *fh = 0; return MPI_SUCCESS; }
static inline int MPI_File_read_all(MPI_File fh, void *buf, int n, MPI_Datatype t, MPI_Status *) {
  return std::fread(buf, (size_t)t, (size_t)n, fh) == (size_t)n ? MPI_SUCCESS : 1;
}
static inline int MPI_File_read_at(MPI_File fh, MPI_Offset off, void *buf, int n, MPI_Datatype t, MPI_Status *) {
  if (fseeko(fh, (off_t)off, SEEK_SET) != 0) return 1;
  return std::fread(buf, (size_t)t, (size_t)n, fh) == (size_t)n ? MPI_SUCCESS : 1;
}
static inline int MPI_File_write_at(MPI_File fh, MPI_Offset off, const void *buf, int n, MPI_Datatype t, MPI_Status *) {
  if (fseeko(fh, (off_t)off, SEEK_SET) != 0) return 1;
  return std::fwrite(buf, (size_t)t, (size_t)n, fh) == (size_t)n ? MPI_SUCCESS : 1;
}

#endif
