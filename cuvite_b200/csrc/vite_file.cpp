// vite_file.cpp — the Vite binary graph format on the host side of the C ABI: which rows a rank owns, and
// those rows read straight into (pinned) staging and uploaded — the reference's loadDistGraphMPIIO
// (distgraph.cpp:69-203) and loadDistGraphMPIIOBalanced (205-337) without MPI-IO.
//
// Layout (64-bit build, little endian, no padding; distgraph.cpp:99-100,131,170):
//   int64 nv; int64 ne; int64 offsets[nv + 1]; struct { int64 tail; double weight; } edges[ne];
// ne counts directed CSR entries.
#include <sys/stat.h>

#include <cstdio>
#include <cstdlib>
#include <new>
#include <vector>

#include "common.cuh"

namespace {

struct File {
  FILE *f = nullptr;
  ~File() {
    if (f) std::fclose(f);
  }
};

bool read_at(FILE *f, int64_t byte_offset, void *dst, size_t bytes) {
  if (fseeko(f, static_cast<off_t>(byte_offset), SEEK_SET) != 0) return false;
  return bytes == 0 || std::fread(dst, 1, bytes, f) == bytes;
}

// balanceEdges, distgraph.cpp:22-66: one pass over the first nv offsets; bin p takes vertices until the entries
// it has SEEN reach ne / P — vertex m is charged offsets[m] - offsets[m-1], i.e. the row of vertex m-1, because
// the loop reads offsets[m] — and the vertex that fills the bin already belongs to bin p+1.
void balanced_parts(const std::vector<int64_t> &seen, int64_t nv, int64_t ne, int nparts, int64_t *parts) {
  const int64_t cap = ne / nparts;
  std::vector<int64_t> bins(nparts, 0), count(nparts + 1, 0);
  int64_t past = 0;
  int p = 0;
  for (int64_t m = 0; m < nv; ++m) {
    const int64_t idx = seen[m];
    if (bins[p] < cap || p == nparts - 1) bins[p] += idx - past;
    if (bins[p] >= cap && p < nparts - 1) ++p;
    ++count[p + 1];
    past = idx;
  }
  parts[0] = 0;
  for (int k = 1; k <= nparts; ++k) parts[k] = parts[k - 1] + count[k];
}

}  // namespace

extern "C" {

int cvb_vite_partition(const char *path, int rank, int nparts, int balanced, cvb_vite_part *out, int64_t *parts_out) {
  CVB_CHECK(path && out && nparts >= 1 && rank >= 0 && rank < nparts, CVB_ERR_ARG, "cvb_vite_partition: bad argument");
  File fh;
  fh.f = std::fopen(path, "rb");
  CVB_CHECK(fh.f, CVB_ERR_ARG, "cvb_vite_partition: cannot open %s", path);
  int64_t hdr[2];
  CVB_CHECK(read_at(fh.f, 0, hdr, sizeof hdr) && hdr[0] >= 0 && hdr[1] >= 0, CVB_ERR_ARG,
            "cvb_vite_partition: %s is not a Vite file (short header)", path);
  const int64_t nv = hdr[0], ne = hdr[1];
  // the header must describe the file: a corrupt nv / ne would otherwise size host allocations
  struct stat sb;
  CVB_CHECK(fstat(fileno(fh.f), &sb) == 0, CVB_ERR_ARG, "cvb_vite_partition: cannot stat %s", path);
  const __int128 need = static_cast<__int128>(16) + (static_cast<__int128>(nv) + 1) * 8 + static_cast<__int128>(ne) * 16;
  CVB_CHECK(need <= static_cast<__int128>(sb.st_size), CVB_ERR_ARG,
            "cvb_vite_partition: %s: header says nv=%lld ne=%lld but the file has only %lld bytes", path,
            (long long)nv, (long long)ne, (long long)sb.st_size);
  std::vector<int64_t> parts(nparts + 1);
  if (balanced) {  // option -b
    std::vector<int64_t> seen;
    try {
      seen.resize(nv > 0 ? nv : 1);
    } catch (const std::bad_alloc &) {
      CVB_CHECK(false, CVB_ERR_ARG, "cvb_vite_partition: %s: cannot allocate %lld offsets", path, (long long)nv);
    }
    CVB_CHECK(read_at(fh.f, 16, seen.data(), static_cast<size_t>(nv) * 8), CVB_ERR_ARG,
              "cvb_vite_partition: %s: truncated offsets", path);
    balanced_parts(seen, nv, ne, nparts, parts.data());
  } else {  // parts[p] = nv*p / P, distgraph.cpp:118-119
    for (int p = 0; p <= nparts; ++p) parts[p] = static_cast<int64_t>((static_cast<__int128>(nv) * p) / nparts);
  }
  int64_t ends[2];
  CVB_CHECK(read_at(fh.f, 16 + parts[rank] * 8, &ends[0], 8) && read_at(fh.f, 16 + parts[rank + 1] * 8, &ends[1], 8) &&
                ends[0] >= 0 && ends[1] >= ends[0] && ends[1] <= ne,
            CVB_ERR_ARG, "cvb_vite_partition: %s: bad offsets", path);
  out->nv_global = nv;
  out->ne_global = ne;
  out->base = parts[rank];
  out->nv_local = parts[rank + 1] - parts[rank];
  out->first_entry = ends[0];
  out->ne_local = ends[1] - ends[0];
  if (parts_out)
    for (int p = 0; p <= nparts; ++p) parts_out[p] = parts[p];
  return CVB_OK;
}

int cvb_vite_read_rows(const char *path, const cvb_vite_part *part, int64_t *offsets, void *edges_aos) {
  CVB_CHECK(path && part && offsets && (edges_aos || part->ne_local == 0), CVB_ERR_ARG, "cvb_vite_read_rows: bad argument");
  File fh;
  fh.f = std::fopen(path, "rb");
  CVB_CHECK(fh.f, CVB_ERR_ARG, "cvb_vite_read_rows: cannot open %s", path);
  CVB_CHECK(read_at(fh.f, 16 + part->base * 8, offsets, static_cast<size_t>(part->nv_local + 1) * 8), CVB_ERR_ARG,
            "cvb_vite_read_rows: %s: truncated offsets", path);
  for (int64_t i = part->nv_local; i >= 0; --i) offsets[i] -= offsets[0];  // rebased to 0, distgraph.cpp:194-195
  CVB_CHECK(read_at(fh.f, 16 + (part->nv_global + 1) * 8 + part->first_entry * 16, edges_aos,
                    static_cast<size_t>(part->ne_local) * 16),
            CVB_ERR_ARG, "cvb_vite_read_rows: %s: truncated edge list", path);
  return CVB_OK;
}

int cvb_graph_load_vite(cvb_engine *e, const char *path, int rank, int nparts, int balanced, cvb_vite_part *out) {
  CVB_CHECK(e, CVB_ERR_ARG, "null engine");
  cvb_vite_part part;
  CVB_TRY(cvb_vite_partition(path, rank, nparts, balanced, &part, nullptr));
  // pinned staging: the file is read once, the H2D copy runs at PCIe speed
  int64_t *h_off = nullptr;
  void *h_edges = nullptr;
  CVB_CUDA(cudaMallocHost(reinterpret_cast<void **>(&h_off), static_cast<size_t>(part.nv_local + 1) * 8));
  cudaError_t err = cudaMallocHost(&h_edges, static_cast<size_t>(part.ne_local > 0 ? part.ne_local : 1) * 16);
  int rc = CVB_OK;
  if (err != cudaSuccess) {
    cvb::set_error("cvb_graph_load_vite: cudaMallocHost failed: %s", cudaGetErrorString(err));
    rc = CVB_ERR_CUDA;
  }
  if (rc == CVB_OK) rc = cvb_vite_read_rows(path, &part, h_off, h_edges);
  if (rc == CVB_OK)
    rc = cvb_graph_upload_host(e, part.nv_local, part.base, part.nv_global, h_off, h_edges, nullptr, nullptr);
  if (rc == CVB_OK) rc = cvb_sync(e);  // the staging buffers are freed below
  if (h_edges) cudaFreeHost(h_edges);
  cudaFreeHost(h_off);
  if (rc == CVB_OK && out) *out = part;
  return rc;
}

}  // extern "C"
