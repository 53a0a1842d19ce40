#!/usr/bin/env bash
# oracle/build_ref.sh — TEST INFRASTRUCTURE ONLY.
#
# Compiles the reference's own CPU Louvain path (pnnl/cuVite: louvain.cpp,
# rebuild.cpp, distgraph.cpp, utils.cpp, coloring.cpp, compare.cpp, main.cpp)
# from the sources where they lie under $REF (default /root/reference) into
#     oracle/_ref/libcuvite_ref.so   (ref_harness.cpp C entry points)
#     oracle/_ref/gc_cpu64           (the reference CLI, 64-bit ids / double)
#     oracle/_ref/gc_gpu64           (the reference CLI with ITS OWN CUDA path — louvain_cuda.cu,
#                                     modularity.cu, GpuGraph.cu — recompiled for sm_100 with
#                                     -DPRINT_TIMEDS: the "recompiled kernels" baseline that
#                                     tools/ref_cuda_baseline.py times on the B200 box)
# Only binaries land in oracle/_ref/ (git-ignored, NOT gpurun-ignored).  The
# three mechanical accommodations SURVEY.md §8c lists are applied to a scratch
# copy under a temp dir, never to the repo:
#   1. single-rank mpi.h shim (oracle/mpi_shim) — the image has no MPI;
#   2. `default(none)` -> `default(shared)`: g++ >= 9 rejects the reference's
#      OpenMP regions that use const locals without listing them;
#   3. a forward declaration of CuVite::GpuGraph, which louvain.cpp:2483 names
#      outside #ifdef RUNONGPU.
# The reference's own build system (CMake + mpicxx, broken as shipped) is not run.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${REF:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$REF" ]; then
  echo "build_ref.sh: $REF absent (GPU box?) — keeping prebuilt $OUT" >&2
  exit 0
fi
mkdir -p "$OUT"
TMP="$(mktemp -d /tmp/cuvite_ref_build.XXXXXX)"
trap 'rm -rf "$TMP"' EXIT
for f in main.cpp coloring.cpp compare.cpp distgraph.cpp louvain.cpp rebuild.cpp utils.cpp \
         coloring.hpp compare.hpp distgraph.hpp edge.hpp graph.hpp locks.hpp louvain.hpp \
         rebuild.hpp utils.hpp; do
  sed 's/default(none)/default(shared)/g' "$REF/$f" > "$TMP/$f"
done
cat > "$TMP/fwd.h" <<'EOF'
namespace CuVite { class GpuGraph; }
using namespace CuVite;
EOF
CXX="${REF_CXX:-/usr/bin/g++}"  # NB: $CXX in this image points at a g++ without libgomp.spec
FLAGS="-std=c++11 -O3 -fopenmp -fpermissive -w -ffp-contract=off -DDONT_CREATE_DIAG_FILES -include $TMP/fwd.h -I$HERE/mpi_shim -I$TMP"
# reference CLI
$CXX $FLAGS "$TMP"/{main,coloring,compare,distgraph,louvain,rebuild,utils}.cpp -o "$OUT/gc_cpu64"
# harness library: ref_harness.cpp #includes louvain.cpp to reach its statics
$CXX $FLAGS -fPIC -shared "$HERE/ref_harness.cpp" "$TMP"/{coloring,distgraph,rebuild,utils}.cpp \
     -o "$OUT/libcuvite_ref.so"
# drop-in proof: the reference's UNMODIFIED main.cpp / louvain.cpp (-DRUNONGPU=1) linked
# against include/GpuGraph.cuh + cuvite_b200/csrc/adapter.cpp + libcuvite_b200.so
# instead of its own louvain_cuda.cu / modularity.cu / GpuGraph.cu
REPO="$(cd "$HERE/.." && pwd)"
if [ -f "$REPO/cuvite_b200/libcuvite_b200.so" ]; then
  $CXX $FLAGS -DRUNONGPU=1 -I"$REPO/include" \
       "$TMP"/{main,coloring,compare,distgraph,louvain,rebuild,utils}.cpp "$REPO/cuvite_b200/csrc/adapter.cpp" \
       -L"$REPO/cuvite_b200" -lcuvite_b200 -Wl,-rpath,'$ORIGIN/../../cuvite_b200' \
       -Wl,-rpath,/usr/local/cuda/lib64 -o "$OUT/gc_gpu_dropin"
  echo "built $OUT/gc_gpu_dropin"
  # outer-boundary drop-in: the GPU overload of distLouvainMethod comes from adapter_resident.cpp
  # (resident engine, one upload per phase), so louvain.cpp is compiled WITHOUT -DRUNONGPU and
  # main.cpp WITH it
  OBJ="$TMP/obj_resident"; mkdir -p "$OBJ"
  for f in coloring compare distgraph louvain rebuild utils; do
    $CXX $FLAGS -c "$TMP/$f.cpp" -o "$OBJ/$f.o"
  done
  $CXX $FLAGS -DRUNONGPU=1 -I"$REPO/include" -c "$TMP/main.cpp" -o "$OBJ/main.o"
  $CXX $FLAGS -DRUNONGPU=1 -I"$REPO/include" -c "$REPO/cuvite_b200/csrc/adapter.cpp" -o "$OBJ/adapter.o"
  $CXX $FLAGS -DRUNONGPU=1 -I"$REPO/include" -c "$REPO/cuvite_b200/csrc/adapter_resident.cpp" -o "$OBJ/adapter_resident.o"
  # adapter.o also defines gpu_for_louvain_iteration / gpu_reduce_mod: unused here, harmless
  $CXX -fopenmp "$OBJ"/{main,coloring,compare,distgraph,louvain,rebuild,utils,adapter,adapter_resident}.o \
       -L"$REPO/cuvite_b200" -lcuvite_b200 -Wl,-rpath,'$ORIGIN/../../cuvite_b200' \
       -Wl,-rpath,/usr/local/cuda/lib64 -o "$OUT/gc_gpu_resident"
  echo "built $OUT/gc_gpu_resident"
fi
# the reference's own CUDA path, recompiled for Blackwell (baseline only; SURVEY.md §8c "also compilable")
if command -v nvcc >/dev/null 2>&1; then
  for f in louvain_cuda.cu modularity.cu GpuGraph.cu GpuGraph.cuh louvain_cuda.cuh louvain_cuda_constants.cuh \
           louvain_cuda_cpp_interface.hpp; do
    sed 's/default(none)/default(shared)/g' "$REF/$f" > "$TMP/$f"
  done
  nvcc -ccbin "$CXX" -std=c++14 -O3 -gencode arch=compute_100,code=sm_100 -w \
       -Xcompiler -fopenmp,-fpermissive,-w -DRUNONGPU=1 -DPRINT_TIMEDS -DDONT_CREATE_DIAG_FILES \
       -I"$HERE/mpi_shim" -I"$TMP" \
       "$TMP"/{main,coloring,compare,distgraph,louvain,rebuild,utils}.cpp \
       "$TMP"/{louvain_cuda,modularity,GpuGraph}.cu -lcudart -lgomp -o "$OUT/gc_gpu64"
  echo "built $OUT/gc_gpu64"
fi
echo "built $OUT/gc_cpu64 $OUT/libcuvite_ref.so"
