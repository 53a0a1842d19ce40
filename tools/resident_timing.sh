#!/usr/bin/env bash
# Diagnostics: per-phase upload / device / read-back split of the resident outer-boundary drop-in
# (oracle/_ref/gc_gpu_resident, CVB_ADAPTER_TIMING=1) on an R-MAT Vite file.
set -e
SCALE=${1:-20}
D=$(mktemp -d)
python - <<PY
import sys; sys.path.insert(0, ".")
from cuvite_b200 import graphgen as gg, vite_io
vite_io.write_vite("$D/g.bin", gg.rmat($SCALE, seed=2, device="cuda").cpu())
PY
(cd $D && CVB_ADAPTER_TIMING=1 OMP_NUM_THREADS=$(nproc) /root/repo/oracle/_ref/gc_gpu_resident -f $D/g.bin) | grep -E "resident phase|Level|Rebuild|Average"
rm -rf $D
