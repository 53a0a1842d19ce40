"""Full multi-level run of one R-MAT graph on one GPU (no CPU checker) — the target of an ncu launch list
that separates the coarsening kernels (CUB sort passes, renumber, row emit, post-upload setup) from the sweeps:
  ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file out.csv \
      python tools/coarsen_profile.py --scale 22
Prints one JSON line with the host-clock seconds of the run."""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=int, default=22)
    ap.add_argument("--runs", type=int, default=2)
    a = ap.parse_args()
    from cuvite_b200.engine import GpuGraph
    from cuvite_b200.graphgen import rmat
    g = rmat(a.scale, seed=2, device="cuda")
    torch.cuda.synchronize()
    gpu = GpuGraph()
    secs = []
    for _ in range(a.runs):
        gpu.upload_device(g.offsets, g.tails, g.weights)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        lq, li, ln, _ = gpu.run()
        secs.append(time.perf_counter() - t0)
    gpu.close()
    print(json.dumps({"scale": a.scale, "csr_entries": g.ne, "seconds_per_run": secs, "level_q": lq, "level_iters": li,
                      "level_nv": ln}))


if __name__ == "__main__":
    main()
