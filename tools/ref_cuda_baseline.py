"""Second baseline (SURVEY.md §8d "optional second baseline"): the reference's OWN CUDA path
(louvain_cuda.cu / modularity.cu / GpuGraph.cu recompiled for sm_100 with -DPRINT_TIMEDS into
oracle/_ref/gc_gpu64 by oracle/build_ref.sh) run on this box's GPU on the same Vite file as

  * oracle/_ref/gc_gpu_dropin — the same reference main.cpp/louvain.cpp linked against
    libcuvite_b200.so through the per-iteration plugin boundary,
  * oracle/_ref/gc_gpu_resident — the same, with the GPU overload of distLouvainMethod supplied
    by adapter_resident.cpp (one upload per phase), oracle/_ref/gc_cpu64 (the CPU reference), and
  * the resident engine (cvb_louvain_run).

Test/measurement infrastructure: prints one JSON object; nothing here is on the product path.
Per-iteration kernel time of the reference = its GPU_kernels stamp minus its GPU_iter1 stamp
(louvain_cuda.cu:2527-2683, both after cudaDeviceSynchronize), i.e. kernels only, no copies.
"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run_cli(exe, path, threads):
    t0 = time.time()
    r = subprocess.run([exe, "-f", path], capture_output=True, text=True, timeout=1500,
                       cwd=os.path.dirname(path), env=dict(os.environ, OMP_NUM_THREADS=str(threads)))
    wall = time.time() - t0
    out = r.stdout
    levels = [(int(a), float(b), float(c), int(d)) for a, b, c, d in re.findall(
        r"Level (\d+), Modularity: ([0-9.eE+-]+), Clustering time: ([0-9.eE+-]+), Iterations: (\d+)", out)]
    stamps = {}
    for name, val in re.findall(r"Time (GPU_\w+): ([0-9.eE+-]+)", out):
        stamps.setdefault(name, []).append(float(val))
    res = {"rc": r.returncode, "wall_s": wall, "levels": levels}
    phases = re.findall(r"cuvite_b200 resident phase: nv=(\d+) entries=(\d+) upload ([0-9.eE+-]+) s, phase ([0-9.eE+-]+) s "
                        r"\((\d+) iterations\), labels ([0-9.eE+-]+) s", out)
    if phases:  # CVB_ADAPTER_TIMING=1: the resident adapter's own split per phase
        res["resident_phases"] = [{"nv": int(a), "entries": int(b), "upload_s": float(c), "phase_s": float(d),
                                   "iterations": int(i), "labels_s": float(f)} for a, b, c, d, i, f in phases]
    if "GPU_kernels" in stamps and "GPU_iter1" in stamps:
        k = [a - b for a, b in zip(stamps["GPU_kernels"], stamps["GPU_iter1"])]
        res["kernel_s_per_iteration"] = k
        res["call_s_per_iteration"] = stamps.get("GPU_locupd", [])
    if r.returncode != 0:
        res["tail"] = (out[-1500:] + r.stderr[-1500:])
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=int, default=20)
    ap.add_argument("--threads", type=int, default=os.cpu_count())
    ap.add_argument("--only", default="", help="comma-separated subset of the CLI binaries to run")
    a = ap.parse_args()
    import numpy as np
    import torch
    from cuvite_b200 import graphgen as gg, vite_io
    from cuvite_b200.engine import GpuGraph
    g = gg.rmat(a.scale, seed=2, device="cuda").cpu()
    ne = int(g.offsets[-1])
    ref = os.path.join(ROOT, "oracle", "_ref")
    out = {"workload": f"R-MAT scale-{a.scale}", "nv": int(g.nv), "entries": ne, "host_threads": a.threads}
    with tempfile.TemporaryDirectory() as d:
        f = os.path.join(d, "g.bin")
        vite_io.write_vite(f, g)
        for tag in ("gc_gpu64", "gc_gpu_dropin", "gc_gpu_resident", "gc_cpu64"):
            if a.only and tag not in a.only.split(","):
                continue
            exe = os.path.join(ref, tag)
            out[tag] = run_cli(exe, f, a.threads) if os.path.exists(exe) else {"unavailable": True}
    ks = out.get("gc_gpu64", {}).get("kernel_s_per_iteration")
    if ks:
        n0 = out["gc_gpu64"]["levels"][0][3] if out["gc_gpu64"]["levels"] else len(ks)
        lvl0 = ks[:n0]  # level 0 runs on the full graph
        out["ref_cuda_level0_sweep_ms"] = [round(x * 1e3, 3) for x in lvl0]
        out["ref_cuda_level0_undirected_edges_per_s"] = ne / 2 / (sum(lvl0) / len(lvl0))
    gpu = GpuGraph()
    o, t, w = g.offsets.numpy(), g.tails.numpy(), g.weights.numpy()
    gpu.upload(o, t, w)
    torch.cuda.synchronize()
    t0 = time.time()
    lq, li, ln, _ = gpu.run()
    out["resident"] = {"wall_s": time.time() - t0, "level_q": [float(x) for x in lq], "level_iters": li,
                       "level_communities": ln}
    gpu.close()
    print(json.dumps(out, default=lambda x: x.tolist() if isinstance(x, np.ndarray) else str(x)))


if __name__ == "__main__":
    main()
