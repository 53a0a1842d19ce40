"""BASELINE.json config C5: power-law (alpha = 1.8) graph, 2^25 vertices, 512 M sampled undirected edges,
N GPUs with the reference's edge-balanced vertex ranges (-b, balanceEdges distgraph.cpp:22-66).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
      --master-port 29579 tools/run_c5.py [--log2nv 25 --edges 536870912] [--phase-only]

Every rank generates the same deterministic graph on its GPU, keeps its balanced row range and runs the
distributed driver (phase 0 sweeps timed per sweep; then, unless --phase-only, the whole multi-level run).
Rank 0 then runs the one-GPU engine on the whole graph and compares (Q bit for bit, iterations, labels).
Prints one JSON object on rank 0.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2nv", type=int, default=25)
    ap.add_argument("--edges", type=int, default=512 * 1024 * 1024)
    ap.add_argument("--phase-only", action="store_true")
    ap.add_argument("--no-check", action="store_true")
    ap.add_argument("--warmup", action="store_true", help="one untimed full run first (NCCL point-to-point channels, allocator)")
    a = ap.parse_args()
    from cuvite_b200.dist import DistLouvain, init_process_group, load_partition
    from cuvite_b200.engine import CLASS_NAMES, GpuGraph, set_gpuDevices
    from cuvite_b200.graphgen import power_law
    from cuvite_b200.vite_io import balanced_ranges, partition_ranges
    rank, local, world = init_process_group("nccl")
    set_gpuDevices(local)
    torch.cuda.set_device(local)
    nv = 1 << a.log2nv
    g = power_law(nv, a.edges, alpha=1.8, max_degree=1 << 20, seed=4, device="cuda")
    torch.cuda.synchronize()
    ne = g.ne
    parts = balanced_ranges(g.offsets.cpu().numpy(), world)  # the reference's -b ranges
    plain = partition_ranges(nv, world)
    off = g.offsets
    ent_bal = [int(off[parts[p + 1]] - off[parts[p]]) for p in range(world)]
    ent_plain = [int(off[plain[p + 1]] - off[plain[p]]) for p in range(world)]
    side = torch.cuda.Stream()
    torch.cuda.set_stream(side)
    gpu = GpuGraph(stream=side.cuda_stream)
    load_partition(gpu, g, rank, world, parts)
    info = gpu.info()
    classes = torch.tensor([int(info.class_count[i]) for i in range(info.num_classes)], dtype=torch.int64, device="cuda")
    all_classes = [torch.zeros_like(classes) for _ in range(world)]
    dist.all_gather(all_classes, classes)
    dl = DistLouvain(gpu, rank, world, nv, parts=parts)
    # phase 0, sweep by sweep
    dl.phase_begin()
    prev, qs, ms = -1.0, [], []
    for _ in range(32):
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        q = dl.sweep()
        torch.cuda.synchronize()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms.append(1e3 * float(t.item()))
        qs.append(q)
        if q - prev < 1e-6:
            break
        prev = q
        dl.advance()
    phase_q, phase_iters = prev, len(qs)
    out = {"workload": f"power-law alpha=1.8, nv=2^{a.log2nv}, {a.edges} sampled undirected edges (C5), phase 0",
           "n_gpus": world, "nv": nv, "csr_entries": ne, "max_degree_rank0": int(info.max_degree),
           "ranges": "edge-balanced (-b, balanceEdges)", "entries_per_rank_balanced": ent_bal,
           "entries_per_rank_plain_ranges": ent_plain,
           "imbalance_balanced": max(ent_bal) / (sum(ent_bal) / world), "imbalance_plain": max(ent_plain) / (sum(ent_plain) / world),
           "class_counts_per_rank": {CLASS_NAMES[i]: [int(c[i]) for c in all_classes] for i in range(info.num_classes)},
           "q_per_iteration": qs, "phase_q": phase_q, "phase_iters": phase_iters, "sweep_ms_max_rank": ms,
           "edges_per_s_first_sweep": (ne / 2) / (ms[0] * 1e-3),
           "edges_per_s_later_sweeps": (ne / 2) / (np.mean(ms[1:]) * 1e-3) if len(ms) > 1 else None}
    if not a.phase_only:
        if a.warmup:
            load_partition(gpu, g, rank, world, parts)
            DistLouvain(gpu, rank, world, nv, parts=parts).run()
        load_partition(gpu, g, rank, world, parts)
        dl = DistLouvain(gpu, rank, world, nv, parts=parts)
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        lq, li, ln, comm = dl.run()
        torch.cuda.synchronize()
        out.update(full_run_seconds=time.perf_counter() - t0, full_run_stage_seconds=dl.seconds, warmup=bool(a.warmup),
                   level_q=lq, level_iters=li, level_nv=ln)
    gpu.close()
    if rank == 0 and not a.no_check:
        one = GpuGraph(stream=side.cuda_stream)
        one.upload_device(g.offsets, g.tails, g.weights)
        q1, i1 = one.phase_run()
        out["one_gpu_phase_q"], out["one_gpu_phase_iters"] = q1, i1
        out["phase_equal_one_gpu"] = bool(q1 == phase_q and i1 == phase_iters)
        if not a.phase_only:
            if a.warmup:
                one.upload_device(g.offsets, g.tails, g.weights)
                one.run()
            one.upload_device(g.offsets, g.tails, g.weights)
            t0 = time.perf_counter()
            lq1, li1, ln1, c1 = one.run()
            out["one_gpu_full_run_seconds"] = time.perf_counter() - t0
            out["levels_equal_one_gpu"] = bool(lq == lq1 and li == li1 and ln == ln1)
            out["assignment_equal_one_gpu"] = bool(np.array_equal(comm.cpu().numpy(), c1))
        one.close()
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
