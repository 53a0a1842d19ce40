"""Full multi-phase Louvain (phases + coarsening) across N GPUs, checked against the one-GPU
resident driver on the same graph.  Launch with torchrun:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port 29577 tools/run_dist_full.py --scale 22 [--permute]

Every rank generates the same deterministic R-MAT graph on its GPU, uploads its vertex range
(parts[p] = nv*p // P) and runs DistLouvain.run(); rank 0 then runs cvb_louvain_run on the whole
graph on its own GPU and compares levels (Q bit for bit, cumulative iterations, vertices per
level) and the final assignment.  Prints one JSON object on rank 0.
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
    ap.add_argument("--scale", type=int, default=20)
    ap.add_argument("--permute", action="store_true")
    ap.add_argument("--no-check", action="store_true")
    ap.add_argument("--no-warmup", action="store_true", help="skip the untimed warm-up runs (large scales)")
    a = ap.parse_args()
    from cuvite_b200.dist import DistLouvain, init_process_group, load_partition
    from cuvite_b200.engine import GpuGraph, set_gpuDevices
    from cuvite_b200.graphgen import rmat
    rank, local, world = init_process_group("nccl")
    set_gpuDevices(local)
    torch.cuda.set_device(local)
    g = rmat(a.scale, seed=2, device="cuda", permute_seed=3 if a.permute else None)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    torch.cuda.set_stream(side)
    gpu = GpuGraph(stream=side.cuda_stream)
    load_partition(gpu, g, rank, world)
    nv, ne = g.nv, g.ne
    dl = DistLouvain(gpu, rank, world, nv)
    if not a.no_warmup:
        dl.run()  # warm-up (NCCL channels, allocator) on the same input
        load_partition(gpu, g, rank, world)
        dl = DistLouvain(gpu, rank, world, nv)
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    lq, li, ln, comm = dl.run()
    torch.cuda.synchronize()
    secs = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    dist.all_reduce(secs, op=dist.ReduceOp.MAX)
    gpu.close()
    out = {"workload": f"R-MAT scale-{a.scale}{' permuted' if a.permute else ''}, full multi-phase Louvain",
           "n_gpus": world, "nv": nv, "csr_entries": ne, "levels": len(lq), "level_q": lq, "level_iters": li,
           "level_nv": ln, "seconds_max_rank": float(secs.item()),
           "stage_seconds_rank0": dl.seconds, "warmup": not a.no_warmup,
           "coarsen_stage_seconds_rank0": getattr(dl, "coarsen_seconds", None),
           "communities": int(comm.max().item()) + 1}
    if rank == 0 and not a.no_check:
        one = GpuGraph(stream=side.cuda_stream)
        one.upload_device(g.offsets, g.tails, g.weights)
        if not a.no_warmup:
            one.run()  # warm-up
            one.upload_device(g.offsets, g.tails, g.weights)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        q1, i1, n1, c1 = one.run()
        out["one_gpu_seconds"] = time.perf_counter() - t0
        one.close()
        out["levels_equal_one_gpu"] = bool(lq == q1 and li == i1 and ln == n1)
        out["assignment_equal_one_gpu"] = bool(np.array_equal(comm.cpu().numpy(), c1))
        out["max_abs_dq"] = max((abs(x - y) for x, y in zip(lq, q1)), default=0.0)
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not a.no_check and not (out["levels_equal_one_gpu"] and out["assignment_equal_one_gpu"]):
        sys.exit(1)


if __name__ == "__main__":
    main()
