"""torchrun worker: N-GPU phase 0 through cuvite_b200.dist must equal the 1-rank
reference (golden fixture) bit for bit; the multi-level run (phases + distributed coarsening)
must reproduce the reference's levels and final assignment."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from cuvite_b200.dist import DistLouvain, init_process_group
    from cuvite_b200.engine import GpuGraph, set_gpuDevices
    from cuvite_b200.vite_io import partition_ranges
    rank, local, world = init_process_group("nccl")
    set_gpuDevices(local)
    torch.cuda.set_device(local)
    stream = torch.cuda.Stream()
    ok = True
    with torch.cuda.stream(stream):
        for name, exchange in (("rmat10_intw", "labels"), ("hub30k", "labels"), ("er1k", "labels"), ("hub30k", "deltas")):
            g = np.load(os.path.join(ROOT, "tests", "golden", f"{name}.npz"))
            o, t, w = g["offsets"], g["tails"].astype(np.int64), g["weights"]
            nv = o.size - 1
            parts = partition_ranges(nv, world)
            b0, b1 = parts[rank], parts[rank + 1]
            gpu = GpuGraph(stream=stream.cuda_stream)
            gpu.upload(o[b0:b1 + 1] - o[b0], t[o[b0]:o[b1]], w[o[b0]:o[b1]], base=b0, nv_global=nv)
            q, iters = DistLouvain(gpu, rank, world, nv, exchange=exchange).run_phase()
            good = iters == int(g["iters"]) and q == float(g["q"]) and np.array_equal(gpu.labels(0), g["cvect"][b0:b1])
            print(f"rank {rank} {name} [{exchange}]: q={q!r} iters={iters} ok={good}", flush=True)
            ok &= bool(good)
            gpu.close()
        # phases + distributed coarsening: levels and final assignment of the 1-rank reference
        for name in ("rmat10_intw", "hub30k", "er1k", "ring6x5"):
            g = np.load(os.path.join(ROOT, "tests", "golden", f"{name}.npz"))
            o, t, w = g["offsets"], g["tails"].astype(np.int64), g["weights"]
            nv = o.size - 1
            parts = partition_ranges(nv, world)
            b0, b1 = parts[rank], parts[rank + 1]
            gpu = GpuGraph(stream=stream.cuda_stream)
            gpu.upload(o[b0:b1 + 1] - o[b0], t[o[b0]:o[b1]], w[o[b0]:o[b1]], base=b0, nv_global=nv)
            lq, li, ln, comm = DistLouvain(gpu, rank, world, nv).run()
            good = (lq == list(g["level_q"]) and li == list(g["level_iters"]) and ln == list(g["level_nv"])
                    and np.array_equal(comm.cpu().numpy(), g["comm"]))
            print(f"rank {rank} {name} [multi-level run]: q={lq} iters={li} nv={ln} ok={good}", flush=True)
            ok &= bool(good)
            gpu.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
