"""All-gather timing probe (torchrun, one rank per GPU): the label all-gather of a distributed sweep in
isolation — 4 B x nv_local per rank — through torch.distributed's NCCL communicator, timed with CUDA events
over back-to-back calls and, separately, with a host synchronize after every call (the sweep loop's pattern).
Prints one JSON line on rank 0.  NCCL_DEBUG=INFO in the environment shows the algorithm NCCL picked."""
import json
import os
import sys

import torch
import torch.distributed as dist


def main():
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    out = {"world": world, "env": {k: v for k, v in os.environ.items() if k.startswith("NCCL_") and k != "NCCL_DEBUG"}}
    for mb in (1, 4, 8, 32):
        n = mb * (1 << 20) // 4
        src = torch.full((n,), rank, dtype=torch.int32, device="cuda")
        dst = torch.empty(n * world, dtype=torch.int32, device="cuda")
        for _ in range(5):
            dist.all_gather_into_tensor(dst, src)
        torch.cuda.synchronize()
        dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            dist.all_gather_into_tensor(dst, src)
        b.record()
        torch.cuda.synchronize()
        back_to_back = a.elapsed_time(b) / 20
        tot = 0.0
        for _ in range(20):
            a.record()
            dist.all_gather_into_tensor(dst, src)
            b.record()
            torch.cuda.synchronize()
            tot += a.elapsed_time(b)
        t = torch.tensor([back_to_back, tot / 20], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out[f"{mb}MiB_per_rank"] = {"ms_back_to_back": round(t[0].item(), 4), "ms_with_sync": round(t[1].item(), 4),
                                    "busbw_GBs": round(mb * (1 << 20) * (world - 1) / (t[0].item() * 1e-3) / 1e9, 1)}
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
