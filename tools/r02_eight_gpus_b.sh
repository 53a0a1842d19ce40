#!/bin/bash
# Round-2 multi-GPU follow-up (under gpurun --gpus 8): the label all-gather in isolation (NCCL algorithm
# variants), then C4 / C5 full multi-level runs WITH the untimed warm-up run (NCCL channels, allocator).
set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/r02_topo.txt 2>&1
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,TUNING $TR --nproc-per-node 8 --master-port 29581 tools/nccl_allgather_probe.py > gpurun_out/r02_allgather_probe_default.json 2> gpurun_out/r02_allgather_probe_default.err
grep -E "NVLS|Ring|Tree|channels|AllGather" gpurun_out/r02_allgather_probe_default.json | head -40 > gpurun_out/r02_nccl_info.txt
NCCL_NVLS_ENABLE=0 $TR --nproc-per-node 8 --master-port 29582 tools/nccl_allgather_probe.py > gpurun_out/r02_allgather_probe_nonvls.json 2>/dev/null
NCCL_ALGO=Ring $TR --nproc-per-node 8 --master-port 29583 tools/nccl_allgather_probe.py > gpurun_out/r02_allgather_probe_ring.json 2>/dev/null
$TR --nproc-per-node 4 --master-port 29584 tools/nccl_allgather_probe.py > gpurun_out/r02_allgather_probe_n4.json 2>/dev/null
export CVB_DIST_TIMING=1
NCCL_NVLS_ENABLE=0 $TR --nproc-per-node 8 --master-port 29585 bench.py --gpus 8 --steps 12 --warmup 3 > gpurun_out/r02_bench_n8_nonvls.json 2> /dev/null
$TR --nproc-per-node 8 --master-port 29586 tools/run_dist_full.py --scale 24 --permute > gpurun_out/r02_c4_full_run_warm.json 2> gpurun_out/r02_c4_full_warm.err
$TR --nproc-per-node 8 --master-port 29587 tools/run_c5.py --warmup > gpurun_out/r02_c5_eight_gpus_warm.json 2> gpurun_out/r02_c5_warm.err
echo done
