#!/bin/bash
# Round-2 final multi-GPU record (under gpurun --gpus 8): bench at N=8 and N=4 with the library's NCCL defaults
# (NVLS off) and the N-rank end-to-end leg; C4 full multi-level run, warm, with coarsening stage times.
set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
export CVB_DIST_TIMING=1
$TR --nproc-per-node 8 --master-port 29571 bench.py --gpus 8 --steps 12 --warmup 3 > gpurun_out/r02_bench_n8_final.json 2> gpurun_out/r02_bench_n8_final.err
$TR --nproc-per-node 4 --master-port 29572 bench.py --gpus 4 --steps 12 --warmup 3 > gpurun_out/r02_bench_n4_final.json 2> gpurun_out/r02_bench_n4_final.err
$TR --nproc-per-node 8 --master-port 29574 tools/run_dist_full.py --scale 24 --permute > gpurun_out/r02_c4_full_run_final.json 2> gpurun_out/r02_c4_full_final.err
tail -2 gpurun_out/r02_bench_n8_final.err gpurun_out/r02_c4_full_final.err
echo done
