#!/bin/bash
# Round-2 multi-GPU evidence (under gpurun --gpus 8): weak-scaling bench at N=8 and N=4 with stage timing,
# C4 (scale 24 on 8 GPUs: bench step + full multi-level run vs one GPU), C5 (power law, 512 M edges, -b ranges).
set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
export CVB_DIST_TIMING=1
$TR --nproc-per-node 8 --master-port 29571 bench.py --gpus 8 --steps 12 --warmup 3 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err
$TR --nproc-per-node 4 --master-port 29572 bench.py --gpus 4 --steps 12 --warmup 3 > gpurun_out/r02_bench_n4.json 2> gpurun_out/r02_bench_n4.err
$TR --nproc-per-node 8 --master-port 29573 bench.py --gpus 8 --scale 24 --steps 8 --warmup 3 > gpurun_out/r02_c4_eight_gpus_scale24.json 2> gpurun_out/r02_c4.err
$TR --nproc-per-node 8 --master-port 29574 tools/run_dist_full.py --scale 24 --permute --no-warmup > gpurun_out/r02_c4_full_run.json 2> gpurun_out/r02_c4_full.err
$TR --nproc-per-node 8 --master-port 29575 tools/run_c5.py > gpurun_out/r02_c5_eight_gpus.json 2> gpurun_out/r02_c5.err
tail -c 400 gpurun_out/r02_c5.err
echo done
