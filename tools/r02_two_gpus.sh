#!/bin/bash
# N=2 check of the N-rank end-to-end leg and of the in-code NVLS default (NCCL init log)
set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
export CVB_DIST_TIMING=1
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,ENV $TR --nproc-per-node 2 --master-port 29591 bench.py --gpus 2 --steps 12 --warmup 3 > gpurun_out/r02_bench_n2_v2.log 2> gpurun_out/r02_bench_n2_v2.err
grep '^{' gpurun_out/r02_bench_n2_v2.log > gpurun_out/r02_bench_n2_v2.json
grep -E "NVLS|NCCL_NVLS" gpurun_out/r02_bench_n2_v2.log | sort | uniq -c | head -10 > gpurun_out/r02_bench_n2_nvls_lines.txt
tail -3 gpurun_out/r02_bench_n2_v2.err
timeout 600 python -m pytest tests/test_gpu_dist.py -q -x 2>&1 | tail -3
echo done
