#!/bin/bash
# Round-2 evidence run on ONE B200 (under gpurun): final bench line, reference arm, ncu launch list with
# DRAM bytes (roofline.traffic), ncu --set full of the three dominant kernels, C3 golden summary.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu.log
python bench.py --steps 12 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_ref.err
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,l1tex__throughput.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct \
    --clock-control none --cache-control none -k regex:"k_sweep|k_apply|k_pack|k_lm|k_big" -c 200 --csv \
    --log-file gpurun_out/r02_launches_scale20.csv python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-scale24 > /dev/null 2>&1
ncu --set full --import-source on --clock-control none --kernel-name-base demangled \
    -k regex:"k_sweep_u32_hash2<.int.256|k_big_stage|k_big_pass|k_lm_first_group<.int.32|k_lm_group4<.int.4, .int.8" -c 8 -f \
    -o gpurun_out/r02_top_kernels python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-scale24 > gpurun_out/r02_ncu_full.log 2>&1
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,l1tex__throughput.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct \
    --clock-control none --cache-control none -k regex:"k_sweep|k_apply|k_pack|k_lm|k_big" -c 120 --csv \
    --log-file gpurun_out/r02_launches_scale24.csv python bench.py --scale 24 --steps 4 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python tools/run_configs.py c3 > gpurun_out/r02_c3.log 2>&1
cp gpurun_out/configs.json gpurun_out/r02_configs_c3.json
echo done
