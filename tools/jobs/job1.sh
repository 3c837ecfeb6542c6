#!/bin/bash
# first GPU pass of round 2: parity of the new small-tile remap kernel, A/B against v1, ncu
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_smi.log 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2_tests.log
tail -5 gpurun_out/r2_tests.log
VKB_REMAP_V1=1 timeout 300 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline > gpurun_out/r2_bench_v1.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline > gpurun_out/r2_bench_v2.log 2>&1
tail -c 1500 gpurun_out/r2_bench_v1.log; echo; tail -c 1500 gpurun_out/r2_bench_v2.log; echo
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 60 --csv \
    --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --batch 64 --skip-cpu-baseline > gpurun_out/r2_ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"grid_remap_tiles" -s 2 -c 1 \
    -o gpurun_out/r2_tiles python bench.py --steps 1 --warmup 1 --batch 32 --skip-cpu-baseline > gpurun_out/r2_ncu_full.log 2>&1
ls -la gpurun_out | tail -8
