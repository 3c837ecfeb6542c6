#!/bin/bash
# final build of round 2: tests, smoke, bench lines (config 2 default, config 3), launch lists, timeline
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r2z_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1
python bench.py > gpurun_out/r2z_bench_c2.log 2>&1; echo rc=$? >> gpurun_out/r2z_bench_c2.log
python bench.py --config 3 > gpurun_out/r2z_bench_c3.log 2>&1; echo rc=$? >> gpurun_out/r2z_bench_c3.log
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 20 -c 20 --csv \
    --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 1 --batch 256 --kernel-only > gpurun_out/r2_ncu_list.log 2>&1
python tools/timeline_probe.py > gpurun_out/r2_timeline.log 2>&1
cat gpurun_out/r2z_tests.log gpurun_out/r2z_smoke.log; tail -c 600 gpurun_out/r2z_bench_c2.log
