#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu -k "blur or chain or golden or photometric" > gpurun_out/r2_tests3.log 2>&1; tail -4 gpurun_out/r2_tests3.log
python tools/blur_tma_probe.py > gpurun_out/r2_blur_tma.log 2>&1; cat gpurun_out/r2_blur_tma.log
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:gaussian_blur --csv \
    --log-file gpurun_out/r02_blur_tma_launches.csv python tools/blur_tma_probe.py --once > gpurun_out/r2_blur_tma_ncu.log 2>&1
tail -3 gpurun_out/r2_blur_tma_ncu.log
