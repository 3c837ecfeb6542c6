#!/bin/bash
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_pre_compositing.py tests/test_gpu_parity.py -x -q -m gpu -k "compose or atlas or lcd or ellipse or jpeg or tma or optimistic or fog_gray or (geometric_vs_golden and 136)" > gpurun_out/r2_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/r2_memcheck.log
tail -6 gpurun_out/r2_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_pre_compositing.py -x -q -m gpu -k "(geometric_vs_golden and c02) or cb00 or tl00" > gpurun_out/r2_racecheck.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/r2_racecheck.log
tail -6 gpurun_out/r2_racecheck.log
