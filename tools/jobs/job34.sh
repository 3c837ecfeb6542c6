#!/bin/bash
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -x -q -m gpu -k "not full_size and not 1024" > gpurun_out/r2_memcheck_final.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/r2_memcheck_final.log
tail -5 gpurun_out/r2_memcheck_final.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests -x -q -m gpu -k "(geometric_vs_golden and c02) or cb00 or tl00 or mls_projection or variants_vs_oracle or batched_text" > gpurun_out/r2_racecheck_final.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/r2_racecheck_final.log
tail -4 gpurun_out/r2_racecheck_final.log
