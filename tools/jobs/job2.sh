#!/bin/bash
# variant sweep of the small-tile remap kernel (kernel-only timing), then parity of the default build
mkdir -p gpurun_out
for v in base nopipe pipe b4 pipe_b4 nopipe_b4 pipe_w7 nopipe_w7 pipe_w7r96 pipe_w10r96 pipe_w10b2; do
  echo "== $v" >> gpurun_out/r2_sweep.log
  VKB_LIB=$PWD/variants/libvkit_$v.so timeout 120 python bench.py --steps 10 --warmup 3 --kernel-only 2>&1 | python -c "
import sys,json
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('value %.0f pages/s  step %.3f ms  remap %.3f ms' % (d['value'], d['ms_per_step'], d['roofline']['launch_ms']))
    else: print(line.rstrip())
" >> gpurun_out/r2_sweep.log 2>&1
done
cat gpurun_out/r2_sweep.log
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests.log 2>&1; tail -3 gpurun_out/r2_tests.log
