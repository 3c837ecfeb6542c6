#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/r2_sweep2.log
for v in base share; do
  echo "== $v" >> gpurun_out/r2_sweep2.log
  for b in 256; do
  VKB_LIB=$PWD/variants/libvkit_$v.so timeout 120 python bench.py --steps 20 --warmup 3 --kernel-only --batch $b 2>&1 | python -c "
import sys,json
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('batch $b value %.0f pages/s  step %.3f ms  remap %.3f ms' % (d['value'], d['ms_per_step'], d['roofline']['launch_ms']))
    else: print(line.rstrip())
" >> gpurun_out/r2_sweep2.log 2>&1
  done
done
cat gpurun_out/r2_sweep2.log
