#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 3 --kernel-only > gpurun_out/r2_bench_k.log 2>&1; python - <<'PY'
import json
for line in open('gpurun_out/r2_bench_k.log'):
    if line.startswith('{'):
        d=json.loads(line); print('value %.0f pages/s  step %.3f ms  remap %.3f ms host %.3f ms' % (d['value'], d['ms_per_step'], d['roofline']['launch_ms'], d.get('host_ms_per_step', 0)))
    else: print(line.rstrip())
PY
