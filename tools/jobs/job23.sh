#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -3
python bench.py > gpurun_out/r2_final_c2.log 2> gpurun_out/r2_final_c2.err; tail -c 600 gpurun_out/r2_final_c2.err; python - <<'PY'
import json
for f in ['gpurun_out/r2_final_c2.log']:
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line); print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches','host_ms_per_step','parity_gate','clocks')}); print(d['roofline']); print(d['e2e']); print(d['cpu_baseline'])
PY
python bench.py --impl reference > gpurun_out/r2_final_ref.log 2>&1; tail -2 gpurun_out/r2_final_ref.log | cut -c1-400
python bench.py --config 3 --skip-cpu-baseline > gpurun_out/r2_final_c3.log 2>&1; tail -1 gpurun_out/r2_final_c3.log | cut -c1-600
