#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_c2.log 2> gpurun_out/r2_bench_c2.err; echo "rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_c2.log').read().strip().split('\n')[-1])
print('c2 value', d['value'], 'e2e', d['e2e']['value'], 'parity ok', d.get('parity',{}).get('ok'), 'frac', d['roofline']['frac'], 'host', d['host_ms_per_step'])
PY
tail -5 gpurun_out/r2_bench_c2.err
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
