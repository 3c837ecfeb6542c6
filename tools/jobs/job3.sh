#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests.log 2>&1; tail -3 gpurun_out/r2_tests.log
timeout 300 python tools/step_probe.py 256 > gpurun_out/r2_step_probe.log 2>&1; cat gpurun_out/r2_step_probe.log
timeout 300 python bench.py --steps 10 --warmup 3 --kernel-only > gpurun_out/r2_bench_k.log 2>&1; python - <<'PY'
import json
for line in open('gpurun_out/r2_bench_k.log'):
    if line.startswith('{'):
        d=json.loads(line); print('value %.0f pages/s  step %.3f ms  remap %.3f ms' % (d['value'], d['ms_per_step'], d['roofline']['launch_ms']))
PY
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 40 --csv \
    --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --batch 64 --kernel-only > gpurun_out/r2_ncu_list.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r2_launches.csv')))
hdr=None; data={}
for r in rows:
    if r and r[0]=='ID': hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r)); data.setdefault(d['ID'],{'k':d['Kernel Name'][:60]})[d['Metric Name']]=d['Metric Value']
for i,v in list(data.items())[-13:]: print(i, v)
PY
