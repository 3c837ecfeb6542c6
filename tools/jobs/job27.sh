#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python bench.py --config 3 --skip-cpu-baseline 2>/dev/null | python -c "
import sys,json
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('c3', d['value'], d['ms_per_step'], d['host_ms_per_step'], d['roofline']['launch_ms'], d['e2e']['value'])
"
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:"grid_project_mls" -c 4 --csv --log-file gpurun_out/r2_mls.csv python bench.py --config 3 --steps 1 --warmup 1 --batch 256 --kernel-only --skip-cpu-baseline > /dev/null 2>&1
python tools/ncu_launch_table.py gpurun_out/r2_mls.csv
