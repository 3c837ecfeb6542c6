#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py --skip-cpu-baseline --kernel-only 2>/dev/null | python -c "
import sys,json
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('value %.0f step %.3f remap %.3f' % (d['value'], d['ms_per_step'], d['roofline']['launch_ms']))
"
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:"grid_masks|grid_cells|grid_tile_records" -c 9 --csv --log-file gpurun_out/r2_masks.csv python bench.py --steps 2 --warmup 1 --batch 256 --kernel-only --skip-cpu-baseline > /dev/null 2>&1
python tools/ncu_launch_table.py gpurun_out/r2_masks.csv
