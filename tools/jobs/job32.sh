#!/bin/bash
run() { python bench.py --skip-cpu-baseline --steps 10 2>/dev/null | python -c "
import sys,json
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('$1', round(d['value']), 'e2e', round(d['e2e']['value']))
"; }
python -m pytest tests -x -q -m gpu -k "pages_host or host" 2>&1 | tail -2
for rep in 1 2 3; do run graded; done
python tools/e2e_timeline_probe.py 2>&1 | grep -v Warn | tail -28
