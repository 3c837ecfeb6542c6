#!/bin/bash
run() { python bench.py --skip-cpu-baseline --steps 10 2>/dev/null | python -c "
import sys,json
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('$1', round(d['value']), 'e2e', round(d['e2e']['value']))
"; }
for rep in 1 2; do
for d in 4 8 6; do VKB_E2E_FIRST_DIV=$d run div$d; done
done
