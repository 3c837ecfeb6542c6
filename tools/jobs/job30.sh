#!/bin/bash
run() { python bench.py --skip-cpu-baseline --steps 10 2>/dev/null | python -c "
import sys,json
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('$1', round(d['value']), 'e2e', round(d['e2e']['value']))
"; }
cp vkit_b200/batch.py /tmp/batch_new.py
for rep in 1 2 3; do
  cp /tmp/batch_new.py vkit_b200/batch.py; run new
  cp tools/batch_old.py.txt vkit_b200/batch.py; run old
done
cp /tmp/batch_new.py vkit_b200/batch.py
