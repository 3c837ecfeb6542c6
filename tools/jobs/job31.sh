#!/bin/bash
cp vkit_b200/batch.py /tmp/batch_new.py
echo "=== new"; python tools/e2e_timeline_probe.py 2>&1 | grep -v Warning | tail -30
cp tools/batch_old.py.txt vkit_b200/batch.py
echo "=== old"; python tools/e2e_timeline_probe.py 2>&1 | grep -v Warning | tail -30
cp /tmp/batch_new.py vkit_b200/batch.py
