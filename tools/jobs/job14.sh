#!/bin/bash
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 120 python bench.py --steps 20 --warmup 3 --kernel-only --batch 32 2>&1 | tail -c 400
echo
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python bench.py --steps 3 --warmup 1 --kernel-only --batch 32 2>&1 | grep -v "^$" | grep -v '^{"metric' | head -60
