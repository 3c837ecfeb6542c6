#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "random_distortion_batch" 2>&1 | tail -25
timeout 600 python tools/bench_random_batch.py --batch 128 --steps 3 2>&1 | tail -3
