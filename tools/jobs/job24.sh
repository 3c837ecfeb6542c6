#!/bin/bash
mkdir -p gpurun_out
python tools/bench_random_batch.py --batch 128 --steps 3 2>&1 | tail -1
python -c "
import cProfile, pstats, sys, runpy
sys.argv=['tools/bench_random_batch.py','--batch','128','--steps','3']
cProfile.run(\"runpy.run_path('tools/bench_random_batch.py', run_name='__main__')\", '/tmp/prof.out')
p=pstats.Stats('/tmp/prof.out'); p.sort_stats('cumulative').print_stats(70)
" > gpurun_out/r2_rdb_profile.log 2>&1
grep -n 'function calls' gpurun_out/r2_rdb_profile.log | head -2
