#!/bin/bash
timeout 600 python - <<'PY' 2>&1 | tail -45
import cProfile, pstats, sys, os
sys.path.insert(0, os.getcwd()); sys.argv=['x','--batch','128','--steps','2']
sys.path.insert(0, 'tools')
import bench_random_batch as b
cProfile.run('b.main()', '/tmp/prof.out')
p = pstats.Stats('/tmp/prof.out'); p.sort_stats('cumulative').print_stats(38)
PY
