python - <<'PY' 2>&1 | cut -c1-170 | head -60
import cProfile, pstats, sys, io
sys.argv=['bench.py','--config','4','--steps','2','--warmup','1']
sys.path.insert(0,'/root/repo')
import runpy
pr=cProfile.Profile(); pr.enable()
try:
    runpy.run_path('/root/repo/bench.py', run_name='__main__')
except SystemExit: pass
pr.disable()
st=pstats.Stats(pr); st.sort_stats('cumulative').print_stats(45)
PY
