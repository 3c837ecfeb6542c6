# usage: bash tools/gpu_ab.sh <tag> "<ENV=.. ENV=..>" ...   -- one kernel-only bench line per variant
tag=$1; shift
for v in "$@"; do
  echo "== $v" >> gpurun_out/${tag}_ab.log
  env $v python bench.py --kernel-only --steps 30 --warmup 5 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('value %.0f step %.4f remap %.4f' % (d['value'], d['ms_per_step'], d['roofline']['launch_ms']))
" >> gpurun_out/${tag}_ab.log
done
cat gpurun_out/${tag}_ab.log
