# usage: bash tools/gpu_ab_lib.sh <tag> <lib or -> ...   -- kernel-only bench line per library ("-" = in-tree)
tag=$1; shift
for rep in 1 2; do
for v in "$@"; do
  echo "== $v" >> gpurun_out/${tag}_ab.log
  if [ "$v" = "-" ]; then unset VKB_LIB; else export VKB_LIB=$PWD/$v; fi
  python bench.py --kernel-only --steps 30 --warmup 5 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('value %.0f step %.4f remap %.4f' % (d['value'], d['ms_per_step'], d['roofline']['launch_ms']))
" >> gpurun_out/${tag}_ab.log
done
done
cat gpurun_out/${tag}_ab.log
