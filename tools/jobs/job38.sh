#!/bin/bash
# bench lines of the final build at N ranks: bash tools/jobs/job38.sh N
N=$1
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2_final_n$N.log 2>&1
echo rc=$? >> gpurun_out/r2_final_n$N.log
python - <<PY
import json
for l in open("gpurun_out/r2_final_n$N.log"):
    if l.startswith("{"):
        d=json.loads(l); print("N=$N value %.0f step %.4f e2e %.0f remap %.4f clocks %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["launch_ms"], d["clocks"]))
    elif "rc=" in l: print(l.strip())
PY
