#!/bin/bash
# e2e scaling probe on 2 GPUs: pure concurrent copies vs the bench's e2e
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo.log 2>&1
lscpu | head -30 >> gpurun_out/r2_topo.log 2>&1
python tools/pcie_probe_multi.py > gpurun_out/r2_pcie_n1.log 2>&1
for n in 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 tools/pcie_probe_multi.py > gpurun_out/r2_pcie_n$n.log 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --steps 10 --warmup 3 --skip-cpu-baseline > gpurun_out/r2_bench_n$n.log 2>&1
done
cat gpurun_out/r2_pcie_n*.log | grep ranks
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench_n*.log')):
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line); print(f, 'value', round(d['value']), 'e2e', round(d['e2e']['value']))
PY
head -20 gpurun_out/r2_topo.log
