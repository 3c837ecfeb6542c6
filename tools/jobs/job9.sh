#!/bin/bash
# pure concurrent copies at 4 and 8 ranks vs the bench's e2e at 8 ranks
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo8.log 2>&1
for n in 4 8; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n tools/pcie_probe_multi.py > gpurun_out/r2_pcie_n$n.log 2>&1
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --steps 10 --warmup 3 --skip-cpu-baseline > gpurun_out/r2_bench_n8.log 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 10 --warmup 3 --skip-cpu-baseline > gpurun_out/r2_bench_n4.log 2>&1
cat gpurun_out/r2_pcie_n4.log gpurun_out/r2_pcie_n8.log | grep ranks
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench_n[48].log')):
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line); print(f, 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'host', d['host_ms_per_step'])
PY
head -12 gpurun_out/r2_topo8.log; nproc
