#!/bin/bash
# launch list of one fog call (1024 x 1024 RGB page) with the field computed on the device
mkdir -p gpurun_out
cat > /tmp/fog_once.py <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from vkit_b200 import element
from vkit_b200.mechanism import distortion
from vkit_b200.mechanism.distortion.photometric import effect
img = element.Image(mat=torch.randint(0, 256, (1024, 1024, 3), dtype=torch.uint8, device='cuda'))
for k in range(2):
    distortion.fog.distort(effect.FogConfig(roughness=0.5), image=img, rng=np.random.default_rng(3 + k))
torch.cuda.synchronize()
PY
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/r02_launches_fog.csv python /tmp/fog_once.py > gpurun_out/r2_ncu_fog.log 2>&1
python tools/ncu_launch_table.py gpurun_out/r02_launches_fog.csv | tail -12
