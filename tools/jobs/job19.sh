#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/r2_tests2.log 2>&1
tail -5 gpurun_out/r2_tests2.log
python - > gpurun_out/r2_compose_probe.log 2>&1 <<'PY'
import sys, time
sys.path.insert(0, 'tests')
import numpy as np, torch
from common import f4_textures
from vkit_b200.background import ImageCombiner, ImageCombinerConfig, Texture
from vkit_b200.element import Image
tex = f4_textures(7005, 12, 100, 300)
comb = ImageCombiner([Texture(n, Image(mat=m), a, s) for n, m, a, s in tex], ImageCombinerConfig(prob_use_only_the_anchor_image=0.3))
for shape in [(1024, 1024), (2522, 2522)]:
    rng = np.random.default_rng(3)
    plans = []
    t0 = time.perf_counter()
    for _ in range(20):
        cands = comb.sample_candidates(rng)
        plans.append(comb.plan(shape[0], shape[1], cands, rng))
    t_plan = (time.perf_counter() - t0) / 20
    for p in plans[:3]:
        comb.compose(shape[0], shape[1], p)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record()
    for p in plans:
        comb.compose(shape[0], shape[1], p)
    ev1.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / 20
    print(f'{shape}: plan {t_plan*1e3:.3f} ms host, compose {ev0.elapsed_time(ev1)/20:.3f} ms device / {wall*1e3:.3f} ms wall per page, segments {np.mean([len(p) for p in plans]):.0f}')
PY
cat gpurun_out/r2_compose_probe.log
