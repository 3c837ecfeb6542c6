"""Host-side time per phase of one 32-page chunk of the e2e pipeline (perf_counter, no extra
syncs): where the Python / driver time of distort_pages_host goes."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, '/root/repo')
import bench  # noqa: E402
from vkit_b200.batch import GeometricBatch  # noqa: E402

n = 32
names, configs = bench.sample_page_configs(0, n, n)
dev = torch.randint(0, 256, (n, 1024, 1024, 3), dtype=torch.uint8, device='cuda')
acc = {}


def tick(key, t0):
    t1 = time.perf_counter()
    acc[key] = acc.get(key, 0.0) + (t1 - t0) * 1e3
    return t1


reps = 20
for it in range(reps + 3):
    if it == 3:
        acc.clear()
    t = time.perf_counter()
    eng = GeometricBatch(names, configs, (1024, 1024))
    t = tick('records', t)
    eng.plan_batch()
    t = tick('plan (project+finalize+sync)', t)
    eng.plan.build()
    t = tick('build (allocs + 6 launches)', t)
    out = eng.run(dev, replan=False)
    t = tick('run (arena, planes, remap launch)', t)
    torch.cuda.synchronize()
    t = tick('drain', t)
for k, v in acc.items():
    print(f'{k:40s} {v / reps:8.3f} ms per 32-page chunk')
