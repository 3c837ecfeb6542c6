"""Where one bench step (256 pages) spends its time: CUDA events between the C-ABI calls of
GeometricBatch.plan_batch / run, host wall clock beside them, and a per-kernel list from
back-to-back launches of the build kernels alone.

    python tools/step_probe.py [pages]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from vkit_b200.batch import GeometricBatch
from vkit_b200.mechanism.distortion.geometric import _gridcore

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
names, configs = bench.sample_page_configs(0, n, 256)
pages = torch.randint(0, 256, (n, 1024, 1024, 3), dtype=torch.uint8, device='cuda')
eng = GeometricBatch(names, configs, (1024, 1024))
OPT = os.environ.get('PROBE_EXACT') is None
for _ in range(3):
    out = eng.run(pages, optimistic=OPT)
torch.cuda.synchronize()

marks = []


def mark(label):
    ev = torch.cuda.Event(enable_timing=True)
    ev.record()
    marks.append((label, ev, time.perf_counter()))


# instrument the phases by wrapping GridBatch methods
orig_init, orig_build, orig_remap = _gridcore.GridBatch.__init__, _gridcore.GridBatch.build, _gridcore.GridBatch.remap


def init(self, *a, **k):
    mark('start')
    orig_init(self, *a, **k)
    mark('project+finalize+sync (GridBatch.__init__)')


def build(self, *a, **k):
    orig_build(self, *a, **k)
    mark('allocs + cells/masks/tiles launches (build)')


def remap(self, *a, **k):
    mark('planes arrays (host)')
    r = orig_remap(self, *a, **k)
    mark('remap launch')
    return r


_gridcore.GridBatch.__init__, _gridcore.GridBatch.build, _gridcore.GridBatch.remap = init, build, remap
reps = 5
rows = {}
for rep in range(reps):
    marks.clear()
    torch.cuda.synchronize()
    out = eng.run(pages, optimistic=OPT)
    mark('end')
    torch.cuda.synchronize()
    for (l0, e0, t0), (l1, e1, t1) in zip(marks, marks[1:]):
        rows.setdefault(l1, []).append((e0.elapsed_time(e1), (t1 - t0) * 1e3))
print(f'{n} pages per step; per segment: GPU ms between events (median) | host ms')
tot_g = tot_h = 0.0
for label, vals in rows.items():
    g = float(np.median([v[0] for v in vals]))
    h = float(np.median([v[1] for v in vals]))
    tot_g += g
    tot_h += h
    print(f'  {label:50s} {g:7.3f} | {h:7.3f}')
print(f'  {"total":50s} {tot_g:7.3f} | {tot_h:7.3f}')
