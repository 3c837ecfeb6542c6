"""Does the prep of batch i+1 overlap with the remap of batch i when the persistent remap leaves a
block slot per SM free?  Two engines (own workspaces) on two streams, steps issued alternately;
VKB_TILES_GRID_BLOCKS=3 limits the remap to 3 of its 4 resident blocks per SM.

    VKB_TILES_GRID_BLOCKS=3 python tools/overlap_probe.py [pages] [steps]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from vkit_b200.batch import GeometricBatch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
names, configs = bench.sample_page_configs(0, n, 256)
pages = torch.randint(0, 256, (n, 1024, 1024, 3), dtype=torch.uint8, device='cuda')
engines = [GeometricBatch(names, configs, (1024, 1024)) for _ in range(2)]
low, high = torch.cuda.Stream.priority_range() if hasattr(torch.cuda.Stream, 'priority_range') else (0, -1)
streams = [torch.cuda.Stream(), torch.cuda.Stream()]


def run(mode):
    outs = []
    for k in range(steps):
        if mode == 'one':
            outs.append(engines[0].run(pages, optimistic=True))
        else:
            with torch.cuda.stream(streams[k & 1]):
                outs.append(engines[k & 1].run(pages, optimistic=True))
        if len(outs) > 4:
            outs.pop(0)
    return outs


for mode in ('one', 'two', 'one', 'two'):
    run(mode)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run(mode)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    print(f'{mode}: {dt * 1e3:.3f} ms per step, {n / dt:.0f} pages/s  (VKB_TILES_GRID_BLOCKS={os.environ.get("VKB_TILES_GRID_BLOCKS")})')
