"""Kernel timeline of bench steps (config 2) from CUPTI through torch.profiler: start, duration and
stream of every kernel / memset of the last step, the idle gaps between them and the step period.

    python tools/timeline_probe.py [pages] [steps]
"""
import json
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

import bench
from vkit_b200.batch import GeometricBatch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
names, configs = bench.sample_page_configs(0, n, 256)
pages = torch.randint(0, 256, (n, 1024, 1024, 3), dtype=torch.uint8, device='cuda')
eng = GeometricBatch(names, configs, (1024, 1024))
for _ in range(3):
    out = eng.run(pages, optimistic=True)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(steps):
        out = eng.run(pages, optimistic=True)
    torch.cuda.synchronize()
path = os.path.join(tempfile.mkdtemp(), 'trace.json')
prof.export_chrome_trace(path)
events = [e for e in json.load(open(path))['traceEvents']
          if e.get('ph') == 'X' and e.get('cat') in ('kernel', 'gpu_memset', 'gpu_memcpy')]
events.sort(key=lambda e: e['ts'])
tiles = [e for e in events if 'grid_remap_tiles' in e['name']]
print(f'{len(events)} device activities, {len(tiles)} steps')
periods = [b['ts'] - a['ts'] for a, b in zip(tiles, tiles[1:])]
print('step period (us, remap start to remap start):', [round(p, 1) for p in periods])
# the last full step: from the end of the previous small-tile remap to the end of the last one
t0 = tiles[-2]['ts'] + tiles[-2]['dur']
t1 = tiles[-1]['ts'] + tiles[-1]['dur']
step = [e for e in events if e['ts'] >= t0 - 200 and e['ts'] + e['dur'] <= t1 + 1]
busy_until = t0
for e in step:
    name = e['name'].split('(')[0].replace('void ', '').replace('vkb::', '')[:44]
    gap = e['ts'] - busy_until
    print(f"{e['ts'] - t0:9.1f} us  +{e['dur']:8.1f}  stream {e['args'].get('stream', '?'):>3}  gap {gap:7.1f}  {name}")
    busy_until = max(busy_until, e['ts'] + e['dur'])
print(f'step span {t1 - t0:.1f} us')
