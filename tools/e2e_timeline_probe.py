"""Copy / kernel timeline of one end-to-end step (distort_pages_host, 256 pages) from CUPTI.

    python tools/e2e_timeline_probe.py
"""
import json
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

import bench

w = bench.Workload2(256, 0, 256, np.arange(256) + bench.BASE_SEED)
w.step(False)
torch.cuda.synchronize()
for _ in range(3):
    w.e2e_step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    w.e2e_step()
    torch.cuda.synchronize()
path = os.path.join(tempfile.mkdtemp(), 'trace.json')
prof.export_chrome_trace(path)
events = [e for e in json.load(open(path))['traceEvents']
          if e.get('ph') == 'X' and e.get('cat') in ('kernel', 'gpu_memset', 'gpu_memcpy')]
events.sort(key=lambda e: e['ts'])
t0 = events[0]['ts']
print(f'{len(events)} device activities, span {(events[-1]["ts"] + events[-1]["dur"] - t0) / 1e3:.2f} ms')
copies = [e for e in events if e['cat'] == 'gpu_memcpy' and e['dur'] > 50]
for e in copies:
    name = e['name']
    nbytes = e['args'].get('bytes', 0)
    print(f"{(e['ts'] - t0) / 1e3:8.3f} ms  +{e['dur'] / 1e3:7.3f} ms  {nbytes / 1e6:8.1f} MB  "
          f"{nbytes / max(e['dur'], 1) / 1e3:6.1f} GB/s  stream {e['args'].get('stream')}  {name[:40]}")
remaps = [e for e in events if 'grid_remap_tiles' in e['name']]
print('remap launches at (ms):', [round((e['ts'] - t0) / 1e3, 2) for e in remaps])
