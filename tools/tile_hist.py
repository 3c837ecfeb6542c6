"""Histogram of candidate cells per 32x32 dst tile for the bench workload."""
import sys
import numpy as np
import torch
sys.path.insert(0, '/root/repo')
import bench
from vkit_b200.batch import GeometricBatch

n = 256
names, configs = bench.sample_page_configs(0, n, n)
eng = GeometricBatch(names, configs, (1024, 1024))
plan = eng.plan_batch()
tc = plan.tile_count.cpu().numpy()
meta = plan.meta
counts = []
for i in range(n):
    tiles = ((int(meta['dst_w'][i]) + 31) // 32) * ((int(meta['dst_h'][i]) + 31) // 32)
    counts.append(tc[i, :tiles])
c = np.concatenate(counts)
print('tiles', c.size, 'mean', c.mean(), 'p50', np.percentile(c, 50), 'p99', np.percentile(c, 99), 'max', c.max())
for thr in (8, 12, 16, 20, 24, 32, 48, 64):
    print(f'count > {thr}: {(c > thr).mean():.5%}')
print('empty tiles', (c == 0).mean())
