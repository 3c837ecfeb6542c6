"""BASELINE config 4's distortion stage at batch scale: RandomDistortionBatch over pages of
1024 x 1024 RGB + mask + 256 points + 64 quadrilaterals per page, the pipeline's default policy
config (page_distortion.py:53-64: defocus_blur / zoom_in_blur disabled, post rotate forced).
Prints one JSON line (pages/s incl. the host-side chain sampling and the coordinate traffic).

    python tools/bench_random_batch.py [--batch 128] [--steps 3]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vkit_b200.mechanism.distortion_policy import random_distortion_factory  # noqa: E402
from vkit_b200.mechanism.distortion_policy.random_distortion_batch import RandomDistortionBatch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=128)
    ap.add_argument('--steps', type=int, default=3)
    args = ap.parse_args()
    n, shape = args.batch, (1024, 1024)
    rd = random_distortion_factory.create({'disabled_policy_names': ['defocus_blur', 'zoom_in_blur'],
                                           'force_post_rotate': True})
    batch = RandomDistortionBatch(rd)
    images = torch.randint(0, 256, (n,) + shape + (3,), dtype=torch.uint8, device='cuda')
    masks = (torch.rand((n,) + shape, device='cuda') > 0.5).to(torch.uint8)
    gen = np.random.default_rng(7)
    points = [gen.uniform(0, 1023, (256, 2)) for _ in range(n)]
    polygons = []
    for _ in range(n):
        polys = []
        for _ in range(64):
            x0, y0 = gen.uniform(0, 900, 2)
            w, h = gen.uniform(10, 120, 2)
            polys.append(np.asarray([(x0, y0), (x0 + w, y0), (x0 + w, y0 + h), (x0, y0 + h)]))
        polygons.append(polys)
    seqs = np.random.SeedSequence(133700).spawn(n * (args.steps + 1))

    def step(k):
        rngs = [np.random.default_rng(s) for s in seqs[k * n:(k + 1) * n]]
        return batch.distort(rngs, images, masks, points, polygons)

    step(0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(args.steps):
        out = step(k + 1)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / args.steps
    names = {}
    for r in out:
        for name in r.chain.names:
            names[name] = names.get(name, 0) + 1
    print(json.dumps({'workload': f'config 4 distortion stage, RandomDistortionBatch, batch {n}, '
                                  '1024x1024 RGB + mask + 256 points + 64 polygons per page',
                      'pages_per_s': n / dt, 'ms_per_page': dt / n * 1e3, 'ops_last_step': names}))


if __name__ == '__main__':
    main()
