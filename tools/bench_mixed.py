"""BASELINE config 5: mixed-resolution sweep (256 .. 4096 px), fixed 10-op chain
[mean_shift, color_shift, brightness_shift, std_shift, gaussian_blur, gaussion_noise, line_streak,
 camera_cubic_curve, similarity_mls, rotate], every stage batched (PhotometricBatch, GeometricBatch
x2, AffineBatch), inputs resident.  One JSON line per size; pages/s and algorithmic GB/s (each
stage's input read once + output written once).

    python tools/bench_mixed.py [--sizes 256,512,1024,2048,4096] [--steps 3]
"""
import argparse
import json
import sys

import numpy as np
import torch

sys.path.insert(0, '/root/repo')
from vkit_b200.batch import AffineBatch, GeometricBatch, PhotometricBatch  # noqa: E402
from vkit_b200.mechanism.distortion_policy import random_distortion as rd  # noqa: E402

OPS = ['mean_shift', 'color_shift', 'brightness_shift', 'std_shift', 'gaussian_blur',
       'gaussion_noise', 'line_streak', 'camera_cubic_curve', 'similarity_mls', 'rotate']
PAGES = {256: 1024, 512: 512, 1024: 128, 2048: 32, 4096: 8}


def policies():
    out = {}
    for group in (rd._PHOTOMETRIC_POLICY_FACTORIES_AND_DEFAULT_WEIGHTS_SUM_PAIRS
                  + rd._GEOMETRIC_POLICY_FACTORIES_AND_DEFAULT_WEIGHTS_SUM_PAIRS):
        for fac in group[0]:
            if fac.name in OPS:
                out[fac.name] = fac.create()
    return out


def sample(n, size, pols):
    cfg = {name: [] for name in OPS}
    seeds = []
    for i, seq in enumerate(np.random.SeedSequence(133700 + size).spawn(n)):
        rng = np.random.default_rng(seq)
        level = int(rng.integers(1, 11))
        shape = (size, size)
        for name in OPS[:9]:
            pol = pols[name]
            gen = pol.config_generator_cls(pol.config_for_config_generator, level)
            cfg[name].append(gen(shape, rng))
        seeds.append(int(rng.integers(0, 2**63 - 1)))
    return cfg, seeds


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--sizes', default='256,512,1024,2048,4096')
    ap.add_argument('--steps', type=int, default=3)
    args = ap.parse_args()
    pols = policies()
    for size in [int(v) for v in args.sizes.split(',')]:
        n = PAGES.get(size, 16)
        cfg, seeds = sample(n, size, pols)
        shapes = [(size, size)] * n
        pages = torch.randint(0, 256, (n * size * size * 3,), dtype=torch.uint8, device='cuda')

        stage_ms = {}

        def mark(name, t0):
            torch.cuda.synchronize()
            import time as _t
            stage_ms[name] = stage_ms.get(name, 0.0) + (_t.perf_counter() - t0) * 1e3
            return _t.perf_counter()

        def step(profile=False):
            import time as _t
            # the same draws every step: result shapes repeat, so the caching allocator reuses
            # its blocks (fresh sizes would put a cudaMalloc of the 1.5 GB arenas in the loop)
            rot_rng = np.random.default_rng(size)
            t0 = _t.perf_counter()
            photo = PhotometricBatch(shapes, 3, [
                ('mean_shift', cfg['mean_shift']), ('color_shift', cfg['color_shift']),
                ('brightness_shift', cfg['brightness_shift']), ('std_shift', cfg['std_shift']),
                ('gaussian_blur', cfg['gaussian_blur']),
                ('gaussion_noise', cfg['gaussion_noise'], seeds),
                ('line_streak', cfg['line_streak'])])
            arena = photo.run(pages.clone())
            if profile:
                t0 = mark('photometric (7 ops, 5 passes)', t0)
            out1 = GeometricBatch(['camera_cubic_curve'] * n, cfg['camera_cubic_curve'],
                                  shapes).run(arena, channels=3)
            if profile:
                t0 = mark('camera_cubic_curve', t0)
            # the MLS / rotate configs depend on the (data dependent) input shapes of their stage
            pol = pols['similarity_mls']
            mls = [pol.config_generator_cls(pol.config_for_config_generator, 5)(s, rot_rng)
                   for s in out1.shapes]
            out2 = GeometricBatch(['similarity_mls'] * n, mls, out1.shapes).run(out1.image_arena,
                                                                                 channels=3)
            if profile:
                t0 = mark('similarity_mls (incl. config sampling)', t0)
            rot = [{'angle': int(rot_rng.integers(1, 360))} for _ in range(n)]
            out3 = AffineBatch(['rotate'] * n, rot, out2.shapes).run(out2.image_arena, channels=3)
            if profile:
                t0 = mark('rotate', t0)
            px = [n * size * size, sum(h * w for h, w in out1.shapes),
                  sum(h * w for h, w in out2.shapes), sum(h * w for h, w in out3.shapes)]
            return out3, px

        step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            out, px = step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        step(profile=True)
        # photometric passes: A (3 ops), stats, std pass, blur+..., noise  ~ 4.5 x (read + write);
        # geometric: read in + write out each
        alg = 3 * (px[0] * 9 + (px[0] + px[1]) + (px[1] + px[2]) + (px[2] + px[3]))
        print(json.dumps({'workload': f'config 5 chain, {n} pages of {size}x{size} RGB',
                          'pages_per_s': n / ms * 1e3, 'ms_per_step': ms,
                          'megapixels_per_s': px[0] / ms / 1e3,
                          'algorithmic_GBps': alg / ms / 1e6,
                          'final_pixels_over_input': px[3] / px[0],
                          'stage_ms_synchronised': {k: round(v, 2) for k, v in stage_ms.items()}}))
        del pages, out


if __name__ == '__main__':
    main()
