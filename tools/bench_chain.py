"""BASELINE config 3: chained similarity_mls -> gaussian_blur -> color_shift over a batch of
1024x1024 RGB pages on one GPU (inputs resident).  Not the bench.py contract line -- a second
measured workload for DESIGN.md.  Prints one JSON line.

    python tools/bench_chain.py [--batch 1024] [--steps 5] [--warmup 3]
"""
import argparse
import json
import sys

import numpy as np
import torch

sys.path.insert(0, '/root/repo')
import bench  # noqa: E402
from vkit_b200.batch import GeometricBatch, PhotometricBatch  # noqa: E402


def sample(n):
    from vkit_b200.mechanism.distortion_policy.geometric.mls import similarity_mls_policy_factory
    from vkit_b200.mechanism.distortion_policy.photometric.blur import gaussian_blur_policy_factory
    from vkit_b200.mechanism.distortion_policy.photometric.color import color_shift_policy_factory
    pols = [f.create() for f in (similarity_mls_policy_factory, gaussian_blur_policy_factory,
                                 color_shift_policy_factory)]
    out = [[], [], []]
    for rng in bench.page_rngs(0, n, n):
        level = int(rng.integers(1, 11))
        for k, pol in enumerate(pols):
            gen = pol.config_generator_cls(pol.config_for_config_generator, level)
            out[k].append(gen(bench.PAGE_SHAPE, rng))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=1024)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    args = ap.parse_args()
    n = args.batch
    mls_cfg, blur_cfg, color_cfg = sample(n)
    pages = torch.randint(0, 256, (n,) + bench.PAGE_SHAPE + (3,), dtype=torch.uint8, device='cuda')
    names = ['similarity_mls'] * n
    scratch = [None]
    kernel_events = []
    remap_events = []

    def step(events=None):
        engine = GeometricBatch(names, mls_cfg, bench.PAGE_SHAPE)
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        out = engine.run(pages, launch_events=remap_events if events is not None else None)
        e[1].record()
        photo = PhotometricBatch(out.shapes, 3, [('gaussian_blur', blur_cfg),
                                                  ('color_shift', color_cfg)])
        if events is not None:
            photo.launch_events = kernel_events
        if scratch[0] is None or scratch[0].numel() != out.image_arena.numel():
            scratch[0] = torch.empty_like(out.image_arena)
        res = photo.run(out.image_arena, scratch[0])
        e[2].record()
        if events is not None:
            events.append(e)
        return out, res

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    events = []
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(args.steps):
        out, res = step(events)
    stop.record()
    torch.cuda.synchronize()
    ms = start.elapsed_time(stop) / args.steps
    geo_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in events]))
    photo_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in events]))
    photo_kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in kernel_events]))
    remap_kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in remap_events]))
    dst_px = int(res.numel()) // 3
    src_px = n * bench.PAGE_SHAPE[0] * bench.PAGE_SHAPE[1]
    peak = bench.measured_peak()[0] if hasattr(bench, 'measured_peak') else 6552.3
    print(json.dumps({
        'workload': 'config 3: similarity_mls -> gaussian_blur -> color_shift, 1024x1024 RGB, '
                    f'batch {n}, inputs resident, configs from the policy generators',
        'pages_per_s': n / ms * 1e3, 'ms_per_step': ms,
        'geometric_ms': geo_ms, 'photometric_ms': photo_ms,
        'remap_kernel_ms': remap_kernel_ms, 'photo_kernel_ms': photo_kernel_ms,
        'fused_photo_GBps': 6.0 * dst_px / photo_kernel_ms / 1e6,
        'chain_algorithmic_GBps': 3.0 * (src_px + dst_px) / ms / 1e6,
        'hbm_peak_GBps': peak,
    }))


if __name__ == '__main__':
    main()
