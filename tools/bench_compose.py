"""BASELINE config 4 shape, per-page API (what vkit.pipeline drives): background 1024x1024 +
64 text lines of float32 coverage (32 x 768) blended in one launch, RandomDistortion (default
stages, force_post_rotate) on image + mask + 256 points + 64 text-line polygons, label
rasterisation of the distorted polygons (mask + height score map).  Prints one JSON line.

    python tools/bench_compose.py [--pages 64]
"""
import argparse
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, '/root/repo')
from vkit_b200 import element  # noqa: E402
from vkit_b200.compositing import assemble_text_lines, fill_polygons  # noqa: E402
from vkit_b200.mechanism.distortion_policy import random_distortion_factory  # noqa: E402

NOT_YET = ['zoom_in_blur', 'jpeg_quality', 'ellipse_streak']


def make_page(seed):
    rng = np.random.default_rng(seed)
    h = w = 1024
    background = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    lines, colors, polys = [], [], []
    for i in range(64):
        lh, lw = 32, 768
        up = 8 + (i % 32) * 31
        left = int(rng.integers(0, w - lw))
        alpha = np.clip(rng.normal(0.8, 0.3, (lh, lw)), 0, 1).astype(np.float32)
        alpha[rng.random((lh, lw)) > 0.4] = 0.0
        box = element.Box(up=up, down=up + lh - 1, left=left, right=left + lw - 1)
        lines.append(element.ScoreMap(mat=alpha, box=box))
        colors.append(tuple(int(v) for v in rng.integers(0, 256, 3)))
        polys.append(element.Polygon.from_xy_pairs(
            [(left, up), (left + lw - 1, up), (left + lw - 1, up + lh - 1), (left, up + lh - 1)]))
    points = element.PointList(element.Point.create(y=float(y), x=float(x))
                               for x, y in rng.uniform(0, 1000, (256, 2)))
    return background, lines, colors, polys, points


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--pages', type=int, default=64)
    args = ap.parse_args()
    rd = random_distortion_factory.create({'disabled_policy_names': NOT_YET,
                                           'force_post_rotate': True})
    pages = [make_page(1000 + i) for i in range(args.pages)]
    acc = {'compose': 0.0, 'distort': 0.0, 'labels': 0.0}

    def one(i, background, lines, colors, polys, points):
        t0 = time.perf_counter()
        image = assemble_text_lines(element.Image(mat=background), lines, colors)
        t1 = time.perf_counter()
        mask = element.Mask.from_shape(image.shape, value=1)
        r = rd.distort(np.random.default_rng(i), image=image, mask=mask, points=points,
                       polygons=polys)
        t2 = time.perf_counter()
        line_mask = fill_polygons(element.Mask.from_shape(r.image.shape), r.polygons, 1)
        heights = [float(10 + k % 30) for k in range(len(r.polygons))]
        height_map = fill_polygons(element.ScoreMap.from_shape(r.image.shape, is_prob=False),
                                   r.polygons, heights)
        t3 = time.perf_counter()
        acc['compose'] += t1 - t0
        acc['distort'] += t2 - t1
        acc['labels'] += t3 - t2
        return r.image, line_mask, height_map

    one(0, *pages[0])
    torch.cuda.synchronize()
    for k in acc:
        acc[k] = 0.0
    t0 = time.perf_counter()
    for i, page in enumerate(pages):
        out = one(i, *page)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    print(json.dumps({'workload': 'config 4 shape, per-page API: 64 text lines blended + '
                                  'RandomDistortion (image, mask, 256 points, 64 polygons) + label fills',
                      'pages': args.pages, 'pages_per_s': args.pages / wall,
                      'ms_per_page': wall / args.pages * 1e3,
                      'host_ms_per_page': {k: v / args.pages * 1e3 for k, v in acc.items()}}))


if __name__ == '__main__':
    main()
