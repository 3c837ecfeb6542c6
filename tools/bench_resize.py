"""page_resizing on the device: the seven resamples of PageResizingStep.run (page image, four
masks, two height score maps; pipeline/text_detection/page_resizing.py:112-180) of a 2522 x 2522
page (the default page area, page_shape.py:26-28) for every interpolation the step samples.
Inputs are resident on the device; times are CUDA events over `--reps` repetitions after warm-up.
Prints one JSON line per interpolation: ms per page and algorithmic GB/s (bytes read + written).

    python tools/bench_resize.py [--side 2522] [--ratio 0.45] [--reps 20]
"""
import argparse
import json
import sys

import numpy as np
import torch

sys.path.insert(0, '/root/repo')
from vkit_b200 import compositing, element  # noqa: E402

CODES = {6: 'NEAREST_EXACT', 5: 'LINEAR_EXACT', 2: 'CUBIC', 4: 'LANCZOS4', 3: 'AREA'}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--side', type=int, default=2522)
    ap.add_argument('--ratio', type=float, default=0.45)
    ap.add_argument('--reps', type=int, default=20)
    args = ap.parse_args()
    rng = np.random.default_rng(7)
    side = args.side
    image = element.Image(mat=torch.from_numpy(rng.integers(0, 256, (side, side, 3), dtype=np.uint8)).cuda())
    masks = [element.Mask(mat=torch.from_numpy((rng.random((side, side)) > 0.5).astype(np.uint8)).cuda())
             for _ in range(4)]
    maps = [element.ScoreMap(mat=torch.from_numpy((rng.random((side, side)) * 40).astype(np.float32)).cuda(),
                             is_prob=False) for _ in range(2)]
    out_side = round(args.ratio * side)
    # image 3 B/px, 4 masks 1 B/px, 2 maps 4 B/px, read at the source size and written at the result size
    algorithmic = (3 + 4 + 8) * (side * side + out_side * out_side)
    for code, name in CODES.items():
        if code == 3 and args.ratio > 1:
            continue
        for _ in range(3):
            compositing.resize_page_elements(image, masks, maps, args.ratio, code)
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        start.record()
        for _ in range(args.reps):
            compositing.resize_page_elements(image, masks, maps, args.ratio, code)
        stop.record()
        torch.cuda.synchronize()
        ms = start.elapsed_time(stop) / args.reps
        print(json.dumps({'workload': f'page_resizing {side}x{side} -> {out_side}x{out_side}',
                          'interpolation': name, 'ms_per_page': round(ms, 3),
                          'pages_per_s': round(1e3 / ms, 1),
                          'algorithmic_GBps': round(algorithmic / ms / 1e6, 1)}))


if __name__ == '__main__':
    main()
