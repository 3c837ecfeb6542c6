"""gaussian_blur with and without the TMA box load of the tile + halo (VKB_BLUR_NO_TMA=1 selects
the per-thread staging loop): identical output, time per launch by CUDA events.

    python tools/blur_tma_probe.py            # timing table
    ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum ... python tools/blur_tma_probe.py --once
"""
import argparse
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vkit_b200 import _native as nv  # noqa: E402
from vkit_b200 import device as dv  # noqa: E402
from vkit_b200.mechanism.distortion.photometric.blur import gaussian_kernel_u8  # noqa: E402


def blur(src, dst, taps):
    h, w, c = src.shape
    arr = (ctypes.c_int32 * len(taps))(*taps)
    nv.check(nv.lib().vkb_gaussian_blur_u8(dv.ptr(src), dv.ptr(dst), h, w, c, arr, len(taps),
                                           dv.stream_ptr()), 'vkb_gaussian_blur_u8')


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument('--once', action='store_true')
    args = parser.parse_args()
    rng = np.random.default_rng(0)
    for side in (1024, 2048, 4096):
        src = dv.to_device(rng.integers(0, 256, (side, side, 3), dtype=np.uint8))
        out = [torch.empty_like(src), torch.empty_like(src)]
        for ksize, sigma in ((5, 1.0), (9, 2.5), (17, 5.0)):
            taps = gaussian_kernel_u8(ksize, sigma)
            times = []
            for k, flag in enumerate(('0', '1')):
                os.environ['VKB_BLUR_NO_TMA'] = flag
                reps = 1 if args.once else 50
                for _ in range(0 if args.once else 5):
                    blur(src, out[k], taps)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    blur(src, out[k], taps)
                e1.record()
                torch.cuda.synchronize()
                times.append(e0.elapsed_time(e1) * 1e3 / reps)
            same = bool(torch.equal(out[0], out[1]))
            gbs = 2 * 3 * side * side / (times[0] * 1e-6) / 1e9
            print(f'{side}x{side} ksize {ksize:2d}: TMA {times[0]:8.1f} us ({gbs:6.0f} GB/s algorithmic)  '
                  f'loop {times[1]:8.1f} us  ratio {times[1] / times[0]:.2f}  identical {same}')
            assert same
    os.environ.pop('VKB_BLUR_NO_TMA', None)


if __name__ == '__main__':
    main()
