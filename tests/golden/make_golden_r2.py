"""Golden fixtures added in round 2, from the LIVE reference (vkit-x/vkit @ 98ada2d under
/root/reference, cv2 4.13.0.92, numpy 2.3.5).  Run in the build container only:

    PYTHONPATH=/root/repo python tests/golden/make_golden_r2.py

  ellipse_streak   the reference op with ONE change: under cv2 >= 4.13 `cv.ellipse(mask.mat, ...)`
                   raises because Mask.mat is a read-only array (streak.py:316), so the generator
                   lets cv.ellipse draw into a writable copy (the op's arithmetic is untouched).
  fog_gray         fog on a GRAYSCALE page (fractional fog value, effect.py:194-197).
  jpeg_quality     cv.imencode / cv.imdecode round trips (effect.py:26-55).
Inputs are regenerated from the seed by tests/common.py, never stored.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))

import make_golden as mg  # noqa: E402  (loads the reference through oracle/refshim)

from vkit.element import Image, ImageMode, Mask  # noqa: E402
from vkit.mechanism import distortion  # noqa: E402

from common import make_inputs  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CASES, ARRAYS = [], {}


class _CvWithWritableEllipse:
    """cv2 as the reference's streak module sees it, with ONE change: cv.ellipse draws into a
    writable copy of the (read-only) mask array it is handed and returns it -- the reference
    assigns the returned array back into the mask (streak.py:314-326), so the op's arithmetic is
    what it would be with a cv2 that accepted the read-only array."""

    def __init__(self, cv):
        self._cv = cv

    def __getattr__(self, name):
        return getattr(self._cv, name)

    def ellipse(self, img, **kwargs):
        return self._cv.ellipse(img.copy(), **kwargs)


def ellipse_cases():
    from vkit.mechanism.distortion.photometric import streak as ref_streak
    real_cv = ref_streak.cv
    ref_streak.cv = _CvWithWritableEllipse(real_cv)
    try:
        specs = [
            ({'thickness': 1}, (64, 96)),
            ({'thickness': 2, 'short_side_min': 7, 'short_side_step': 9, 'alpha': 0.6,
              'color': [255, 0, 9]}, (64, 96)),
            ({'thickness': 3, 'aspect_ratio': 0.7, 'short_side_min': 6, 'short_side_step': 11,
              'alpha': 0.35, 'color': [10, 200, 30]}, (100, 133)),
            ({'thickness': 1, 'aspect_ratio': 1.4, 'short_side_min': 5, 'short_side_step': 4,
              'alpha': 1.0}, (77, 50)),
            ({'thickness': 2, 'aspect_ratio': 0.5, 'short_side_min': 12, 'short_side_step': 30,
              'alpha': 0.8}, (257, 300)),
            ({'thickness': 3, 'short_side_min': 40, 'short_side_step': 57, 'alpha': 0.5,
              'color': [0, 0, 255]}, (1024, 1024)),
            ({'thickness': 1, 'aspect_ratio': 1.5, 'short_side_min': 25, 'short_side_step': 21,
              'alpha': 0.9}, (1024, 1024)),
        ]
        for k, (cfg, shape) in enumerate(specs):
            seed = 9100 + k
            image, _, _ = make_inputs(seed, shape)
            r = distortion.ellipse_streak.distort(cfg, image=Image(mat=image), get_config=True)
            CASES.append({'id': f'el{k:02d}', 'kind': 'ellipse_streak', 'op': 'ellipse_streak',
                          'config': mg.plain(r.config), 'shape': list(shape), 'seed': seed,
                          'sha': {'image': mg.sha(r.image.mat)}})
    finally:
        ref_streak.cv = real_cv


def fog_gray_cases():
    for k, (cfg, shape) in enumerate([({'roughness': 0.5}, (64, 96)),
                                      ({'roughness': 0.8, 'ratio_max': 0.9, 'ratio_min': 0.2,
                                        'fog_rgb': [200, 30, 99]}, (100, 133))]):
        seed = 9200 + k
        image, _, _ = make_inputs(seed, shape)
        gray = Image(mat=image).to_target_mode_image(ImageMode.GRAYSCALE)
        rng = np.random.default_rng(seed + 1)
        r = distortion.fog.distort(cfg, image=gray, rng=rng, get_config=True)
        CASES.append({'id': f'fg{k:02d}', 'kind': 'fog_gray', 'op': 'fog', 'config': mg.plain(r.config),
                      'shape': list(shape), 'seed': seed, 'rng_seed': seed + 1,
                      'sha': {'image': mg.sha(r.image.mat)}})


def jpeg_cases():
    specs = [(95, (64, 96)), (75, (64, 96)), (50, (100, 133)), (30, (77, 50)), (10, (64, 96)),
             (88, (257, 300)), (60, (1024, 1024))]
    for k, (quality, shape) in enumerate(specs):
        seed = 9300 + k
        image, _, _ = make_inputs(seed, shape)
        if k % 2 == 1:
            # smooth content next to the noise pages: blurred noise compresses like a photo
            import cv2
            image = cv2.GaussianBlur(image, (0, 0), 3.0)
        r = distortion.jpeg_quality.distort({'quality': quality}, image=Image(mat=image))
        cid = f'jq{k:02d}'
        CASES.append({'id': cid, 'kind': 'jpeg_quality', 'op': 'jpeg_quality',
                      'config': {'quality': quality}, 'shape': list(shape), 'seed': seed,
                      'smooth': k % 2 == 1, 'sha': {'image': mg.sha(r.image.mat)}})
        if shape[0] * shape[1] <= 100000:
            ARRAYS[f'{cid}/image'] = r.image.mat


def main():
    ellipse_cases()
    fog_gray_cases()
    jpeg_cases()
    with open(os.path.join(HERE, 'r2_cases.json'), 'w') as fout:
        json.dump({'reference': 'vkit-x/vkit@98ada2d', 'cv2': __import__('cv2').__version__,
                   'numpy': np.__version__, 'cases': CASES}, fout, indent=1)
    np.savez_compressed(os.path.join(HERE, 'r2_arrays.npz'), **ARRAYS)
    print(len(CASES), 'cases;', sum(v.nbytes for v in ARRAYS.values()) / 1e6, 'MB raw arrays')


if __name__ == '__main__':
    main()
