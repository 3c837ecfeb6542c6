"""Golden fixtures for the CHAINED configurations, from the LIVE reference (vkit-x/vkit @ 98ada2d
under /root/reference, cv2 4.13.0.92, numpy 2.3.5).  Run in the build container only:

    PYTHONPATH=/root/repo python tests/golden/make_golden_chain.py

  random_distortion   BASELINE config 4's distortion stage: random_distortion_factory with the
                      pipeline's flags (page_distortion.py:53-64 minus the ops that are "next"
                      rows), image + mask + points + polygons, one case per rng seed: the chosen
                      policy names / levels / configs, the result shape, hashes and arrays.
  fixed_chain         BASELINE config 5's 10-op chain at several page sizes: per-stage hashes,
                      final arrays for the smallest size.
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

from vkit.element import Image, Mask, Point, PointList, Polygon  # noqa: E402
from vkit.mechanism import distortion  # noqa: E402
from vkit.mechanism.distortion_policy.random_distortion import (  # noqa: E402
    RandomDistortionDebug, random_distortion_factory)

from common import make_inputs, make_points, make_polygons  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
NOT_YET = ['defocus_blur', 'zoom_in_blur', 'motion_blur', 'glass_blur', 'jpeg_quality',
           'pixelation', 'fog', 'ellipse_streak']
CHAIN_OPS = ['mean_shift', 'color_shift', 'brightness_shift', 'std_shift', 'gaussian_blur',
             'gaussion_noise', 'line_streak', 'camera_cubic_curve', 'similarity_mls', 'rotate']

CASES, ARRAYS = [], {}


def xy(points):
    return np.asarray([(p.smooth_x, p.smooth_y) for p in points], dtype=np.float64)


NOT_YET_2 = ['zoom_in_blur', 'jpeg_quality', 'ellipse_streak']


def random_distortion_cases(prefix='rd', disabled=NOT_YET, seeds=range(24)):
    shape = (160, 208)
    rd = random_distortion_factory.create({'disabled_policy_names': disabled,
                                           'force_post_rotate': True})
    for seed in seeds:
        image, mask, _ = make_inputs(1000 + seed, shape)
        pts = PointList(Point.create(y=y, x=x) for x, y in make_points(1000 + seed, shape, 16))
        polys = [Polygon.from_xy_pairs(p) for p in make_polygons(1000 + seed, shape, 4)]
        debug = RandomDistortionDebug()
        rng = np.random.default_rng(seed)
        r = rd.distort(rng, image=Image(mat=image), mask=Mask(mat=mask), points=pts,
                       polygons=polys, debug=debug)
        cid = f'{prefix}{seed:02d}'
        case = {
            'id': cid, 'kind': 'random_distortion', 'disabled': list(disabled), 'shape': list(shape), 'seed': 1000 + seed,
            'rng_seed': seed, 'names': list(debug.distortion_names),
            'levels': [int(v) for v in debug.distortion_levels],
            'configs': [mg.plain(c) for c in debug.distortion_configs],
            'result_shape': list(r.image.shape),
            'stage_sha': [mg.sha(im.mat) for im in debug.distortion_images],
            'sha': {'image': mg.sha(r.image.mat), 'mask': mg.sha(r.mask.mat)},
            'rng_after': float(rng.random()),
        }
        ARRAYS[f'{cid}/image'] = r.image.mat
        ARRAYS[f'{cid}/mask'] = r.mask.mat
        ARRAYS[f'{cid}/points'] = xy(r.points)
        ARRAYS[f'{cid}/polygons'] = np.asarray([xy(p.points) for p in r.polygons])
        CASES.append(case)


def fixed_chain_cases():
    for size, keep in ((256, True), (512, False), (1024, False)):
        shape = (size, size)
        seed = 5000 + size
        image, _, _ = make_inputs(seed, shape)
        rng = np.random.default_rng(seed)
        cur = Image(mat=image)
        stage_sha, configs, shapes = [], [], []
        for name in CHAIN_OPS:
            config = mg.policy_config(name, 6, cur.shape, int(rng.integers(0, 2**31)))
            op_rng = np.random.default_rng(int(rng.integers(0, 2**31)))
            r = getattr(distortion, name).distort(config, image=cur, rng=op_rng, get_config=True)
            cur = r.image
            configs.append(mg.plain(r.config))
            stage_sha.append(mg.sha(cur.mat))
            shapes.append(list(cur.shape))
        cid = f'fc{size}'
        case = {'id': cid, 'kind': 'fixed_chain', 'shape': list(shape), 'seed': seed,
                'ops': CHAIN_OPS, 'configs': configs, 'stage_sha': stage_sha,
                'stage_shapes': shapes, 'sha': {'image': mg.sha(cur.mat)}}
        if keep:
            ARRAYS[f'{cid}/image'] = cur.mat
        CASES.append(case)


def label_polygons(seed, shape, n, max_side):
    """n rotated boxes (as the distorted char / text-line polygons look), overlapping freely."""
    rng = np.random.default_rng(seed)
    height, width = shape
    polys = []
    for _ in range(n):
        cx, cy = rng.uniform(8, width - 8), rng.uniform(8, height - 8)
        hw, hh = rng.uniform(2, max_side), rng.uniform(2, max_side / 2)
        theta = rng.uniform(-0.6, 0.6)
        c, s = np.cos(theta), np.sin(theta)
        corners = [(-hw, -hh), (hw, -hh), (hw, hh), (-hw, hh)]
        xy = [(np.clip(cx + c * x - s * y, 0, width - 1), np.clip(cy + s * x + c * y, 0, height - 1))
              for x, y in corners]
        polys.append([(float(x), float(y)) for x, y in xy])
    return polys


def label_cases():
    """Label rasterisation after distortion (page_distortion.py:163-314): text-line mask, height
    score map in list order, combined char mask (keep max), char heights from tall to small."""
    from vkit.element import ScoreMap
    for cid, shape, n, max_side, seed in (('lb0', (160, 208), 40, 30, 70),
                                          ('lb1', (512, 640), 600, 14, 71),
                                          ('lb2', (300, 333), 25, 120, 72)):
        xy_lists = label_polygons(seed, shape, n, max_side)
        polys = [Polygon.from_xy_pairs(p) for p in xy_lists]
        rng = np.random.default_rng(seed + 1)
        heights = [float(v) for v in rng.uniform(4, 60, n).astype(np.float32)]
        mask = Mask.from_shape(shape)
        for poly in polys:
            poly.fill_mask(mask)
        height_map = ScoreMap.from_shape(shape, is_prob=False)
        for poly, hv in zip(polys, heights):
            poly.fill_score_map(score_map=height_map, value=hv)
        char_mask = Mask.from_shape(shape)
        for poly in polys:
            poly.fill_mask(char_mask, keep_max_value=True)
        order = list(reversed(np.asarray(heights).argsort()))
        char_heights = ScoreMap.from_shape(shape, is_prob=False)
        for idx in order:
            polys[idx].fill_score_map(score_map=char_heights, value=heights[idx])
        case = {'id': cid, 'kind': 'labels', 'shape': list(shape), 'seed': seed,
                'polygons': xy_lists, 'heights': heights,
                'sha': {'mask': mg.sha(mask.mat), 'height_map': mg.sha(height_map.mat),
                        'char_mask': mg.sha(char_mask.mat),
                        'char_heights': mg.sha(char_heights.mat)}}
        if cid == 'lb0':
            ARRAYS[f'{cid}/mask'] = mask.mat
            ARRAYS[f'{cid}/height_map'] = height_map.mat
        CASES.append(case)


def filter_blur_cases():
    """defocus_blur / motion_blur (cv.filter2D with a host-built float32 kernel)."""
    shape = (64, 96)
    specs = [('defocus_blur', {'radius': 1}), ('defocus_blur', {'radius': 2}),
             ('defocus_blur', {'radius': 3}), ('defocus_blur', {'radius': 4, 'anti_aliasing_sigma': 0.9}),
             ('defocus_blur', {'radius': 7}),  # 17 x 17 taps: cv2 switches to its DFT path
             ('motion_blur', {'radius': 2, 'angle': 30}), ('motion_blur', {'radius': 4, 'angle': 135}),
             ('motion_blur', {'radius': 3, 'angle': 300}), ('motion_blur', {'radius': 5, 'angle': 0}),
             ('motion_blur', {'radius': 1, 'angle': 77, 'anti_aliasing_sigma': 0.8})]
    for k, (name, cfg) in enumerate(specs):
        seed = 8000 + k
        image, _, _ = make_inputs(seed, shape)
        r = getattr(distortion, name).distort(cfg, image=Image(mat=image), get_config=True)
        cid = f'fb{k:02d}'
        CASES.append({'id': cid, 'kind': 'filter_blur', 'op': name, 'config': mg.plain(r.config),
                      'shape': list(shape), 'seed': seed, 'sha': {'image': mg.sha(r.image.mat)}})
        ARRAYS[f'{cid}/image'] = r.image.mat
    # one full-size page per op (hash only)
    for k, (name, cfg) in enumerate([('defocus_blur', {'radius': 3}),
                                     ('motion_blur', {'radius': 4, 'angle': 200})]):
        seed = 8100 + k
        image, _, _ = make_inputs(seed, (1024, 1024))
        r = getattr(distortion, name).distort(cfg, image=Image(mat=image), get_config=True)
        CASES.append({'id': f'fbL{k}', 'kind': 'filter_blur', 'op': name,
                      'config': mg.plain(r.config), 'shape': [1024, 1024], 'seed': seed,
                      'sha': {'image': mg.sha(r.image.mat)},
                      'sum': int(r.image.mat.astype(np.int64).sum())})


def effect_cases():
    """pixelation (cv.resize linear down + nearest up) and fog (diamond-square field + blend)."""
    specs = [('pixelation', {'ratio': 0.13}, (64, 96)), ('pixelation', {'ratio': 0.5}, (64, 96)),
             ('pixelation', {'ratio': 0.37}, (100, 133)), ('pixelation', {'ratio': 0.93}, (77, 50)),
             ('pixelation', {'ratio': 0.25}, (1024, 1024)),
             ('fog', {'roughness': 0.5}, (64, 96)), ('fog', {'roughness': 0.2, 'ratio_max': 0.7,
                                                               'ratio_min': 0.1}, (100, 133)),
             ('fog', {'roughness': 0.9, 'fog_rgb': [10, 20, 250]}, (257, 300)),
             ('fog', {'roughness': 0.6}, (1024, 1024)),
             ('glass_blur', {'sigma': 0.7}, (64, 96)),
             ('glass_blur', {'sigma': 1.2, 'delta': 2, 'loop': 3}, (100, 133)),
             ('glass_blur', {'sigma': 0.9, 'delta': 3, 'loop': 7}, (33, 47)),
             ('glass_blur', {'sigma': 0.8}, (1024, 1024))]
    for k, (name, cfg, shape) in enumerate(specs):
        seed = 8200 + k
        image, _, _ = make_inputs(seed, shape)
        rng = np.random.default_rng(seed + 1) if name in ('fog', 'glass_blur') else None
        r = getattr(distortion, name).distort(cfg, image=Image(mat=image), rng=rng, get_config=True)
        CASES.append({'id': f'ef{k:02d}', 'kind': 'effect', 'op': name, 'config': mg.plain(r.config),
                      'shape': list(shape), 'seed': seed, 'rng_seed': seed + 1 if rng else None,
                      'sha': {'image': mg.sha(r.image.mat)}})


def cubic_cases():
    """zoom_in_blur and Image.to_resized_image (default INTER_CUBIC): cv2 runs cubic in Intel IPP,
    so these are tolerance cases (+-1); arrays are kept."""
    for k, (cfg, shape) in enumerate([({}, (64, 96)), ({'ratio': 0.05, 'step': 0.02, 'alpha': 0.7},
                                                       (100, 133)), ({'ratio': 0.2, 'step': 0.05},
                                                                     (77, 50))]):
        seed = 8300 + k
        image, _, _ = make_inputs(seed, shape)
        r = distortion.zoom_in_blur.distort(cfg, image=Image(mat=image), get_config=True)
        cid = f'zb{k}'
        CASES.append({'id': cid, 'kind': 'cubic', 'op': 'zoom_in_blur', 'config': mg.plain(r.config),
                      'shape': list(shape), 'seed': seed, 'sha': {'image': mg.sha(r.image.mat)}})
        ARRAYS[f'{cid}/image'] = r.image.mat
    for k, (shape, dsize) in enumerate([((64, 96), (70, 101)), ((100, 133), (61, 200)),
                                        ((77, 50), (150, 33))]):
        seed = 8400 + k
        image, _, _ = make_inputs(seed, shape)
        out = Image(mat=image).to_resized_image(resized_height=dsize[0], resized_width=dsize[1])
        cid = f'rc{k}'
        CASES.append({'id': cid, 'kind': 'cubic', 'op': 'to_resized_image', 'shape': list(shape),
                      'resized': list(dsize), 'seed': seed, 'sha': {'image': mg.sha(out.mat)}})
        ARRAYS[f'{cid}/image'] = out.mat


def main():
    random_distortion_cases()
    fixed_chain_cases()
    label_cases()
    filter_blur_cases()
    effect_cases()
    random_distortion_cases('rx', NOT_YET_2, range(100, 124))
    cubic_cases()
    with open(os.path.join(HERE, 'chain_cases.json'), 'w') as fout:
        json.dump({'reference': 'vkit-x/vkit@98ada2d', 'cv2': __import__('cv2').__version__,
                   'numpy': np.__version__, 'cases': CASES}, fout, indent=1)
    np.savez_compressed(os.path.join(HERE, 'chain_arrays.npz'), **ARRAYS)
    print(len(CASES), 'cases;', sum(v.nbytes for v in ARRAYS.values()) / 1e6, 'MB raw arrays')
    for c in CASES:
        print(c['id'], c.get('names', c.get('ops', c['kind'])), c.get('result_shape') or (c.get('stage_shapes') or [c['shape']])[-1])


if __name__ == '__main__':
    main()
