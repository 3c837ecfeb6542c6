"""Generates tests/golden/cases.json + arrays.npz by running the LIVE reference
(vkit-x/vkit @ 98ada2d under /root/reference, cv2 4.13.0.92, numpy 2.3.5) on seeded synthetic
inputs.  Run in the build container only:

    PYTHONPATH=/root/repo python tests/golden/make_golden.py

Every case stores: op name, config (plain dict), input seed + shape, output shape, sha256 of each
output array, and -- for small cases -- the arrays themselves.  Inputs are regenerated from the
seed by `tests/common.py:make_inputs`, never stored.
"""
import hashlib
import json
import os
import sys

import attrs
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from oracle import refshim  # noqa: E402

vkit = refshim.load()

from vkit.element import Image, ImageMode, Mask, Point, PointList, Polygon, ScoreMap  # noqa: E402
from vkit.mechanism import distortion  # noqa: E402
from vkit.mechanism.distortion_policy.random_distortion import random_distortion_factory  # noqa: E402

from common import make_inputs, make_points, make_polygons  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def sha(arr: np.ndarray) -> str:
    arr = np.ascontiguousarray(arr)
    return hashlib.sha256(arr.tobytes()).hexdigest()


def plain(obj):
    """Reference config -> JSON-able structure."""
    if attrs.has(type(obj)):
        out = {}
        for field in attrs.fields(type(obj)):
            if field.name == '_rng_state':
                continue
            out[field.name.lstrip('_')] = plain(getattr(obj, field.name))
        return out
    if isinstance(obj, Point):
        return [obj.smooth_x, obj.smooth_y]
    if isinstance(obj, (list, tuple)):
        if len(obj) and isinstance(obj[0], Point):
            return [[p.smooth_x, p.smooth_y] for p in obj]
        return [plain(x) for x in obj]
    if hasattr(obj, 'value') and type(obj).__module__.startswith('vkit'):
        return obj.value
    if isinstance(obj, np.integer):
        return int(obj)
    if isinstance(obj, np.floating):
        return float(obj)
    return obj


CASES = []
ARRAYS = {}


def add_case(case, arrays, keep_arrays):
    cid = f"c{len(CASES):03d}"
    case['id'] = cid
    case['sha'] = {k: sha(v) for k, v in arrays.items()}
    case['has_arrays'] = bool(keep_arrays)
    if keep_arrays:
        for k, v in arrays.items():
            ARRAYS[f'{cid}/{k}'] = v
    CASES.append(case)


def geometric_case(name, config, shape, seed, keep_arrays, with_labels=True):
    image, mask, score_map = make_inputs(seed, shape)
    op = getattr(distortion, name)
    kwargs = {}
    if with_labels:
        pts = make_points(seed, shape, 24)
        polys = make_polygons(seed, shape, 6)
        kwargs['points'] = PointList(Point.create(y=y, x=x) for x, y in pts)
        kwargs['polygons'] = [Polygon.from_xy_pairs(p) for p in polys]
    r = op.distort(config, image=Image(mat=image), mask=Mask(mat=mask),
                   score_map=ScoreMap(mat=score_map), get_active_mask=True, get_config=True,
                   get_state=True, disable_clip_result_elements=True, **kwargs)
    arrays = {
        'image': r.image.mat,
        'mask': r.mask.mat,
        'score_map': r.score_map.mat,
        'active_mask': r.active_mask.mat,
    }
    if with_labels:
        arrays['points'] = np.asarray([(p.smooth_x, p.smooth_y) for p in r.points], dtype=np.float64)
        arrays['polygons'] = np.asarray([[(p.smooth_x, p.smooth_y) for p in poly.points]
                                         for poly in r.polygons], dtype=np.float64)
    small = {}
    if hasattr(r.state, 'dst_image_grid'):
        lat = np.asarray([[p.x, p.y] for p in r.state.dst_image_grid.flatten_points], dtype=np.int32)
        small['lattice'] = lat
    case = {'kind': 'geometric', 'op': name, 'config': plain(r.config), 'shape': list(shape),
            'seed': seed, 'result_shape': list(r.shape), 'labels': with_labels}
    cid_arrays = dict(arrays)
    cid_arrays.update(small)
    # lattices are always stored (small); big pixel arrays only for small cases
    # pixel arrays are pinned by sha256 only; lattices / points / polygons are stored (small)
    add_case(case, cid_arrays, False)
    if name.startswith('skew'):
        # skew ops are compared with a tie tolerance, so the pixels are stored
        for key in ('image', 'mask', 'score_map'):
            ARRAYS[f"{case['id']}/{key}"] = arrays[key]
    if 'lattice' in small:
        ARRAYS[f"{case['id']}/lattice"] = small['lattice']
    if with_labels:
        ARRAYS[f"{case['id']}/points"] = arrays['points']
        ARRAYS[f"{case['id']}/polygons"] = arrays['polygons']


def policy_config(factory_name, level, shape, seed):
    from vkit.mechanism.distortion_policy import random_distortion as rd
    for group in (rd._PHOTOMETRIC_POLICY_FACTORIES_AND_DEFAULT_WEIGHTS_SUM_PAIRS
                  + rd._GEOMETRIC_POLICY_FACTORIES_AND_DEFAULT_WEIGHTS_SUM_PAIRS):
        for fac in group[0]:
            if fac.name == factory_name:
                pol = fac.create()
                gen = pol.config_generator_cls(pol.config_for_config_generator, level)
                return gen(shape, np.random.default_rng(seed))
    raise KeyError(factory_name)


def photometric_case(name, config, shape, seed, keep_arrays, mode=None, rng_seed=None):
    image, _, _ = make_inputs(seed, shape)
    op = getattr(distortion, name)
    img = Image(mat=image)
    if mode:
        img = img.to_target_mode_image(ImageMode(mode))
    rng = np.random.default_rng(rng_seed) if rng_seed is not None else None
    r = op.distort(config, image=img, rng=rng, get_config=True)
    case = {'kind': 'photometric', 'op': name, 'config': plain(r.config), 'shape': list(shape),
            'seed': seed, 'mode': mode, 'rng_seed': rng_seed, 'result_mode': r.image.mode.value}
    tolerant = name in ('color_shift', 'brightness_shift', 'gaussion_noise', 'poisson_noise',
                        'impulse_noise', 'speckle_noise', 'std_shift')
    add_case(case, {'image': r.image.mat}, keep_arrays and tolerant)


def blend_cases():
    from vkit.element import Box
    shape = (72, 120)
    for idx, spec in enumerate([
        dict(kind='score_map_color', box=(10, 41, 20, 99)),
        dict(kind='box_alpha_scalar', box=(0, 71, 0, 119), alpha=0.3),
        dict(kind='box_value_image_alpha', box=(5, 60, 7, 100), alpha=0.65),
        dict(kind='mask_assign', box=(8, 50, 30, 90)),
        dict(kind='score_keep_max', box=(3, 40, 3, 80)),
        dict(kind='inactive_fill', box=(0, 71, 0, 119)),
    ]):
        seed = 900 + idx
        image, mask, score_map = make_inputs(seed, shape)
        rng = np.random.default_rng(seed + 5000)
        up, down, left, right = spec['box']
        box = Box(up=up, down=down, left=left, right=right)
        bh, bw = box.height, box.width
        img = Image(mat=image.copy())
        arrays = {}
        if spec['kind'] == 'score_map_color':
            alpha = rng.random((bh, bw)).astype(np.float32)
            alpha[rng.random((bh, bw)) < 0.6] = 0.0
            sm = ScoreMap(mat=alpha, box=box)
            sm.fill_image(img, (17, 99, 201))
            arrays['alpha'] = alpha
        elif spec['kind'] == 'box_alpha_scalar':
            box.fill_image(img, (250, 3, 77), alpha=spec['alpha'])
        elif spec['kind'] == 'box_value_image_alpha':
            value = rng.integers(0, 256, (bh, bw, 3), dtype=np.uint8)
            box.fill_image(img, value, alpha=spec['alpha'])
            arrays['value'] = value
        elif spec['kind'] == 'mask_assign':
            m = (rng.random((bh, bw)) > 0.5).astype(np.uint8)
            Mask(mat=m, box=box).fill_image(img, (1, 2, 3))
            arrays['m'] = m
        elif spec['kind'] == 'score_keep_max':
            base = ScoreMap(mat=score_map.copy())
            value = rng.random((bh, bw)).astype(np.float32)
            box.fill_score_map(base, value, keep_max_value=True)
            arrays['value'] = value
            arrays['out_score'] = base.mat
        elif spec['kind'] == 'inactive_fill':
            bottom = rng.integers(0, 256, image.shape, dtype=np.uint8)
            Mask(mat=mask.copy()).to_inverted_mask().fill_image(img, Image(mat=bottom))
            arrays['bottom'] = bottom
        arrays['out_image'] = img.mat
        case = {'kind': 'blend', 'op': spec['kind'], 'shape': list(shape), 'seed': seed,
                'box': list(spec['box']), 'alpha': spec.get('alpha')}
        add_case(case, arrays, True)


def main():
    small = (136, 176)
    # ---- affine family -------------------------------------------------------------------
    for name, cfg in [('rotate', {'angle': 30}), ('rotate', {'angle': 137}),
                      ('rotate', {'angle': 200}), ('rotate', {'angle': 301}),
                      ('rotate', {'angle': 0}),
                      ('shear_hori', {'angle': -20}), ('shear_hori', {'angle': 13}),
                      ('shear_vert', {'angle': 17}), ('shear_vert', {'angle': -9}),
                      ('skew_hori', {'ratio': 0.3}), ('skew_hori', {'ratio': -0.12}),
                      ('skew_vert', {'ratio': -0.2}), ('skew_vert', {'ratio': 0.33})]:
        geometric_case(name, cfg, small, 11, keep_arrays=True)
    # BASELINE config 1: 512x512 rotate 30 (hash only)
    geometric_case('rotate', {'angle': 30}, (512, 512), 133700, keep_arrays=False)
    geometric_case('shear_vert', {'angle': 11}, (1024, 1024), 133701, keep_arrays=False)

    # ---- grid ops: small with arrays, 1024^2 with hashes + lattices ------------------------
    grid_ops = ['camera_plane_only', 'camera_cubic_curve', 'camera_plane_line_fold',
                'camera_plane_line_curve', 'similarity_mls']
    for i, name in enumerate(grid_ops):
        for level, seed in ((3, 21 + i), (9, 41 + i)):
            cfg = policy_config(name, level, small, seed)
            geometric_case(name, cfg, small, seed, keep_arrays=True)
    for i, name in enumerate(grid_ops):
        cfg = policy_config(name, 7, (1024, 1024), 61 + i)
        geometric_case(name, cfg, (1024, 1024), 61 + i, keep_arrays=False)
    # non-square page, resize_as_src for MLS
    cfg = policy_config('similarity_mls', 6, (200, 333), 77)
    cfg.resize_as_src = True
    geometric_case('similarity_mls', cfg, (200, 333), 77, keep_arrays=True, with_labels=False)
    cfg = policy_config('camera_cubic_curve', 10, (333, 200), 78)
    geometric_case('camera_cubic_curve', cfg, (333, 200), 78, keep_arrays=True)

    # ---- photometric -----------------------------------------------------------------------
    pshape = (64, 96)
    photometric = [
        ('mean_shift', {'delta': 40}), ('mean_shift', {'delta': -77, 'threshold': 100}),
        ('mean_shift', {'delta': 60, 'threshold': 170, 'channels': [0, 2]}),
        ('mean_shift', {'delta': 200, 'oob_behavior': 'cycle', 'channels': [1]}),
        ('color_shift', {'delta': 37}), ('color_shift', {'delta': -120}),
        ('brightness_shift', {'delta': 45}), ('brightness_shift', {'delta': -90}),
        ('brightness_shift', {'delta': 30, 'intermediate_image_mode': 'hsv'}),
        ('std_shift', {'scale': 1.7}), ('std_shift', {'scale': 0.55, 'channels': [1]}),
        ('boundary_equalization', {}), ('boundary_equalization', {'channels': [0, 1]}),
        ('complement', {}), ('complement', {'threshold': 100, 'enable_threshold_lte': True}),
        ('complement', {'threshold': 150, 'channels': [2]}),
        ('posterization', {'num_bits': 3}), ('posterization', {'num_bits': 6, 'channels': [0]}),
        ('color_balance', {'ratio': 0.35}), ('color_balance', {'ratio': 0.9}),
        ('gaussian_blur', {'sigma': 0.5}), ('gaussian_blur', {'sigma': 0.8}),
        ('gaussian_blur', {'sigma': 1.0}), ('gaussian_blur', {'sigma': 2.0}),
        ('line_streak', {'thickness': 2, 'gap': 7, 'alpha': 0.45, 'color': [10, 200, 30]}),
        ('line_streak', {'thickness': 1, 'gap': 5, 'dash_thickness': 4, 'dash_gap': 3,
                         'alpha': 1.0, 'enable_hori': False}),
        ('line_streak', {'thickness': 3, 'gap': 9, 'dash_thickness': 5, 'dash_gap': 2,
                         'alpha': 0.7}),
        ('rectangle_streak', {'thickness': 2, 'short_side_min': 8, 'short_side_step': 9,
                              'alpha': 0.6, 'color': [255, 0, 9]}),
        ('rectangle_streak', {'thickness': 1, 'aspect_ratio': 0.7, 'dash_thickness': 3,
                              'dash_gap': 2, 'short_side_min': 6, 'short_side_step': 7,
                              'alpha': 1.0}),
    ]
    for name, cfg in photometric:
        photometric_case(name, cfg, pshape, 300 + len(CASES), keep_arrays=True)
    photometric_case('gaussian_blur', {'sigma': 0.7}, pshape, 401, keep_arrays=True, mode='grayscale')
    photometric_case('mean_shift', {'delta': 33}, pshape, 402, keep_arrays=True, mode='grayscale')
    photometric_case('channel_permutation', {}, pshape, 403, keep_arrays=True, rng_seed=0)
    photometric_case('gaussion_noise', {'std': 12.0}, pshape, 404, keep_arrays=True, rng_seed=5)
    photometric_case('poisson_noise', {}, pshape, 405, keep_arrays=True, rng_seed=6)
    photometric_case('impulse_noise', {'prob_salt': 0.03, 'prob_pepper': 0.02}, pshape, 406,
                     keep_arrays=True, rng_seed=7)
    photometric_case('speckle_noise', {'std': 0.2}, pshape, 407, keep_arrays=True, rng_seed=8)
    # 1024^2 hashes for the exact ops (config 3 ingredients)
    for name, cfg in [('gaussian_blur', {'sigma': 0.9}), ('mean_shift', {'delta': 50}),
                      ('complement', {}), ('posterization', {'num_bits': 4})]:
        photometric_case(name, cfg, (1024, 1024), 500 + len(CASES), keep_arrays=False)

    blend_cases()

    # appended later (keeps the ids / seeds of the cases above stable)
    photometric_case('histogram_equalization', {}, (64, 96), 601, keep_arrays=False)
    photometric_case('histogram_equalization', {'channels': [1]}, (64, 96), 602, keep_arrays=False)
    photometric_case('histogram_equalization', {}, (1024, 1024), 603, keep_arrays=False)
    photometric_case('histogram_equalization', {}, (64, 96), 604, keep_arrays=False,
                     mode='grayscale')

    with open(os.path.join(HERE, 'cases.json'), 'w') as fout:
        json.dump({'reference': 'vkit-x/vkit@98ada2d', 'cv2': __import__('cv2').__version__,
                   'numpy': np.__version__, 'cases': CASES}, fout, indent=1)
    np.savez_compressed(os.path.join(HERE, 'arrays.npz'), **ARRAYS)
    print(len(CASES), 'cases;', sum(v.nbytes for v in ARRAYS.values()) / 1e6, 'MB raw arrays')


if __name__ == '__main__':
    main()
