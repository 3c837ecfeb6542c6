"""Shared test helpers: seeded synthetic inputs, golden fixtures."""
import hashlib
import json
import os

import numpy as np

TESTS_DIR = os.path.dirname(os.path.abspath(__file__))
GOLDEN_DIR = os.path.join(TESTS_DIR, 'golden')


def make_inputs(seed: int, shape):
    """image uint8 HxWx3, mask uint8 HxW (0/1), score_map float32 HxW in [0, 1)."""
    height, width = shape
    rng = np.random.default_rng(seed)
    image = rng.integers(0, 256, (height, width, 3), dtype=np.uint8)
    mask = (rng.random((height, width)) > 0.5).astype(np.uint8)
    score_map = rng.random((height, width)).astype(np.float32)
    return image, mask, score_map


def make_points(seed: int, shape, n: int):
    """n smooth (x, y) points strictly inside the page."""
    height, width = shape
    rng = np.random.default_rng(seed + 100000)
    xs = rng.uniform(0, width - 2, n)
    ys = rng.uniform(0, height - 2, n)
    return [(float(x), float(y)) for x, y in zip(xs, ys)]


def make_polygons(seed: int, shape, n: int):
    """n axis-aligned quads as lists of (x, y)."""
    height, width = shape
    rng = np.random.default_rng(seed + 200000)
    polys = []
    for _ in range(n):
        x0 = int(rng.integers(0, width - 12))
        y0 = int(rng.integers(0, height - 12))
        x1 = int(rng.integers(x0 + 4, min(width - 1, x0 + 60)))
        y1 = int(rng.integers(y0 + 4, min(height - 1, y0 + 40)))
        polys.append([(x0, y0), (x1, y0), (x1, y1), (x0, y1)])
    return polys


def sha(arr: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(arr).tobytes()).hexdigest()


_GOLDEN = None


def golden():
    """(cases list, arrays NpzFile) from tests/golden."""
    global _GOLDEN
    if _GOLDEN is None:
        with open(os.path.join(GOLDEN_DIR, 'cases.json')) as fin:
            meta = json.load(fin)
        arrays = np.load(os.path.join(GOLDEN_DIR, 'arrays.npz'))
        _GOLDEN = (meta['cases'], arrays)
    return _GOLDEN


def golden_cases(kind=None, **filters):
    cases, _ = golden()
    out = []
    for case in cases:
        if kind and case['kind'] != kind:
            continue
        if any(case.get(k) != v for k, v in filters.items()):
            continue
        out.append(case)
    return out


def golden_array(case, key):
    _, arrays = golden()
    name = f"{case['id']}/{key}"
    return arrays[name] if name in arrays.files else None


# ---------------------------------------------------------------------------------------------
# Golden case -> oracle run / product config
# ---------------------------------------------------------------------------------------------
AFFINE_OPS = ('rotate', 'shear_hori', 'shear_vert', 'skew_hori', 'skew_vert')


# skew_hori / skew_vert: the 3x3 matrix comes from cv.getPerspectiveTransform(DECOMP_SVD), whose
# LAPACK solve differs from any closed form in the last bits; integer page corners put many source
# coordinates EXACTLY on 1/64 px rounding ties, where those bits decide (DESIGN.md "skew ties").
# A flipped tie moves one tap by one 1/32 px step, so per container:
#   image     <= 0.5 % of the pixels, each by at most 16 levels (random-noise content),
#   mask      <= 0.5 % of the pixels, values stay in {0, 1},
#   score_map <= 0.5 % of the pixels, each by at most 2/32 (one step per axis on values in [0, 1)).
SKEW_TIE_FRACTION = 0.005


def assert_skew_close(case, got, where=''):
    """got: {'image': ..., 'mask': ..., 'score_map': ...} arrays of one skew case."""
    for key, bound in (('image', 16), ('mask', 1), ('score_map', 2.0 / 32 + 1e-6)):
        ref = golden_array(case, key)
        assert got[key].shape == ref.shape and got[key].dtype == ref.dtype, (where, case['id'], key)
        diff = np.abs(got[key].astype(np.float64) - ref.astype(np.float64))
        if diff.ndim == 3:
            diff = diff.max(axis=-1)
        frac = float((diff > 0).mean())
        assert frac <= SKEW_TIE_FRACTION and diff.max() <= bound, (
            f"{where}{case['id']} {key}: {frac:.5f} of the pixels differ, max abs {diff.max():.4g}")
        if key == 'mask':
            assert set(np.unique(got[key]).tolist()) <= {0, 1}, (where, case['id'], 'mask values')


def oracle_geometric(case, port, want=('image', 'mask', 'score_map'), given_lattice=None):
    """Run the oracle on a golden geometric case -> dict of arrays."""
    shape = tuple(case['shape'])
    image, mask, score_map = make_inputs(case['seed'], shape)
    name, cfg = case['op'], case['config']
    mats = {'image': image, 'mask': mask, 'score_map': score_map}
    if name in AFFINE_OPS:
        trans_mat, dsize = port.affine_state(name, cfg, shape)
        out = {k: port.affine_apply(mats[k], trans_mat, dsize) for k in want}
        out['shape'] = (dsize[1], dsize[0]) if dsize else shape
        out['trans_mat'] = trans_mat
        return out
    res = port.grid_distort(name, cfg, shape, **{k: mats[k] for k in want},
                            given_lattice=given_lattice)
    return res


def product_config(case):
    """Config argument for vkit_b200.mechanism.distortion.<op>.distort from a golden case."""
    cfg = dict(case['config'])
    if case['op'] == 'similarity_mls':
        from vkit_b200.element import PointTuple
        cfg['src_handle_points'] = PointTuple.from_xy_pairs(cfg['src_handle_points'])
        cfg['dst_handle_points'] = PointTuple.from_xy_pairs(cfg['dst_handle_points'])
    return cfg


# ---------------------------------------------------------------------------------------------
# Chained fixtures (tests/golden/make_golden_chain.py)
# ---------------------------------------------------------------------------------------------
_CHAIN = None


def chain_golden():
    global _CHAIN
    if _CHAIN is None:
        with open(os.path.join(GOLDEN_DIR, 'chain_cases.json')) as fin:
            meta = json.load(fin)
        _CHAIN = (meta['cases'], np.load(os.path.join(GOLDEN_DIR, 'chain_arrays.npz')))
    return _CHAIN


def chain_cases(kind):
    return [c for c in chain_golden()[0] if c['kind'] == kind]


def chain_array(case, key):
    arrays = chain_golden()[1]
    name = f"{case['id']}/{key}"
    return arrays[name] if name in arrays.files else None


def plain_config(obj):
    """Product config (attrs) -> the JSON-able structure make_golden.plain produces."""
    import attrs
    if hasattr(obj, 'smooth_x') and hasattr(obj, 'smooth_y'):
        return [obj.smooth_x, obj.smooth_y]
    if attrs.has(type(obj)):
        out = {}
        for field in attrs.fields(type(obj)):
            if field.name == '_rng_state':
                continue
            out[field.name.lstrip('_')] = plain_config(getattr(obj, field.name))
        return out
    if isinstance(obj, (list, tuple)):
        return [plain_config(x) for x in obj]
    if hasattr(obj, 'value') and type(obj).__module__.startswith('vkit_b200'):
        return obj.value
    if isinstance(obj, np.integer):
        return int(obj)
    if isinstance(obj, np.floating):
        return float(obj)
    return obj


def product_config_for(op, config):
    return product_config({'op': op, 'config': config})


# ---------------------------------------------------------------------------------------------
# Round-2 fixtures (tests/golden/make_golden_r2.py): ellipse_streak, fog on GRAYSCALE, jpeg_quality
# ---------------------------------------------------------------------------------------------
_R2 = None


def r2_golden():
    global _R2
    if _R2 is None:
        with open(os.path.join(GOLDEN_DIR, 'r2_cases.json')) as fin:
            meta = json.load(fin)
        _R2 = (meta['cases'], np.load(os.path.join(GOLDEN_DIR, 'r2_arrays.npz')))
    return _R2


def r2_cases(kind):
    return [c for c in r2_golden()[0] if c['kind'] == kind]


def r2_array(case, key):
    arrays = r2_golden()[1]
    name = f"{case['id']}/{key}"
    return arrays[name] if name in arrays.files else None


def rgb_to_gray(rgb):
    """cv.cvtColor(RGB2GRAY) on uint8 (Image.to_target_mode_image, image.py:771-814), through the
    oracle's pinned model."""
    from oracle import cv2_model
    return cv2_model.cvt_rgb2gray(rgb)


# ---------------------------------------------------------------------------------------------
# Fixtures of the step before compositing (tests/golden/make_golden_f4.py): background
# synthesis, glyph preparation, LCD / default text-line rendering
# ---------------------------------------------------------------------------------------------
_F4 = None


def f4_golden():
    global _F4
    if _F4 is None:
        with open(os.path.join(GOLDEN_DIR, 'f4_cases.json')) as fin:
            meta = json.load(fin)
        _F4 = (meta['cases'], np.load(os.path.join(GOLDEN_DIR, 'f4_arrays.npz')))
    return _F4


def f4_cases(kind):
    return [c for c in f4_golden()[0] if c['kind'] == kind]


def f4_array(case, key):
    arrays = f4_golden()[1]
    name = f"{case['id']}/{key}"
    return arrays[name] if name in arrays.files else None


def f4_textures(seed, count, size_min, size_max):
    """Synthetic background textures: smooth colour ramps + noise, each with its own size and
    brightness so the grayscale-mean ordering and the sigma window of the combiner matter.
    Returns [(name, HxWx3 uint8, grayscale_mean, grayscale_std)]; the statistics are rounded to
    three decimals so JSON and both sides see the same numbers."""
    rng = np.random.default_rng(seed)
    textures = []
    for k in range(count):
        h, w = (int(v) for v in rng.integers(size_min, size_max + 1, 2))
        base = rng.integers(20, 236, 3)
        yy, xx = np.mgrid[0:h, 0:w]
        ramp = (yy * int(rng.integers(-3, 4)) + xx * int(rng.integers(-3, 4))) // 4
        mat = base[None, None, :] + ramp[:, :, None] + rng.integers(-12, 13, (h, w, 3))
        mat = np.clip(mat, 0, 255).astype(np.uint8)
        gray = mat.mean(axis=2)
        textures.append((f'tex{k:02d}.png', mat, round(float(gray.mean()), 3),
                         round(float(gray.std()), 3)))
    return textures


def f4_glyph_bitmaps(seed, count, lcd):
    """Synthetic FreeType-like coverage bitmaps with empty margins (so trimming has work to do):
    [(bitmap, bitmap_top, bitmap_left, advance_x)]."""
    rng = np.random.default_rng(seed)
    glyphs = []
    for _ in range(count):
        h, w = int(rng.integers(6, 30)), int(rng.integers(4, 24))
        shape = (h, w, 3) if lcd else (h, w)
        body = rng.integers(0, 256, shape)
        body[rng.random(shape) < 0.45] = 0
        pad = [int(v) for v in rng.integers(0, 4, 4)]
        full = np.zeros((h + pad[0] + pad[1], w + pad[2] + pad[3]) + shape[2:], dtype=np.uint8)
        full[pad[0]:pad[0] + h, pad[2]:pad[2] + w] = body
        # one certainly covered pixel per border row / column of the body keeps the trim exact
        full[pad[0], pad[2]] = 255
        full[pad[0] + h - 1, pad[2] + w - 1] = 255
        bitmap_top = int(rng.integers(-2, h + 4))
        bitmap_left = int(rng.integers(-2, 4))
        advance_x = int(rng.integers(1, (full.shape[1] + 6) * 64))
        glyphs.append((full, bitmap_top, bitmap_left, advance_x))
    return glyphs
