"""The per-element numerics of the CUDA kernels (vkit_b200/csrc/*.cuh), compiled for the CPU
by tests/hostsim, against the golden fixtures of the live reference.  CPU only: this pins the
arithmetic the kernels execute without needing a GPU (the kernels' plumbing is covered by
`-m gpu` tests)."""
import ctypes

import numpy as np
import pytest

from common import AFFINE_OPS, golden_array, golden_cases, make_inputs, product_config, sha
from vkit_b200 import _native as nv

vp = ctypes.c_void_p


@pytest.fixture(scope='module')
def hs(hostsim):
    hostsim.hs_warp.argtypes = [vp]
    hostsim.hs_grid_project.argtypes = [vp, vp]
    hostsim.hs_grid_finalize.argtypes = [vp, vp, vp, vp]
    hostsim.hs_grid_remap.argtypes = [vp, vp, vp, vp, vp]
    hostsim.hs_fast_path_stats.argtypes = [vp, vp, ctypes.c_int, ctypes.c_int, vp, vp]
    hostsim.hs_fill_poly4.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp]
    hostsim.hs_ellipse.argtypes = [ctypes.c_int] * 7 + [vp]
    hostsim.hs_homography.argtypes = [vp, vp, vp]
    assert hostsim.hs_sizeof_grid_page() == ctypes.sizeof(nv.GridPage)
    assert hostsim.hs_sizeof_warp_page() == ctypes.sizeof(nv.WarpPage)
    assert hostsim.hs_sizeof_planes() == ctypes.sizeof(nv.Planes)
    assert hostsim.hs_sizeof_grid_meta() == ctypes.sizeof(nv.GridMeta)
    return hostsim


def _planes(image, mask, score_map, dst_shape):
    dh, dw = dst_shape
    h, w = image.shape[:2]
    out = (np.zeros((dh, dw, 3), np.uint8), np.zeros((dh, dw), np.uint8),
           np.zeros((dh, dw), np.float32))
    pl = nv.Planes(image.ctypes.data, out[0].ctypes.data, mask.ctypes.data, out[1].ctypes.data,
                   score_map.ctypes.data, out[2].ctypes.data, 3, h, w, dh, dw, 0)
    return pl, out


def _grid_page(case):
    """Page record built by the PRODUCT's host code (camera / mls modules), no CUDA needed for
    the camera ops; MLS handles are passed as host pointers here."""
    from vkit_b200.mechanism.distortion.geometric import camera as cam
    from vkit_b200.mechanism.distortion.geometric._gridcore import new_grid_page
    from vkit_b200.utility import dyn_structure
    shape = tuple(case['shape'])
    cfg = product_config(case)
    keep = []
    if case['op'] == 'similarity_mls':
        rec = new_grid_page(shape[0], shape[1], cfg['grid_size'])
        rec['projector'] = nv.PROJ_MLS
        rec['resize_as_src'] = int(cfg.get('resize_as_src', False))
        src = np.ascontiguousarray(cfg['src_handle_points'].to_smooth_np_array())
        dst = np.ascontiguousarray(cfg['dst_handle_points'].to_smooth_np_array())
        keep += [src, dst]
        rec['n_handles'] = len(src)
        rec['handles_src'] = src.ctypes.data
        rec['handles_dst'] = dst.ctypes.data
    else:
        from vkit_b200.batch import _PAGE_BUILDERS
        config_cls, builder = _PAGE_BUILDERS[case['op']]
        config = dyn_structure(cfg, config_cls)
        rec = builder(config, shape)
        cam.fill_camera_model(rec, cam.complete_camera_model_config(shape[0], shape[1],
                                                                    config.camera_model_config))
    return nv.GridPage.from_buffer_copy(rec.tobytes()), keep


GRID_CASES = [c for c in golden_cases('geometric') if c['op'] not in AFFINE_OPS]
SMALL_GRID = [c for c in GRID_CASES if c['shape'][0] < 400]


@pytest.mark.parametrize('case', [c for c in golden_cases('geometric') if c['op'] in AFFINE_OPS
                                  and not c['op'].startswith('skew')],
                         ids=lambda c: f"{c['id']}-{c['op']}-{c['shape'][0]}")
def test_affine_numerics(hs, case):
    from vkit_b200.mechanism import distortion
    from vkit_b200.mechanism.distortion.geometric._hostmath import invert_affine
    shape = tuple(case['shape'])
    image, mask, score_map = make_inputs(case['seed'], shape)
    op = getattr(distortion, case['op'])
    config = op.config_cls(**case['config'])
    if config.is_nop:
        pytest.skip('nop')
    state = op.state_cls(config, shape, None)
    page = nv.WarpPage()
    page.planes, out = _planes(image, mask, score_map, state.result_shape)
    page.kind = nv.WARP_AFFINE
    inv = invert_affine(state.trans_mat).reshape(-1)
    for i in range(6):
        page.inv[i] = inv[i]
    hs.hs_warp(ctypes.byref(page))
    assert tuple(state.result_shape) == tuple(case['result_shape'])
    for key, arr in zip(('image', 'mask', 'score_map'), out):
        assert sha(arr) == case['sha'][key], key


@pytest.mark.parametrize('case', GRID_CASES, ids=lambda c: f"{c['id']}-{c['op']}-{c['shape'][0]}")
def test_lattice_numerics(hs, case):
    page, keep = _grid_page(case)
    n_points = page.rows * page.cols
    lattice_f = np.zeros((n_points, 2), np.float64)
    hs.hs_grid_project(ctypes.byref(page), lattice_f.ctypes.data)
    lattice_i = np.zeros((n_points, 2), np.int32)
    meta = nv.GridMeta()
    hs.hs_grid_finalize(ctypes.byref(page), lattice_f.ctypes.data, lattice_i.ctypes.data,
                        ctypes.byref(meta))
    ref = golden_array(case, 'lattice')
    flips = int((lattice_i != ref).any(axis=1).sum())
    if case['op'] == 'similarity_mls':
        assert flips <= 4
    else:
        assert flips == 0
        assert (meta.dst_h, meta.dst_w) == tuple(case['result_shape'])


@pytest.mark.parametrize('case', SMALL_GRID + [c for c in GRID_CASES if c['op'] == 'camera_plane_only'
                                               and c['shape'][0] == 1024],
                         ids=lambda c: f"{c['id']}-{c['op']}-{c['shape'][0]}")
def test_remap_numerics_given_reference_lattice(hs, case):
    """Owner (exact cv.fillPoly coverage, last writer wins) + homography + fixed-point bilinear,
    fed with the reference's lattice: bit-exact image, mask and score map."""
    page, keep = _grid_page(case)
    shape = tuple(case['shape'])
    image, mask, score_map = make_inputs(case['seed'], shape)
    lattice = np.ascontiguousarray(golden_array(case, 'lattice'), dtype=np.int32)
    dst_shape = tuple(case['result_shape'])
    pl, out = _planes(image, mask, score_map, dst_shape)
    owner = np.zeros(dst_shape, np.int32)
    hs.hs_grid_remap(ctypes.byref(page), lattice.ctypes.data, ctypes.byref(pl), owner.ctypes.data,
                     None)
    for key, arr in zip(('image', 'mask', 'score_map'), out):
        assert sha(arr) == case['sha'][key], key
    # float32 fast path of the coordinates: never wrong where it claims to be valid
    stats = (ctypes.c_longlong * 4)()
    hs.hs_fast_path_stats(ctypes.byref(page), lattice.ctypes.data, dst_shape[1], dst_shape[0],
                          owner.ctypes.data, stats)
    pixels, ok, wrong, max_err = list(stats)
    assert wrong == 0
    assert ok / pixels > 0.985
    # kFastSlack with a 3x margin; the error includes the rounding of the estimate to whole
    # fast-path units (2^-12 of 1/32 px: up to 1.2e-4)
    assert max_err / 1e9 < 1.0e-3 / 3
    print(f'fast path: {ok / pixels:.5f} accepted, max error {max_err / 1e9:.2e}')


def test_fill_poly_rows_random_quads(hs):
    from oracle import cv2_model as cm
    rng = np.random.default_rng(0)
    for it in range(400):
        side = int(rng.integers(2, 50))
        jitter = int(rng.integers(1, 8))
        quad = np.array([[0, 0], [side, 0], [side, side], [0, side]]) + rng.integers(
            -jitter, jitter + 1, (4, 2))
        if it % 7 == 0:
            quad = rng.integers(0, 70, (4, 2))  # arbitrary, incl. self-intersecting
        quad -= quad.min(axis=0)
        w, h = int(quad[:, 0].max() + 1), int(quad[:, 1].max() + 1)
        ref = cm.fill_poly((h, w), quad)
        out = np.zeros((h, w), np.uint8)
        pts = np.ascontiguousarray(quad.astype(np.int32))
        hs.hs_fill_poly4(pts.ctypes.data, h, w, out.ctypes.data)
        assert np.array_equal(out, ref), quad.tolist()
    # quads that fit one 32-bit window per row: also through the walked-outline + row-fill form of
    # the masks kernel (hs_fill_poly4 marks a row on which that form disagrees with 11)
    for it in range(3000):
        if it % 3 == 0:
            quad = rng.integers(0, 32, (4, 2))  # arbitrary, incl. self-intersecting / degenerate
        else:
            side = int(rng.integers(1, 24))
            jitter = int(rng.integers(0, 5))
            quad = np.array([[0, 0], [side, 0], [side, side], [0, side]]) + rng.integers(
                -jitter, jitter + 1, (4, 2))
        quad -= quad.min(axis=0)
        w, h = int(quad[:, 0].max() + 1), int(quad[:, 1].max() + 1)
        if w > 32:
            continue
        ref = cm.fill_poly((h, w), quad)
        out = np.zeros((h, w), np.uint8)
        pts = np.ascontiguousarray(quad.astype(np.int32))
        hs.hs_fill_poly4(pts.ctypes.data, h, w, out.ctypes.data)
        assert np.array_equal(out, ref), quad.tolist()


def test_homography_closed_form(hs):
    from oracle import cv2_model as cm
    rng = np.random.default_rng(1)
    for _ in range(200):
        src = np.array([[0, 0], [15, 0], [15, 15], [0, 15]], np.float64) + rng.integers(0, 900, 2)
        dst = src + rng.integers(-4, 5, (4, 2))
        H = np.zeros(9)
        hs.hs_homography(np.ascontiguousarray(dst.reshape(-1)).ctypes.data,
                         np.ascontiguousarray(src.reshape(-1)).ctypes.data, H.ctypes.data)
        ref = cm.get_perspective_transform(dst, src)
        pts = np.c_[dst, np.ones(4)]
        a = (H.reshape(3, 3) @ pts.T)
        b = (ref @ pts.T)
        # both map the dst corners onto the src corners; the closed form does so to ~1e-10,
        # the SVD solve (what cv2 runs) to ~1e-6
        np.testing.assert_allclose(a[:2] / a[2], src.T, rtol=0, atol=1e-9)
        np.testing.assert_allclose(b[:2] / b[2], src.T, rtol=0, atol=1e-4)


def test_ellipse_drawing_vs_oracle(hs):
    """The product's drawing code (vkb_draw_host.h vertices + vkb_draw.cuh primitives, host build)
    against the oracle's restatement of cv.ellipse -- and against cv2 itself where importable."""
    from oracle import cv2_draw as cd
    try:
        import cv2
    except ImportError:
        cv2 = None
    rng = np.random.default_rng(5)
    for it in range(400 if cv2 is None else 1500):
        h, w = int(rng.integers(8, 200)), int(rng.integers(8, 200))
        axes = (int(rng.integers(0, 260)), int(rng.integers(0, 260)))
        t = int(rng.integers(1, 4))
        got = np.zeros((h, w), np.uint8)
        hs.hs_ellipse(h, w, w // 2, h // 2, axes[0], axes[1], t, got.ctypes.data)
        if cv2 is not None:
            ref = np.zeros((h, w), np.uint8)
            cv2.ellipse(ref, (w // 2, h // 2), axes, 0, 0, 360, 1, t)
        else:
            ref = cd.ellipse(np.zeros((h, w), np.uint8), (w // 2, h // 2), axes, t)
        assert np.array_equal(got, ref), (h, w, axes, t)
        if cv2 is not None and it % 10 == 0:
            assert np.array_equal(cd.ellipse(np.zeros((h, w), np.uint8), (w // 2, h // 2), axes, t), ref)
