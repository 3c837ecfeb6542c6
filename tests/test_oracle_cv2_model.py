"""oracle/cv2_model.py against the real cv2 wheel (skipped when cv2 is not importable).
Pins the NumPy restatement of every cv2 function on the path to the installed OpenCV."""
import math

import numpy as np
import pytest

cv = pytest.importorskip('cv2')
from oracle import cv2_model as cm  # noqa: E402


def _same(a, b):
    return np.array_equal(np.asarray(a), np.asarray(b))


@pytest.fixture(scope='module')
def rng():
    return np.random.default_rng(1)


@pytest.mark.parametrize('channels', [None, 3, 4])
def test_remap_u8(rng, channels):
    shape = (97, 131) if channels is None else (97, 131, channels)
    src = rng.integers(0, 256, shape, dtype=np.uint8)
    map_x = rng.uniform(-5, 136, (120, 150)).astype(np.float32)
    map_y = rng.uniform(-5, 102, (120, 150)).astype(np.float32)
    map_x[:10, :10] = np.arange(10, dtype=np.float32)[None, :]  # exact integer coordinates
    map_y[:10, :10] = np.arange(10, dtype=np.float32)[:, None]
    map_x[20:30, :20] = (np.arange(20, dtype=np.float32) / 2)[None, :]  # exact half pixels
    map_y[20:30, :20] = 7.5
    assert _same(cm.remap_linear(src, map_x, map_y), cv.remap(src, map_x, map_y, cv.INTER_LINEAR))


def test_remap_f32(rng):
    src = rng.random((97, 131)).astype(np.float32)
    map_x = rng.uniform(-5, 136, (120, 150)).astype(np.float32)
    map_y = rng.uniform(-5, 102, (120, 150)).astype(np.float32)
    assert _same(cm.remap_linear(src, map_x, map_y), cv.remap(src, map_x, map_y, cv.INTER_LINEAR))


@pytest.mark.parametrize('angle', [30, 77, 200, 0.5])
def test_warp_affine(rng, angle):
    rad = math.radians(angle)
    M = np.asarray([[math.cos(rad), -math.sin(rad), 40], [math.sin(rad), math.cos(rad), -10.3]],
                   dtype=np.float32)
    for src in (rng.integers(0, 256, (211, 173, 3), dtype=np.uint8),
                rng.integers(0, 2, (211, 173), dtype=np.uint8),
                rng.random((211, 173)).astype(np.float32)):
        assert _same(cm.warp_affine(src, M, (300, 260)), cv.warpAffine(src, M, (300, 260)))


def test_warp_perspective_and_homography(rng):
    src_q = np.asarray([[0, 0], [172, 0], [172, 210], [0, 210]], dtype=np.float32)
    dst_q = np.asarray([[0, 23], [172, 0], [172, 210], [0, 181]], dtype=np.float32)
    M = cv.getPerspectiveTransform(src_q, dst_q, cv.DECOMP_SVD)
    np.testing.assert_allclose(cm.get_perspective_transform(src_q, dst_q), M, rtol=0, atol=1e-9)
    src = rng.integers(0, 256, (211, 173, 3), dtype=np.uint8)
    got, ref = cm.warp_perspective(src, M, (173, 211)), cv.warpPerspective(src, M, (173, 211))
    assert (got != ref).any(axis=-1).mean() <= 0.005  # exact but for 1/64 px ties (block order)
    srcf = rng.random((211, 173)).astype(np.float32)
    got, ref = cm.warp_perspective(srcf, M, (173, 211)), cv.warpPerspective(srcf, M, (173, 211))
    assert (got != ref).mean() <= 0.005


def test_fill_poly(rng):
    for it in range(600):
        side = int(rng.integers(3, 40))
        quad = np.array([[0, 0], [side, 0], [side, side], [0, side]]) + rng.integers(-4, 5, (4, 2))
        if it % 9 == 0:
            quad = rng.integers(0, 60, (4, 2))
        quad -= quad.min(axis=0)
        w, h = int(quad[:, 0].max() + 1), int(quad[:, 1].max() + 1)
        ref = np.zeros((h, w), np.uint8)
        cv.fillPoly(ref, [quad.astype(np.int32)], 1)
        assert _same(cm.fill_poly((h, w), quad), ref), quad.tolist()


@pytest.mark.parametrize('sigma', [0.5, 0.62, 0.75, 0.9, 1.0, 2.0])
def test_gaussian_blur(rng, sigma):
    k = max(3, round(3 * sigma) + 1)
    k += (k % 2 == 0)
    for shape in ((77, 91, 3), (40, 33)):
        src = rng.integers(0, 256, shape, dtype=np.uint8)
        assert _same(cm.gaussian_blur_u8(src, k, sigma), cv.GaussianBlur(src, (k, k), sigma))


def test_colour_conversions_all_colours():
    grid = np.stack(np.meshgrid(np.arange(256), np.arange(256), np.arange(256), indexing='ij'),
                    -1).astype(np.uint8).reshape(4096, 4096, 3)
    assert _same(cm.cvt_rgb2hsv_full(grid), cv.cvtColor(grid, cv.COLOR_RGB2HSV_FULL))
    assert _same(cm.cvt_rgb2gray(grid), cv.cvtColor(grid, cv.COLOR_RGB2GRAY))
    for model, code, max_bad in ((cm.cvt_hsv2rgb_full, cv.COLOR_HSV2RGB_FULL, 200),
                                 (cm.cvt_hls2rgb_full, cv.COLOR_HLS2RGB_FULL, 32)):
        diff = np.abs(model(grid).astype(int) - cv.cvtColor(grid, code).astype(int))
        assert diff.max() <= 1 and (diff > 0).any(axis=-1).sum() <= max_bad
    # RGB -> HLS_FULL: the wheel's IPP routine, reciprocals by RCPPS -- all 2^24 colours exact
    # (needs the x86 host the fixtures were made on: Intel's RCPPS table is part of the model)
    assert _same(cm.cvt_rgb2hls_full(grid), cv.cvtColor(grid, cv.COLOR_RGB2HLS_FULL))


def test_camera_functions(rng):
    rvec = (np.asarray([0.6, 0.7, 0.3], np.float32) / np.float32(np.linalg.norm([0.6, 0.7, 0.3]))
            ) * np.float32(0.25)
    R, _ = cv.Rodrigues(rvec.astype(np.float64))
    assert _same(cm.rodrigues(rvec), R)
    pts = np.hstack([rng.uniform(0, 1024, (500, 2)), rng.uniform(-100, 100, (500, 1))])
    tvec = np.asarray([[-500.], [-510.], [1024.]], np.float32)
    K = np.asarray([[1024, 0, 0], [0, 1024, 0], [0, 0, 1]], np.float32)
    for dtype in (np.float32, np.float64):
        p = pts.astype(dtype)
        ref, _ = cv.projectPoints(p, rvec, tvec, K, np.zeros(5))
        assert _same(cm.project_points(p, rvec, tvec, K), ref.reshape(-1, 2))


@pytest.mark.parametrize('shape', [(64, 96), (100, 133), (77, 50)])
def test_resize_models_vs_cv2(shape):
    """The oracle's cv.resize restatements against the cv2 wheel: NEAREST, LINEAR, NEAREST_EXACT
    and LINEAR_EXACT bit for bit; CUBIC bit for bit against cv2 with IPP switched off (its own path:
    integer horizontal pass, float32 vertical pass) and, for the wheel's default, the float64
    bicubic against IPP (+-1 on < 3e-4 of the pixels)."""
    import cv2 as cv
    from oracle import vkit_port as port
    port.use_cv2(False)
    rng = np.random.default_rng(shape[0])
    img = rng.integers(0, 256, shape + (3,), dtype=np.uint8)
    for ratio in (0.13, 0.37, 0.5, 0.8, 1.07, 1.5, 2.3):
        dsize = (max(1, round(shape[1] * ratio)), max(1, round(shape[0] * ratio)))
        assert np.array_equal(port.resize_u8(img, dsize), cv.resize(img, dsize, interpolation=cv.INTER_LINEAR))
        assert np.array_equal(port.resize_u8(img, dsize, nearest=True),
                              cv.resize(img, dsize, interpolation=cv.INTER_NEAREST))
        assert np.array_equal(port.resize_exact_u8(img, dsize),
                              cv.resize(img, dsize, interpolation=cv.INTER_LINEAR_EXACT))
        assert np.array_equal(port.resize_exact_u8(img, dsize, nearest=True),
                              cv.resize(img, dsize, interpolation=cv.INTER_NEAREST_EXACT))
        assert np.array_equal(port.resize_lanczos4_u8(img, dsize),
                              cv.resize(img, dsize, interpolation=cv.INTER_LANCZOS4))
        use_ipp = cv.ipp.useIPP()
        cv.ipp.setUseIPP(False)
        try:
            ref = cv.resize(img, dsize, interpolation=cv.INTER_CUBIC)
        finally:
            cv.ipp.setUseIPP(use_ipp)
        assert np.array_equal(port.resize_cubic_u8(img, dsize, ipp=False), ref)
        gray = np.ascontiguousarray(img[:, :, 0])
        cv.ipp.setUseIPP(False)
        try:
            ref = cv.resize(gray, dsize, interpolation=cv.INTER_CUBIC)
        finally:
            cv.ipp.setUseIPP(use_ipp)
        assert np.array_equal(port.resize_cubic_u8(gray, dsize, ipp=False), ref)
        if use_ipp:
            # the wheel's default (Intel IPP for sources of at least 4 x 4): the float64 bicubic
            for src in (img, gray):
                diff = np.abs(port.resize_cubic_u8(src, dsize).astype(int)
                              - cv.resize(src, dsize, interpolation=cv.INTER_CUBIC).astype(int))
                assert diff.max() <= 1 and (diff > 0).mean() <= 3e-4, (dsize, (diff > 0).mean())
    tiny = img[:3, :9]  # below 4 x 4 the wheel takes cv2's own path even with IPP
    for dsize in ((14, 4), (30, 10), (7, 2)):
        assert np.array_equal(port.resize_cubic_u8(tiny, dsize),
                              cv.resize(tiny, dsize, interpolation=cv.INTER_CUBIC))


@pytest.mark.parametrize('shape', [(67, 91), (40, 33)])
def test_resize_f32_model_vs_cv2(shape):
    """ScoreMap resize (float32): the oracle's plain-float32 restatement against cv2.  Without IPP
    every code is bit identical (cv2's vector code adds the rows last to first, its scalar tail
    first to last); the
    wheel's default IPP backend agrees to 1e-5."""
    import cv2 as cv
    from oracle import vkit_port as port
    port.use_cv2(False)
    rng = np.random.default_rng(shape[1])
    mat = rng.random(shape, dtype=np.float32)
    use_ipp = cv.ipp.useIPP()
    try:
        for dsize in ((40, 30), (200, 150), (91, 120), (133, 67), (300, 300), (93, 50), (94, 77)):
            for inter in (0, 1, 2, 4, 5, 6):
                got = port.resize_f32(mat, dsize, inter, ipp=False)
                cv.ipp.setUseIPP(False)
                own = cv.resize(mat, dsize, interpolation=inter)
                cv.ipp.setUseIPP(True)
                ipp = cv.resize(mat, dsize, interpolation=inter)
                assert np.array_equal(got, own), (dsize, inter)
                assert np.abs(got - ipp).max() <= 1e-5
                if use_ipp:
                    # the default model follows the wheel's backend (IPP: double coordinates)
                    assert np.abs(port.resize_f32(mat, dsize, inter) - ipp).max() <= 5e-7, (dsize, inter)
                clipped = port.resize_f32(mat, dsize, inter, clip01=True, ipp=False)
                assert np.array_equal(clipped, np.clip(got, 0.0, 1.0))
    finally:
        cv.ipp.setUseIPP(use_ipp)


@pytest.mark.parametrize('shape', [(120, 180), (97, 131)])
def test_resize_area_model_vs_cv2(shape):
    """INTER_AREA (page_resizing's fifth interpolation, sampled when shrinking): the oracle's
    restatement of cv2's box-sum and weight-table paths, and of the "area mode" bilinear passes
    cv2 runs when an axis enlarges.  uint8 bit for bit with and without IPP; float32 bit for
    bit without IPP, within 1e-5 with it."""
    import cv2 as cv
    from oracle import vkit_port as port
    port.use_cv2(False)
    rng = np.random.default_rng(shape[0])
    img = rng.integers(0, 256, shape + (3,), dtype=np.uint8)
    mat = rng.random(shape, dtype=np.float32)
    h, w = shape
    sizes = [(w // 2, h // 2), (w // 3, h // 3), (w // 4, h // 4), (w, h // 2), (77, 53), (100, 89),
             (w - 1, h), (31, 17), (w // 3, h // 4), (w, h)]
    use_ipp = cv.ipp.useIPP()
    try:
        for dsize in sizes:
            got_rgb, got_gray = port.resize_area(img, dsize), port.resize_area(img[:, :, 0], dsize)
            got = port.resize_area(mat, dsize)
            for ipp in (False, True):
                cv.ipp.setUseIPP(ipp)
                assert np.array_equal(got_rgb, cv.resize(img, dsize, interpolation=cv.INTER_AREA)), dsize
                assert np.array_equal(got_gray, cv.resize(img[:, :, 0], dsize, interpolation=cv.INTER_AREA))
                ref = cv.resize(mat, dsize, interpolation=cv.INTER_AREA)
                if ipp:
                    assert np.abs(got - ref).max() <= 1e-5
                else:
                    assert np.array_equal(got, ref), dsize
    finally:
        cv.ipp.setUseIPP(use_ipp)
    # an enlarging axis: cv2 runs the bilinear passes with "area mode" fractions
    try:
        for dsize in ((w + 1, h), (w * 2, h * 2), (round(w * 1.37), round(h * 1.37)), (w // 2, h * 2),
                      (w * 3, h // 3)):
            for ipp in (False, True):
                cv.ipp.setUseIPP(ipp)
                assert np.array_equal(port.resize_area(img, dsize),
                                      cv.resize(img, dsize, interpolation=cv.INTER_AREA)), dsize
                assert np.array_equal(port.resize_area(mat, dsize),
                                      cv.resize(mat, dsize, interpolation=cv.INTER_AREA)), dsize
    finally:
        cv.ipp.setUseIPP(use_ipp)


def test_draw_models_vs_cv2():
    """oracle/cv2_draw.py (what ellipse_streak's restatement draws with) against the wheel: thin and
    thick lines with a 16-bit shift, filled circles, convex quads and whole ellipses, clipped by
    the canvas in every direction."""
    from oracle import cv2_draw as cd
    rng = np.random.default_rng(20261017)
    one = 65536
    for _ in range(300):
        h, w = int(rng.integers(5, 80)), int(rng.integers(5, 80))
        p1 = (int(rng.integers(-20 * one, (w + 20) * one)), int(rng.integers(-20 * one, (h + 20) * one)))
        p2 = (int(rng.integers(-20 * one, (w + 20) * one)), int(rng.integers(-20 * one, (h + 20) * one)))
        t = int(rng.integers(1, 5))
        ref = np.zeros((h, w), np.uint8)
        cv.line(ref, p1, p2, 1, t, 8, 16)
        got = np.zeros((h, w), np.uint8)
        cd.thick_line(got, p1, p2, t, 3)
        assert np.array_equal(ref, got), ('line', h, w, p1, p2, t)
        c, r = (int(rng.integers(-5, w + 5)), int(rng.integers(-5, h + 5))), int(rng.integers(0, 12))
        ref = np.zeros((h, w), np.uint8)
        cv.circle(ref, c, r, 1, -1)
        got = np.zeros((h, w), np.uint8)
        cd.circle_fill(got, c[0], c[1], r)
        assert np.array_equal(ref, got), ('circle', h, w, c, r)
    for _ in range(400):
        h, w = int(rng.integers(8, 300)), int(rng.integers(8, 300))
        axes = (int(rng.integers(0, 320)), int(rng.integers(0, 320)))
        t = int(rng.integers(1, 4))
        ref = np.zeros((h, w), np.uint8)
        cv.ellipse(ref, (w // 2, h // 2), axes, 0, 0, 360, 1, t)
        got = cd.ellipse(np.zeros((h, w), np.uint8), (w // 2, h // 2), axes, t)
        assert np.array_equal(ref, got), ('ellipse', h, w, axes, t)


def test_jpeg_model_vs_cv2():
    """oracle/jpeg_model.py against cv.imencode / cv.imdecode: every quality, sides from 1 px up
    (ragged MCUs, chroma planes of 1 - 2 samples), noise / smooth / structured content, RGB and
    GRAYSCALE."""
    from oracle import jpeg_model as jm
    rng = np.random.default_rng(20261017)

    def round_trip(mat, q):
        _, buf = cv.imencode('.jpeg', mat, [cv.IMWRITE_JPEG_QUALITY, q])
        return cv.imdecode(buf, cv.IMREAD_UNCHANGED)

    for it in range(500):
        h, w = int(rng.integers(1, 90)), int(rng.integers(1, 90))
        q = int(rng.integers(0, 101))
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        if it % 3 == 1:
            img = cv.GaussianBlur(img, (0, 0), 2.5)
        if it % 3 == 2:
            img = (np.indices((h, w)).sum(0)[..., None] * np.array([3, 5, 7]) % 256).astype(np.uint8)
        assert np.array_equal(round_trip(img, q), jm.jpeg_round_trip(img, q)), ('rgb', h, w, q)
        gray = np.ascontiguousarray(img[..., 0])
        assert np.array_equal(round_trip(gray, q), jm.jpeg_round_trip(gray, q)), ('gray', h, w, q)
    img = cv.GaussianBlur(rng.integers(0, 256, (512, 640, 3), dtype=np.uint8), (0, 0), 4.0)
    assert np.array_equal(round_trip(img, 12), jm.jpeg_round_trip(img, 12))
