"""NumPy restatement of the OpenCV (opencv-python-headless 4.13.0.92) numerics that the
reference's distortion / blend path delegates to.  TEST INFRASTRUCTURE ONLY: imported by
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline leg; never by `vkit_b200/`.

The reference (vkit-x/vkit @ 98ada2d) holds no arithmetic of its own for these steps; it calls
cv2, an un-vendored dependency pinned only by minimum version (`setup.cfg:21-25`).  Each
function below names the reference call site it stands in for and is pinned bit-exactly (or to
the stated tolerance) against the real cv2 wheel in `tests/test_oracle_cv2_model.py` (which
runs wherever cv2 is importable) and against fixtures in `tests/golden/`.

Reference call sites:
  remap_linear              vkit/mechanism/distortion/geometric/grid_rendering/grid_blender.py:60,70,80
  warp_affine               vkit/mechanism/distortion/geometric/affine.py:40
  warp_perspective          vkit/mechanism/distortion/geometric/affine.py:43
  get_perspective_transform vkit/mechanism/distortion/geometric/grid_rendering/type.py:172,189; affine.py:326,386
  fill_poly                 vkit/element/polygon.py:75
  gaussian_blur_u8          vkit/mechanism/distortion/photometric/blur.py:65
  cvt_*                     vkit/element/image.py:794,800,808
  rodrigues/project_points  vkit/mechanism/distortion/geometric/camera.py:96,189
"""
import numpy as np

INTER_BITS = 5
INTER_TAB_SIZE = 1 << INTER_BITS  # 32 sub-pixel positions per axis
REMAP_COEF_BITS = 15


# --------------------------------------------------------------------------------------------
# Bilinear sampling shared by remap / warpAffine / warpPerspective (BORDER_CONSTANT, value 0).
# --------------------------------------------------------------------------------------------
def _tap(src, yy, xx):
    """src[yy, xx] with zeros outside the image. yy/xx int arrays of equal shape."""
    h, w = src.shape[:2]
    inside = (yy >= 0) & (yy < h) & (xx >= 0) & (xx < w)
    yc = np.clip(yy, 0, h - 1)
    xc = np.clip(xx, 0, w - 1)
    val = src[yc, xc]
    if src.ndim == 3:
        inside = inside[..., None]
    return np.where(inside, val, np.zeros((), dtype=src.dtype))


def bilinear_fixed(src, X, Y):
    """Sample `src` at 1/32-pixel fixed-point coordinates (X, Y) the way cv::remap does.

    uint8: integer weights that sum to 2^15, result (acc + 2^14) >> 15.
    float32: float32 weights ((32-f)/32 products), float32 multiply/adds in tap order.
    """
    X = np.asarray(X, dtype=np.int64)
    Y = np.asarray(Y, dtype=np.int64)
    # cv::remap stores integer coordinates as int16 (saturate_cast<short>).
    x0 = np.clip(X >> INTER_BITS, -32768, 32767)
    y0 = np.clip(Y >> INTER_BITS, -32768, 32767)
    fx = X & (INTER_TAB_SIZE - 1)
    fy = Y & (INTER_TAB_SIZE - 1)

    p00 = _tap(src, y0, x0)
    p01 = _tap(src, y0, x0 + 1)
    p10 = _tap(src, y0 + 1, x0)
    p11 = _tap(src, y0 + 1, x0 + 1)

    if src.dtype == np.uint8:
        w00 = (32 - fy) * (32 - fx) * 32
        w01 = (32 - fy) * fx * 32
        w10 = fy * (32 - fx) * 32
        w11 = fy * fx * 32
        if src.ndim == 3:
            w00, w01, w10, w11 = (w[..., None] for w in (w00, w01, w10, w11))
        acc = (p00.astype(np.int64) * w00 + p01.astype(np.int64) * w01
               + p10.astype(np.int64) * w10 + p11.astype(np.int64) * w11)
        return ((acc + (1 << (REMAP_COEF_BITS - 1))) >> REMAP_COEF_BITS).astype(np.uint8)

    if src.dtype == np.float32:
        one = np.float32(1.0)
        scale = np.float32(1.0 / INTER_TAB_SIZE)
        ax = fx.astype(np.float32) * scale
        ay = fy.astype(np.float32) * scale
        w00 = (one - ay) * (one - ax)
        w01 = (one - ay) * ax
        w10 = ay * (one - ax)
        w11 = ay * ax
        if src.ndim == 3:
            w00, w01, w10, w11 = (w[..., None] for w in (w00, w01, w10, w11))
        out = p00 * w00
        out = out + p01 * w01
        out = out + p10 * w10
        out = out + p11 * w11
        return out.astype(np.float32)

    raise NotImplementedError(src.dtype)


def cv_round(x):
    """cvRound: round-half-to-even on a double, to int64."""
    return np.rint(np.asarray(x, dtype=np.float64)).astype(np.int64)


def remap_linear(src, map_x, map_y):
    """cv.remap(src, map_x, map_y, cv.INTER_LINEAR) with BORDER_CONSTANT(0)."""
    assert map_x.dtype == np.float32 and map_y.dtype == np.float32
    # float32 * 32 is exact; saturate to int32 like saturate_cast<int>.
    X = np.clip(cv_round(map_x.astype(np.float64) * INTER_TAB_SIZE), -2**31, 2**31 - 1)
    Y = np.clip(cv_round(map_y.astype(np.float64) * INTER_TAB_SIZE), -2**31, 2**31 - 1)
    return bilinear_fixed(src, X, Y)


def invert_affine(M):
    """Inverse of a 2x3 affine map exactly as cv::warpAffine computes it (double)."""
    M = np.asarray(M, dtype=np.float64).copy()
    D = M[0, 0] * M[1, 1] - M[0, 1] * M[1, 0]
    D = 1.0 / D if D != 0 else 0.0
    A11 = M[1, 1] * D
    A22 = M[0, 0] * D
    M[0, 0] = A11
    M[0, 1] *= -D
    M[1, 0] *= -D
    M[1, 1] = A22
    b1 = -M[0, 0] * M[0, 2] - M[0, 1] * M[1, 2]
    b2 = -M[1, 0] * M[0, 2] - M[1, 1] * M[1, 2]
    M[0, 2] = b1
    M[1, 2] = b2
    return M


def affine_fixed_coords(M, dsize):
    """1/32-pixel source coordinates for every dst pixel of cv.warpAffine(src, M, dsize)."""
    width, height = dsize
    Mi = invert_affine(M)
    AB_BITS = 10
    AB_SCALE = 1 << AB_BITS
    round_delta = AB_SCALE // INTER_TAB_SIZE // 2  # 16
    xs = np.arange(width, dtype=np.float64)
    ys = np.arange(height, dtype=np.float64)
    adelta = cv_round(Mi[0, 0] * xs * AB_SCALE)
    bdelta = cv_round(Mi[1, 0] * xs * AB_SCALE)
    X0 = cv_round((Mi[0, 1] * ys + Mi[0, 2]) * AB_SCALE) + round_delta
    Y0 = cv_round((Mi[1, 1] * ys + Mi[1, 2]) * AB_SCALE) + round_delta
    X = (X0[:, None] + adelta[None, :]) >> (AB_BITS - INTER_BITS)
    Y = (Y0[:, None] + bdelta[None, :]) >> (AB_BITS - INTER_BITS)
    return X, Y


def warp_affine(src, M, dsize):
    """cv.warpAffine(src, M, dsize) (INTER_LINEAR, BORDER_CONSTANT 0)."""
    X, Y = affine_fixed_coords(M, dsize)
    return bilinear_fixed(src, X, Y)


def perspective_fixed_coords(M, dsize):
    width, height = dsize
    Mi = np.linalg.inv(np.asarray(M, dtype=np.float64))
    xs = np.arange(width, dtype=np.float64)[None, :]
    ys = np.arange(height, dtype=np.float64)[:, None]
    with np.errstate(divide='ignore', invalid='ignore'):
        W = Mi[2, 0] * xs + Mi[2, 1] * ys + Mi[2, 2]
        W = np.where(W != 0, INTER_TAB_SIZE / W, 0.0)
        fX = np.clip((Mi[0, 0] * xs + Mi[0, 1] * ys + Mi[0, 2]) * W, -2.0**31, 2.0**31 - 1)
        fY = np.clip((Mi[1, 0] * xs + Mi[1, 1] * ys + Mi[1, 2]) * W, -2.0**31, 2.0**31 - 1)
    return cv_round(fX), cv_round(fY)


def warp_perspective(src, M, dsize):
    """cv.warpPerspective(src, M, dsize) (INTER_LINEAR, BORDER_CONSTANT 0)."""
    X, Y = perspective_fixed_coords(M, dsize)
    return bilinear_fixed(src, X, Y)


# --------------------------------------------------------------------------------------------
# Homography from 4 point pairs.
# --------------------------------------------------------------------------------------------
def get_perspective_transform(src, dst):
    """cv.getPerspectiveTransform(src, dst, cv.DECOMP_SVD): 8x8 system solved in double."""
    src = np.asarray(src, dtype=np.float64)
    dst = np.asarray(dst, dtype=np.float64)
    A = np.zeros((8, 8), dtype=np.float64)
    b = np.zeros(8, dtype=np.float64)
    for i in range(4):
        sx, sy = src[i]
        dx, dy = dst[i]
        A[i] = [sx, sy, 1, 0, 0, 0, -sx * dx, -sy * dx]
        A[i + 4] = [0, 0, 0, sx, sy, 1, -sx * dy, -sy * dy]
        b[i] = dx
        b[i + 4] = dy
    U, w, Vt = np.linalg.svd(A)
    x = Vt.T @ ((U.T @ b) / w)
    return np.append(x, 1.0).reshape(3, 3)


# --------------------------------------------------------------------------------------------
# cv.fillPoly(img, [pts int32], 1) for one polygon, lineType 8, shift 0.
# --------------------------------------------------------------------------------------------
def _line8(mask, x0, y0, x1, y1):
    """cv::LineIterator(connectivity=8, leftToRight=true); both endpoints inside the canvas."""
    h, w = mask.shape
    if x0 > x1:
        x0, y0, x1, y1 = x1, y1, x0, y0
    dx = x1 - x0
    dy = y1 - y0
    sy = 1 if dy >= 0 else -1
    dy = abs(dy)
    steep = dy > dx
    if steep:
        dx, dy = dy, dx
    err = dx - 2 * dy
    x, y = x0, y0
    for _ in range(dx + 1):
        if 0 <= x < w and 0 <= y < h:
            mask[y, x] = 1
        m = err < 0
        err += -2 * dy + (2 * dx if m else 0)
        if steep:
            y += sy
            if m:
                x += 1
        else:
            x += 1
            if m:
                y += sy


def fill_poly(shape, pts):
    """Coverage (uint8 0/1) of cv.fillPoly on a zero canvas of `shape` for one polygon.

    outline (8-connected Bresenham per edge) UNION scan-line fill with 16.16 fixed-point edges,
    spans [ceil(xl), floor(xr)], edges active on [ymin, ymax).  Polygon assumed inside canvas
    (the reference always rasterises on the polygon's own bounding box, polygon.py:70-77).
    """
    h, w = shape
    mask = np.zeros((h, w), dtype=np.uint8)
    pts = [(int(p[0]), int(p[1])) for p in pts]
    n = len(pts)
    edges = []
    for i in range(n):
        x0, y0 = pts[i - 1]
        x1, y1 = pts[i]
        _line8(mask, x0, y0, x1, y1)
        if y0 == y1:
            continue
        num = (x1 - x0) << 16
        den = y1 - y0
        dxf = abs(num) // abs(den)
        if (num < 0) != (den < 0):
            dxf = -dxf  # C truncating division
        if y0 < y1:
            edges.append((y0, y1, x0 << 16, dxf))
        else:
            edges.append((y1, y0, x1 << 16, dxf))
    if not edges:
        return mask
    y_lo = min(e[0] for e in edges)
    y_hi = max(e[1] for e in edges)
    for y in range(max(y_lo, 0), min(y_hi, h)):
        xs = sorted(x + d * (y - ya) for (ya, yb, x, d) in edges if ya <= y < yb)
        for k in range(0, len(xs) - 1, 2):
            xl = (xs[k] + 65535) >> 16
            xr = xs[k + 1] >> 16
            xl = max(xl, 0)
            xr = min(xr, w - 1)
            if xl <= xr:
                mask[y, xl:xr + 1] = 1
    return mask


# --------------------------------------------------------------------------------------------
# cv.GaussianBlur(uint8, (k, k), sigma), BORDER_REFLECT_101: 8.8 fixed-point separable filter.
# --------------------------------------------------------------------------------------------
def gaussian_kernel_u8(ksize, sigma):
    """Integer (sum 256) kernel cv::getGaussianKernelBitExact builds for uint8 images."""
    r = ksize // 2
    xs = np.arange(ksize, dtype=np.float64) - r
    k = np.exp(-(xs * xs) / (2.0 * sigma * sigma))
    k = k / k.sum()
    out = np.zeros(ksize, dtype=np.int64)
    err = 0.0
    for i in range(r):
        adj = k[i] * 256.0 + err
        v = np.rint(adj)
        err = adj - v
        out[i] = out[ksize - 1 - i] = int(v)
    out[r] = 256 - out.sum()
    return out


def _reflect101(idx, n):
    if n == 1:
        return np.zeros_like(idx)
    period = 2 * (n - 1)
    idx = np.abs(idx) % period
    return np.where(idx >= n, period - idx, idx)


def gaussian_blur_u8(src, ksize, sigma):
    assert src.dtype == np.uint8
    K = gaussian_kernel_u8(ksize, sigma)
    r = ksize // 2
    h, w = src.shape[:2]
    s = src.astype(np.int64)
    cols = _reflect101(np.arange(-r, w + r), w)
    rows = _reflect101(np.arange(-r, h + r), h)
    sp = s[:, cols]
    acc = np.zeros_like(s)
    for i in range(ksize):
        acc += sp[:, i:i + w] * K[i]
    acc = np.minimum(acc, 65535)
    ap = acc[rows]
    out = np.zeros_like(s)
    for i in range(ksize):
        out += ap[i:i + h] * K[i]
    out = np.minimum((out + (1 << 15)) >> 16, 255)
    return out.astype(np.uint8)


# --------------------------------------------------------------------------------------------
# Colour conversions (uint8).
# --------------------------------------------------------------------------------------------
def _div_table(scale):
    t = np.zeros(256, dtype=np.int64)
    i = np.arange(1, 256, dtype=np.float64)
    t[1:] = np.rint(scale / i).astype(np.int64)
    return t


_SDIV = _div_table(255 << 12)
_HDIV256 = _div_table((256 << 12) / 6.0)


def cvt_rgb2hsv_full(rgb):
    """cv.cvtColor(uint8, COLOR_RGB2HSV_FULL): integer tables, hue range 256."""
    r = rgb[..., 0].astype(np.int64)
    g = rgb[..., 1].astype(np.int64)
    b = rgb[..., 2].astype(np.int64)
    v = np.maximum(np.maximum(r, g), b)
    vmin = np.minimum(np.minimum(r, g), b)
    diff = v - vmin
    vr = v == r
    vg = v == g
    h = np.where(vr, g - b, np.where(vg, b - r + 2 * diff, r - g + 4 * diff))
    s = (diff * _SDIV[v] + (1 << 11)) >> 12
    h = (h * _HDIV256[diff] + (1 << 11)) >> 12
    h = h + np.where(h < 0, 256, 0)
    out = np.stack([np.clip(h, 0, 255), s, v], axis=-1)
    return out.astype(np.uint8)


_SECTOR = np.array([[1, 3, 0], [1, 0, 2], [3, 0, 1], [0, 2, 1], [0, 1, 3], [2, 1, 0]])


def _sector_pick(tab, sector):
    """tab: (..., 4) float32, sector: (...) int in [0, 6). Returns (b, g, r) picks."""
    idx = _SECTOR[sector]  # (..., 3)
    return np.take_along_axis(tab, idx, axis=-1)


def cvt_hsv2rgb_full(hsv):
    """cv.cvtColor(uint8, COLOR_HSV2RGB_FULL): float32 path, hue scale 6/255 (+-1 vs cv2 on
    <= 78 of 2^24 inputs; cv2 itself is backend dependent here)."""
    f32 = np.float32
    h = hsv[..., 0].astype(f32) * f32(6.0 / 255.0)
    s = hsv[..., 1].astype(f32) * f32(1.0 / 255.0)
    v = hsv[..., 2].astype(f32) * f32(1.0 / 255.0)
    sector = np.floor(h).astype(np.int64)
    frac = h - sector.astype(f32)
    sector = sector % 6
    one = f32(1.0)
    tab = np.stack([v, v * (one - s), v * (one - s * frac), v * (one - s * (one - frac))],
                   axis=-1).astype(f32)
    bgr = _sector_pick(tab, sector)
    gray = np.stack([v, v, v], axis=-1)
    bgr = np.where((hsv[..., 1] == 0)[..., None], gray, bgr)
    rgb = bgr[..., ::-1]
    return np.clip(np.rint(rgb * f32(255.0)), 0, 255).astype(np.uint8)


# RCPPS(d), d = 1..255: x86's 12-bit hardware reciprocal approximation (Intel's table; the values
# _mm_rcp_ss returns -- generator: oracle/ipp_rcp_table.c).  IPP's RGB -> HLS divides with it.
IPP_RCP = np.array([
    0.0,
    0.999755859, 0.49987793, 0.333251953, 0.249938965, 0.199951172, 0.166625977, 0.142822266,
    0.124969482, 0.111083984, 0.0999755859, 0.0908966064, 0.0833129883, 0.0769042969, 0.0714111328,
    0.0666503906, 0.0624847412, 0.058807373, 0.0555419922, 0.0526199341, 0.049987793, 0.0476074219,
    0.0454483032, 0.04347229, 0.0416564941, 0.0399932861, 0.0384521484, 0.0370330811, 0.0357055664,
    0.0344772339, 0.0333251953, 0.0322570801, 0.0312423706, 0.0302963257, 0.0294036865, 0.0285644531,
    0.0277709961, 0.0270195007, 0.026309967, 0.0256347656, 0.0249938965, 0.0243873596, 0.0238037109,
    0.0232505798, 0.0227241516, 0.0222167969, 0.021736145, 0.0212745667, 0.0208282471, 0.0204048157,
    0.0199966431, 0.0196037292, 0.0192260742, 0.018863678, 0.0185165405, 0.0181808472, 0.0178527832,
    0.017539978, 0.0172386169, 0.0169487, 0.0166625977, 0.0163917542, 0.01612854, 0.0158691406,
    0.0156211853, 0.0153808594, 0.0151481628, 0.0149211884, 0.0147018433, 0.0144901276, 0.0142822266,
    0.014081955, 0.013885498, 0.0136947632, 0.0135097504, 0.0133304596, 0.0131549835, 0.0129852295,
    0.0128173828, 0.0126552582, 0.0124969482, 0.012342453, 0.0121936798, 0.012046814, 0.0119018555,
    0.011762619, 0.0116252899, 0.0114917755, 0.0113620758, 0.0112342834, 0.0111083984, 0.0109863281,
    0.0108680725, 0.0107517242, 0.0106372833, 0.0105247498, 0.0104141235, 0.010307312, 0.0102024078,
    0.010099411, 0.00999832153, 0.0098991394, 0.00980186462, 0.00970649719, 0.00961303711, 0.00952148438,
    0.00943183899, 0.00934410095, 0.00925827026, 0.00917243958, 0.00909042358, 0.00900840759, 0.0089263916,
    0.00884819031, 0.00876998901, 0.00869369507, 0.00861930847, 0.00854492188, 0.00847434998, 0.00840187073,
    0.00833129883, 0.00826263428, 0.00819587708, 0.00812911987, 0.00806427002, 0.00799942017, 0.00793457031,
    0.00787353516, 0.00781059265, 0.00775051117, 0.00769042969, 0.00763130188, 0.00757408142, 0.00751686096,
    0.00746059418, 0.00740528107, 0.00735092163, 0.00729751587, 0.00724506378, 0.00719261169, 0.00714111328,
    0.00709056854, 0.00704097748, 0.00699138641, 0.00694274902, 0.00689506531, 0.00684738159, 0.00680160522,
    0.00675487518, 0.00671005249, 0.0066652298, 0.00662136078, 0.00657749176, 0.00653457642, 0.00649261475,
    0.00645065308, 0.00640869141, 0.00636768341, 0.00632762909, 0.00628852844, 0.00624847412, 0.00621032715,
    0.0061712265, 0.0061340332, 0.0060968399, 0.00605964661, 0.00602340698, 0.00598716736, 0.00595092773,
    0.00591564178, 0.00588130951, 0.00584697723, 0.00581264496, 0.00577926636, 0.00574588776, 0.00571346283,
    0.0056810379, 0.00564861298, 0.00561714172, 0.00558567047, 0.00555419922, 0.00552368164, 0.00549316406,
    0.00546360016, 0.00543403625, 0.00540447235, 0.00537586212, 0.00534629822, 0.00531864166, 0.00529003143,
    0.00526237488, 0.00523471832, 0.00520706177, 0.00518035889, 0.00515365601, 0.00512695312, 0.00510120392,
    0.00507545471, 0.00504970551, 0.0050239563, 0.00499916077, 0.00497436523, 0.0049495697, 0.00492572784,
    0.00490093231, 0.00487709045, 0.0048532486, 0.00483036041, 0.00480651855, 0.00478363037, 0.00476074219,
    0.00473880768, 0.00471591949, 0.00469398499, 0.00467205048, 0.00465011597, 0.00462913513, 0.00460720062,
    0.00458621979, 0.00456523895, 0.00454521179, 0.00452423096, 0.0045042038, 0.00448322296, 0.0044631958,
    0.00444412231, 0.00442409515, 0.00440502167, 0.00438499451, 0.00436592102, 0.00434684753, 0.00432872772,
    0.00430965424, 0.00429153442, 0.00427246094, 0.00425434113, 0.00423717499, 0.00421905518, 0.00420093536,
    0.00418376923, 0.00416564941, 0.00414848328, 0.00413131714, 0.00411510468, 0.00409793854, 0.0040807724,
    0.00406455994, 0.00404834747, 0.00403213501, 0.00401592255, 0.00399971008, 0.00398349762, 0.00396728516,
    0.00395202637, 0.00393676758, 0.00392150879,
], dtype=np.float32)


def cvt_rgb2hls_full(rgb):
    """cv.cvtColor(uint8, COLOR_RGB2HLS_FULL) as the default (IPP) backend of the 4.13 wheel
    computes it, bit exact on all 2^24 colours (test_cvt_hls_full_colour_cube):
      L = round_half_even((max + min) / 2);
      S = rint(diff * RCPPS(den) * 255) with den = max + min when <= 255, else 510 - (max + min);
      H = rint(h * 42.5), h = (g - b) * RCPPS(diff) (+ 2 / + 4 when green / blue is the maximum,
          tested in the order R, G, B), + 6 when negative, 256 wraps to 0; float32 products.
    (OpenCV's own path -- cv.ipp.setUseIPP(False) -- differs from this on 60 % of the colours.)"""
    f32 = np.float32
    ri = rgb[..., 0].astype(np.int64)
    gi = rgb[..., 1].astype(np.int64)
    bi = rgb[..., 2].astype(np.int64)
    imax = np.maximum(np.maximum(ri, gi), bi)
    imin = np.minimum(np.minimum(ri, gi), bi)
    isum = imax + imin
    diff = imax - imin
    half = isum >> 1
    l_int = np.where((isum & 1) == 0, half, half + (half & 1))
    den = np.where(isum <= 255, isum, 510 - isum)
    s_int = np.rint((diff.astype(f32) * IPP_RCP[den]) * f32(255.0)).astype(np.int64)
    rd = IPP_RCP[diff]
    h = np.where(imax == ri, (gi - bi).astype(f32) * rd,
                 np.where(imax == gi, (bi - ri).astype(f32) * rd + f32(2.0),
                          (ri - gi).astype(f32) * rd + f32(4.0))).astype(f32)
    h = np.where(h < 0, h + f32(6.0), h).astype(f32)
    h_int = np.rint(h * f32(42.5)).astype(np.int64)
    h_int = np.where(h_int >= 256, h_int - 256, h_int)
    zero = diff == 0
    h_int = np.where(zero, 0, h_int)
    s_int = np.where(zero, 0, s_int)
    return np.stack([h_int, l_int, s_int], axis=-1).astype(np.uint8)


def cvt_hls2rgb_full(hls):
    """cv.cvtColor(uint8, COLOR_HLS2RGB_FULL): float32 path, hue scale 6/255.  Equal to OpenCV's own
    path on all 2^24 HLS triples and to the wheel's default (IPP) on all but 20 of them (+-1 in
    one channel, float32 ties)."""
    f32 = np.float32
    h = hls[..., 0].astype(f32) * f32(6.0 / 255.0)
    l = hls[..., 1].astype(f32) * f32(1.0 / 255.0)
    s = hls[..., 2].astype(f32) * f32(1.0 / 255.0)
    one = f32(1.0)
    p2 = np.where(l <= f32(0.5), l * (one + s), l + s - l * s)
    p1 = f32(2.0) * l - p2
    sector = np.floor(h).astype(np.int64)
    frac = h - sector.astype(f32)
    sector = sector % 6
    tab = np.stack([p2, p1, p1 + (p2 - p1) * (one - frac), p1 + (p2 - p1) * frac],
                   axis=-1).astype(f32)
    bgr = _sector_pick(tab, sector)
    gray = np.stack([l, l, l], axis=-1)
    bgr = np.where((hls[..., 2] == 0)[..., None], gray, bgr)
    rgb = bgr[..., ::-1]
    return np.clip(np.rint(rgb * f32(255.0)), 0, 255).astype(np.uint8)


def cvt_rgb2gray(rgb):
    """cv.cvtColor(uint8, COLOR_RGB2GRAY): (R*9798 + G*19235 + B*3735 + 2^14) >> 15
    (pinned over all 2^24 colours against cv2 4.13.0)."""
    r = rgb[..., 0].astype(np.int64)
    g = rgb[..., 1].astype(np.int64)
    b = rgb[..., 2].astype(np.int64)
    return ((r * 9798 + g * 19235 + b * 3735 + (1 << 14)) >> 15).astype(np.uint8)


def equalize_hist(gray):
    """cv.equalizeHist(uint8 HxW) (vkit/mechanism/distortion/photometric/color.py:284)."""
    hist = np.bincount(gray.reshape(-1), minlength=256)
    total = int(hist.sum())
    first = int(np.nonzero(hist)[0][0])
    if int(hist[first]) == total:
        return np.full_like(gray, first)
    scale = np.float32(255.0) / np.float32(total - int(hist[first]))
    lut = np.zeros(256, dtype=np.uint8)
    cumulative = np.cumsum(hist[first + 1:].astype(np.int64))
    lut[first + 1:] = np.clip(np.rint(cumulative.astype(np.float32) * scale), 0, 255).astype(np.uint8)
    return lut[gray]


def cvt_gray2rgb(gray):
    return np.repeat(gray[..., None], 3, axis=-1)


# --------------------------------------------------------------------------------------------
# Pin-hole camera helpers (double precision, as inside cv2).
# --------------------------------------------------------------------------------------------
def rodrigues(rvec):
    """cv.Rodrigues(rvec) -> 3x3 rotation matrix, computed in double."""
    r = np.asarray(rvec, dtype=np.float64).reshape(3)
    theta = np.sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2])
    if theta < np.finfo(np.float64).eps:
        return np.eye(3)
    c = np.cos(theta)
    s = np.sin(theta)
    c1 = 1.0 - c
    x, y, z = r / theta
    rrt = np.array([[x * x, x * y, x * z], [x * y, y * y, y * z], [x * z, y * z, z * z]])
    rx = np.array([[0, -z, y], [z, 0, -x], [-y, x, 0]])
    return c * np.eye(3) + c1 * rrt + s * rx


def project_points(pts3d, rvec, tvec, K):
    """cv.projectPoints(pts3d, rvec, tvec, K, zeros(5)) -> (N, 2), output dtype = input dtype."""
    out_dtype = pts3d.dtype
    P = np.asarray(pts3d, dtype=np.float64)
    R = rodrigues(np.asarray(rvec, dtype=np.float64))
    t = np.asarray(tvec, dtype=np.float64).reshape(3)
    K = np.asarray(K, dtype=np.float64)
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    X, Y, Z = P[:, 0], P[:, 1], P[:, 2]
    x = R[0, 0] * X + R[0, 1] * Y + R[0, 2] * Z + t[0]
    y = R[1, 0] * X + R[1, 1] * Y + R[1, 2] * Z + t[1]
    z = R[2, 0] * X + R[2, 1] * Y + R[2, 2] * Z + t[2]
    z = np.where(z != 0, 1.0 / np.where(z != 0, z, 1.0), 1.0)
    x = x * z
    y = y * z
    return np.stack([x * fx + cx, y * fy + cy], axis=-1).astype(out_dtype)
