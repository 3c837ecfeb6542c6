"""cv2_draw.py -- TEST INFRASTRUCTURE (oracle): restatement of the OpenCV 4.13 drawing routines that
`cv.ellipse(img, center, axes, 0, 0, 360, color, thickness)` runs through (imgproc/src/drawing.cpp:
ellipse -> EllipseEx -> ellipse2Poly -> PolyLine -> ThickLine -> Line2 / FillConvexPoly / Circle),
for 8-bit single-channel canvases and lineType = LINE_8, which is how vkit's ellipse_streak draws
its concentric ellipses (vkit/mechanism/distortion/photometric/streak.py:282-337).

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs import this package.  The arithmetic
lives in opencv-python-headless (not vendored in /root/reference); it is pinned against the
installed wheel (4.13.0.92) by tests/test_oracle_cv2_model.py::test_draw_models_vs_cv2 on random
ellipses, thick lines, convex quads and circles.

All coordinates below are 16.16 fixed point (XY_SHIFT = 16) unless a name says `px`.
"""
import math

import numpy as np

XY_SHIFT = 16
XY_ONE = 1 << XY_SHIFT

# SinTable of drawing.cpp: sin of 0 .. 450 degrees as float literals with 7 decimals
SIN_TABLE = np.array([np.float32(round(math.sin(math.radians(d)), 7)) for d in range(451)],
                     dtype=np.float32)


def cv_round(v: float) -> int:
    """cvRound: round half to even (lrint / cvtsd2si)."""
    return int(np.rint(v))


def ellipse2poly(cx: float, cy: float, ax: float, ay: float, delta: int):
    """ellipse2Poly(Point2d center, Size2d axes, angle = 0, 0, 360, delta): double vertices."""
    pts = []
    alpha, beta = float(SIN_TABLE[450]), float(SIN_TABLE[0])
    i = 0
    while i < 360 + delta:
        angle = min(i, 360)
        x = ax * float(SIN_TABLE[450 - angle])
        y = ay * float(SIN_TABLE[angle])
        pts.append((cx + x * alpha - y * beta, cy + x * beta + y * alpha))
        i += delta
    if len(pts) == 1:
        pts = [(cx, cy), (cx, cy)]
    return pts


def ellipse_vertices(center_px, axes_px):
    """EllipseEx up to the vertex list (16.16 integers, consecutive duplicates removed)."""
    cx, cy = int(center_px[0]) << XY_SHIFT, int(center_px[1]) << XY_SHIFT
    ax, ay = abs(int(axes_px[0])) << XY_SHIFT, abs(int(axes_px[1])) << XY_SHIFT
    delta = (max(ax, ay) + (XY_ONE >> 1)) >> XY_SHIFT
    delta = 90 if delta < 3 else 30 if delta < 10 else 18 if delta < 15 else 5
    out = []
    prev = None
    for vx, vy in ellipse2poly(float(cx), float(cy), float(ax), float(ay), delta):
        px = cv_round(vx / XY_ONE) << XY_SHIFT
        py = cv_round(vy / XY_ONE) << XY_SHIFT
        px += cv_round(vx - px)
        py += cv_round(vy - py)
        if (px, py) != prev:
            out.append((px, py))
            prev = (px, py)
    if len(out) <= 1:
        out = [(cx, cy), (cx, cy)]
    return out


def _trunc_div(a: float, b: float) -> int:
    return int(a / b)  # (int64)(double): truncation toward zero


def clip_line(width: int, height: int, p1, p2):
    """clipLine(Size2l, Point2l&, Point2l&): (inside?, p1, p2)."""
    x1, y1 = p1
    x2, y2 = p2
    right, bottom = width - 1, height - 1
    if width <= 0 or height <= 0:
        return False, p1, p2
    c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8
    c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8
    if (c1 & c2) == 0 and (c1 | c2) != 0:
        if c1 & 12:
            a = 0 if c1 < 8 else bottom
            x1 += _trunc_div(float(a - y1) * (x2 - x1), (y2 - y1))
            y1 = a
            c1 = (x1 < 0) + (x1 > right) * 2
        if c2 & 12:
            a = 0 if c2 < 8 else bottom
            x2 += _trunc_div(float(a - y2) * (x2 - x1), (y2 - y1))
            y2 = a
            c2 = (x2 < 0) + (x2 > right) * 2
        if (c1 & c2) == 0 and (c1 | c2) != 0:
            if c1:
                a = 0 if c1 == 1 else right
                y1 += _trunc_div(float(a - x1) * (y2 - y1), (x2 - x1))
                x1 = a
                c1 = 0
            if c2:
                a = 0 if c2 == 1 else right
                y2 += _trunc_div(float(a - x2) * (y2 - y1), (x2 - x1))
                x2 = a
                c2 = 0
    return (c1 | c2) == 0, (x1, y1), (x2, y2)


def _c_div(a: int, b: int) -> int:
    """C integer division (truncation toward zero)."""
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q


def line2(img: np.ndarray, p1, p2, value=1):
    """Line2: fixed-point DDA between two 16.16 points (both end points drawn)."""
    height, width = img.shape
    ok, p1, p2 = clip_line(width << XY_SHIFT, height << XY_SHIFT, p1, p2)
    if not ok:
        return
    x1, y1 = p1
    x2, y2 = p2
    dx, dy = x2 - x1, y2 - y1
    ax, ay = abs(dx), abs(dy)

    def put(x, y):
        if 0 <= x < width and 0 <= y < height:
            img[y, x] = value

    if ax > ay:
        if dx < 0:
            dy = -dy
            x1, x2, y1, y2 = x2, x1, y2, y1
        x_step, y_step = XY_ONE, _c_div(dy << XY_SHIFT, ax | 1)
        ecount = (x2 - x1) >> XY_SHIFT
    else:
        if dy < 0:
            dx = -dx
            x1, x2, y1, y2 = x2, x1, y2, y1
        x_step, y_step = _c_div(dx << XY_SHIFT, ay | 1), XY_ONE
        ecount = (y2 - y1) >> XY_SHIFT
    x1 += XY_ONE >> 1
    y1 += XY_ONE >> 1
    put((x2 + (XY_ONE >> 1)) >> XY_SHIFT, (y2 + (XY_ONE >> 1)) >> XY_SHIFT)
    if ax > ay:
        x = x1 >> XY_SHIFT
        while ecount >= 0:
            put(x, y1 >> XY_SHIFT)
            x += 1
            y1 += y_step
            ecount -= 1
    else:
        y = y1 >> XY_SHIFT
        while ecount >= 0:
            put(x1 >> XY_SHIFT, y)
            x1 += x_step
            y += 1
            ecount -= 1


def line_bresenham(img: np.ndarray, p1_px, p2_px, value=1):
    """Line(img, pt1, pt2, color, 8): clipLine on the pixel end points, then cv::LineIterator
    (8-connected, left to right).  After j major steps the minor coordinate is
    floor((2 * minor_extent * j + major_extent - 1) / (2 * major_extent))."""
    height, width = img.shape
    ok, (x0, y0), (x1, y1) = clip_line(width, height, p1_px, p2_px)
    if not ok:
        return
    if x0 > x1:
        x0, y0, x1, y1 = x1, y1, x0, y0
    dx, dy = x1 - x0, abs(y1 - y0)
    sy = 1 if y1 >= y0 else -1
    steep = dy > dx
    major, minor_ext = (dy, dx) if steep else (dx, dy)
    for j in range(major + 1):
        m = (2 * minor_ext * j + major - 1) // (2 * major) if major else 0
        x, y = (x0 + m, y0 + sy * j) if steep else (x0 + j, y0 + sy * m)
        img[y, x] = value


def fill_convex_poly(img: np.ndarray, v, value=1):
    """FillConvexPoly(img, v, npts, color, LINE_8, shift = XY_SHIFT): outline by Line2 + scan."""
    height, width = img.shape
    npts = len(v)
    delta = XY_ONE >> 1
    delta1 = delta2 = XY_ONE >> 1
    p0 = v[-1]
    xmin = xmax = v[0][0]
    ymin = ymax = v[0][1]
    imin = 0
    for i, p in enumerate(v):
        if p[1] < ymin:
            ymin = p[1]
            imin = i
        ymax = max(ymax, p[1])
        xmax = max(xmax, p[0])
        xmin = min(xmin, p[0])
        line2(img, p0, p, value)
        p0 = p
    xmin = (xmin + delta) >> XY_SHIFT
    xmax = (xmax + delta) >> XY_SHIFT
    ymin = (ymin + delta) >> XY_SHIFT
    ymax = (ymax + delta) >> XY_SHIFT
    if npts < 3 or xmax < 0 or ymax < 0 or xmin >= width or ymin >= height:
        return
    ymax = min(ymax, height - 1)
    edge = [{'idx': imin, 'di': 1, 'x': -XY_ONE, 'dx': 0, 'ye': ymin},
            {'idx': imin, 'di': npts - 1, 'x': -XY_ONE, 'dx': 0, 'ye': ymin}]
    edges = npts
    y = ymin
    while True:
        for e in edge:
            if y >= e['ye']:
                idx0, di = e['idx'], e['di']
                idx = idx0 + di
                if idx >= npts:
                    idx -= npts
                while True:
                    edges -= 1
                    if edges < 0:  # `for (; edges-- > 0; )`
                        break
                    ty = (v[idx][1] + delta) >> XY_SHIFT
                    if ty > y:
                        xs, xe = v[idx0][0], v[idx][0]
                        e['ye'] = ty
                        e['dx'] = _c_div((xe - xs) * 2 + (ty - y), 2 * (ty - y))
                        e['x'] = xs
                        e['idx'] = idx
                        break
                    idx0 = idx
                    idx += di
                    if idx >= npts:
                        idx -= npts
        if edges < 0:
            break
        if y >= 0:
            left, right = (1, 0) if edge[0]['x'] > edge[1]['x'] else (0, 1)
            xx1 = (edge[left]['x'] + delta1) >> XY_SHIFT
            xx2 = (edge[right]['x'] + delta2) >> XY_SHIFT
            if xx2 >= 0 and xx1 < width:
                xx1 = max(xx1, 0)
                xx2 = min(xx2, width - 1)
                if xx1 <= xx2:
                    img[y, xx1:xx2 + 1] = value
        edge[0]['x'] += edge[0]['dx']
        edge[1]['x'] += edge[1]['dx']
        y += 1
        if y > ymax:
            break


def circle_fill(img: np.ndarray, cx: int, cy: int, radius: int, value=1):
    """Circle(img, center, radius, color, fill = 1): midpoint circle, filled by row spans."""
    height, width = img.shape
    err, dx, dy, plus, minus = 0, radius, 0, 1, (radius << 1) - 1

    def hline(y, xa, xb):
        if 0 <= y < height:
            xa, xb = max(xa, 0), min(xb, width - 1)
            if xa <= xb:
                img[y, xa:xb + 1] = value

    while dx >= dy:
        y11, y12, y21, y22 = cy - dy, cy + dy, cy - dx, cy + dx
        x11, x12, x21, x22 = cx - dx, cx + dx, cx - dy, cx + dy
        if x11 < width and x12 >= 0 and y21 < height and y22 >= 0:
            hline(y11, x11, x12)
            hline(y12, x11, x12)
            if x21 < width and x22 >= 0:
                hline(y21, x21, x22)
                hline(y22, x21, x22)
        dy += 1
        err += plus
        plus += 2
        mask = (1 if err <= 0 else 0) - 1
        err -= minus & mask
        dx += mask
        minus -= mask & 2


def thick_line(img: np.ndarray, p0, p1, thickness: int, flags: int, value=1):
    """ThickLine(img, p0, p1, color, thickness, LINE_8, flags, shift = XY_SHIFT)."""
    if thickness <= 1:
        # LINE_8 thin lines: end points rounded to pixels, plain Bresenham (not Line2)
        half = XY_ONE >> 1
        line_bresenham(img, ((p0[0] + half) >> XY_SHIFT, (p0[1] + half) >> XY_SHIFT),
                       ((p1[0] + half) >> XY_SHIFT, (p1[1] + half) >> XY_SHIFT), value)
        return
    dx = (p0[0] - p1[0]) * (1.0 / XY_ONE)
    dy = (p1[1] - p0[1]) * (1.0 / XY_ONE)
    r = dx * dx + dy * dy
    odd = thickness & 1
    thickness <<= XY_SHIFT - 1
    if abs(r) > np.finfo(np.float64).eps:
        r = (thickness + odd * XY_ONE * 0.5) / math.sqrt(r)
        dpx, dpy = cv_round(dy * r), cv_round(dx * r)
        quad = [(p0[0] + dpx, p0[1] + dpy), (p0[0] - dpx, p0[1] - dpy),
                (p1[0] - dpx, p1[1] - dpy), (p1[0] + dpx, p1[1] + dpy)]
        fill_convex_poly(img, quad, value)
    for i in range(2):
        if flags & (i + 1):
            circle_fill(img, (p0[0] + (XY_ONE >> 1)) >> XY_SHIFT, (p0[1] + (XY_ONE >> 1)) >> XY_SHIFT,
                        (thickness + (XY_ONE >> 1)) >> XY_SHIFT, value)
        p0 = p1


def poly_line(img: np.ndarray, v, thickness: int, value=1):
    """PolyLine(img, v, count, is_closed = false, color, thickness, LINE_8, XY_SHIFT)."""
    flags = 3
    p0 = v[0]
    for p in v[1:]:
        thick_line(img, p0, p, thickness, flags, value)
        p0 = p
        flags = 2


def ellipse(img: np.ndarray, center_px, axes_px, thickness: int, value=1):
    """cv.ellipse(img, center, axes, 0, 0, 360, value, thickness) on a uint8 HxW canvas."""
    assert thickness >= 1
    poly_line(img, ellipse_vertices(center_px, axes_px), thickness, value)
    return img
