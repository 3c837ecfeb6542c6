"""Scalar host math shared by the geometric ops: the per-op constants the reference derives
with NumPy / cv2 before any pixel is touched (a few dozen flops per page).

Every function keeps the dtype sequence of the reference line it restates, because these
constants feed the lattice projection whose results are rounded to integers.
"""
import math

import numpy as np


def rodrigues(rvec) -> np.ndarray:
    """cv.Rodrigues(rvec) in double (3x3).  Used by camera.py:96 (float32 out) and inside
    cv.projectPoints (camera.py:189, double).  Plain Python floats: same IEEE operations as the
    NumPy expression c*I + c1*rrt + s*r_x, without the small-array overhead."""
    r0, r1, r2 = float(rvec[0]), float(rvec[1]), float(rvec[2])
    theta = math.sqrt(r0 * r0 + r1 * r1 + r2 * r2)
    if theta < 2.220446049250313e-16:
        return np.eye(3, dtype=np.float64)
    c = math.cos(theta)
    s = math.sin(theta)
    c1 = 1.0 - c
    itheta = 1.0 / theta
    x, y, z = r0 * itheta, r1 * itheta, r2 * itheta
    # element (i, j): c*I + c1*r_i*r_j + s*[r]x, summed left to right like the matrix expression
    return np.array([
        [c * 1.0 + c1 * (x * x) + s * 0.0, c * 0.0 + c1 * (x * y) + s * -z, c * 0.0 + c1 * (x * z) + s * y],
        [c * 0.0 + c1 * (x * y) + s * z, c * 1.0 + c1 * (y * y) + s * 0.0, c * 0.0 + c1 * (y * z) + s * -x],
        [c * 0.0 + c1 * (x * z) + s * -y, c * 0.0 + c1 * (y * z) + s * x, c * 1.0 + c1 * (z * z) + s * 0.0],
    ], dtype=np.float64)


def invert_affine(trans_mat) -> np.ndarray:
    """Inverse of a 2x3 map the way cv::warpAffine derives it (double)."""
    m = np.asarray(trans_mat, dtype=np.float64).copy()
    det = m[0, 0] * m[1, 1] - m[0, 1] * m[1, 0]
    det = 1.0 / det if det != 0 else 0.0
    a11 = m[1, 1] * det
    a22 = m[0, 0] * det
    m[0, 0] = a11
    m[0, 1] *= -det
    m[1, 0] *= -det
    m[1, 1] = a22
    b1 = -m[0, 0] * m[0, 2] - m[0, 1] * m[1, 2]
    b2 = -m[1, 0] * m[0, 2] - m[1, 1] * m[1, 2]
    m[0, 2] = b1
    m[1, 2] = b2
    return m


def _unit_square_to_quad(q):
    (x0, y0), (x1, y1), (x2, y2), (x3, y3) = q
    dx1, dx2, sx = x1 - x2, x3 - x2, x0 - x1 + x2 - x3
    dy1, dy2, sy = y1 - y2, y3 - y2, y0 - y1 + y2 - y3
    den = dx1 * dy2 - dx2 * dy1
    g = (sx * dy2 - dx2 * sy) / den
    h = (dx1 * sy - sx * dy1) / den
    return np.array([
        [x1 - x0 + g * x1, x3 - x0 + h * x3, x0],
        [y1 - y0 + g * y1, y3 - y0 + h * y3, y0],
        [g, h, 1.0],
    ], dtype=np.float64)


def homography_4pt(src_quad, dst_quad) -> np.ndarray:
    """Closed-form stand-in for cv.getPerspectiveTransform(src, dst, DECOMP_SVD) (double)."""
    a = _unit_square_to_quad(np.asarray(src_quad, dtype=np.float64))
    b = _unit_square_to_quad(np.asarray(dst_quad, dtype=np.float64))
    adj = np.array([
        [a[1, 1] * a[2, 2] - a[1, 2] * a[2, 1], a[0, 2] * a[2, 1] - a[0, 1] * a[2, 2],
         a[0, 1] * a[1, 2] - a[0, 2] * a[1, 1]],
        [a[1, 2] * a[2, 0] - a[1, 0] * a[2, 2], a[0, 0] * a[2, 2] - a[0, 2] * a[2, 0],
         a[0, 2] * a[1, 0] - a[0, 0] * a[1, 2]],
        [a[1, 0] * a[2, 1] - a[1, 1] * a[2, 0], a[0, 1] * a[2, 0] - a[0, 0] * a[2, 1],
         a[0, 0] * a[1, 1] - a[0, 1] * a[1, 0]],
    ])
    h = b @ adj
    big = np.abs(h).max()
    if abs(h[2, 2]) > 1e-9 * big:
        return h / h[2, 2]
    return h / big


def lattice_axis(size: int, grid_size: int):
    """create_src_image_grid (grid_creator.py:22-41): range(0, size, g) + [size - 1]."""
    coords = list(range(0, size, grid_size))
    if coords[-1] != size - 1:
        coords.append(size - 1)
    return coords
