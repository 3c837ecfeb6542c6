// vkb_draw.cuh -- the OpenCV drawing primitives behind cv.ellipse(..., thickness >= 1, LINE_8) as
// `__host__ __device__` code (imgproc/src/drawing.cpp of OpenCV 4.13: ThickLine, Line / LineIterator,
// Line2, FillConvexPoly, Circle, clipLine), used by the ellipse_streak kernel
// (vkit/mechanism/distortion/photometric/streak.py:282-337).  tests/hostsim compiles this header
// with g++ and checks it against the oracle (oracle/cv2_draw.py) and cv2 itself.
//
// Pixels are produced through two callbacks: plot(x, y) and hline(y, xa, xb) (inclusive span,
// already clipped to the canvas).  Every write means "mask = 1", so the order of the primitives and
// of the segments of an ellipse does not matter: one thread per segment, no ordering between them.
#pragma once
#include <stdint.h>
#include <math.h>
#include "vkb_math.cuh"

namespace vkb {

constexpr int kXyShift = 16;
constexpr long long kXyOne = 1ll << kXyShift;

struct DrawPoint {
    long long x, y;  // 16.16 fixed point (or pixels where a function says so)
};

// clipLine(Size2l, Point2l&, Point2l&)
VKB_HD bool draw_clip_line(long long width, long long height, DrawPoint& p1, DrawPoint& p2) {
    const long long right = width - 1, bottom = height - 1;
    if (width <= 0 || height <= 0) return false;
    long long &x1 = p1.x, &y1 = p1.y, &x2 = p2.x, &y2 = p2.y;
    int c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
    int c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
        long long a;
        if (c1 & 12) {
            a = c1 < 8 ? 0 : bottom;
            x1 += (long long)((double)(a - y1) * (double)(x2 - x1) / (double)(y2 - y1));
            y1 = a;
            c1 = (x1 < 0) + (x1 > right) * 2;
        }
        if (c2 & 12) {
            a = c2 < 8 ? 0 : bottom;
            x2 += (long long)((double)(a - y2) * (double)(x2 - x1) / (double)(y2 - y1));
            y2 = a;
            c2 = (x2 < 0) + (x2 > right) * 2;
        }
        if ((c1 & c2) == 0 && (c1 | c2) != 0) {
            if (c1) {
                a = c1 == 1 ? 0 : right;
                y1 += (long long)((double)(a - x1) * (double)(y2 - y1) / (double)(x2 - x1));
                x1 = a;
                c1 = 0;
            }
            if (c2) {
                a = c2 == 1 ? 0 : right;
                y2 += (long long)((double)(a - x2) * (double)(y2 - y1) / (double)(x2 - x1));
                x2 = a;
                c2 = 0;
            }
        }
    }
    return (c1 | c2) == 0;
}

// Line(img, pt1, pt2, color, 8): pixel end points, clipped, cv::LineIterator left to right
template <typename Plot>
VKB_HD void draw_line_bresenham(int width, int height, DrawPoint p1, DrawPoint p2, Plot plot) {
    if (!draw_clip_line(width, height, p1, p2)) return;
    edge_walk((int)p1.x, (int)p1.y, (int)p2.x, (int)p2.y, plot);
}

// Line2: fixed-point DDA between two 16.16 points
template <typename Plot>
VKB_HD void draw_line2(int width, int height, DrawPoint p1, DrawPoint p2, Plot plot) {
    if (!draw_clip_line((long long)width << kXyShift, (long long)height << kXyShift, p1, p2)) return;
    long long dx = p2.x - p1.x, dy = p2.y - p1.y;
    const long long ax = dx < 0 ? -dx : dx, ay = dy < 0 ? -dy : dy;
    long long x_step, y_step;
    int ecount;
    if (ax > ay) {
        if (dx < 0) {
            dy = -dy;
            const DrawPoint t = p1; p1 = p2; p2 = t;
        }
        x_step = kXyOne;
        y_step = (dy << kXyShift) / (ax | 1);
        ecount = (int)((p2.x - p1.x) >> kXyShift);
    } else {
        if (dy < 0) {
            dx = -dx;
            const DrawPoint t = p1; p1 = p2; p2 = t;
        }
        x_step = (dx << kXyShift) / (ay | 1);
        y_step = kXyOne;
        ecount = (int)((p2.y - p1.y) >> kXyShift);
    }
    p1.x += kXyOne >> 1;
    p1.y += kXyOne >> 1;
    auto put = [&](int x, int y) {
        if (0 <= x && x < width && 0 <= y && y < height) plot(x, y);
    };
    put((int)((p2.x + (kXyOne >> 1)) >> kXyShift), (int)((p2.y + (kXyOne >> 1)) >> kXyShift));
    if (ax > ay) {
        long long x = p1.x >> kXyShift;
        while (ecount >= 0) {
            put((int)x, (int)(p1.y >> kXyShift));
            ++x;
            p1.y += y_step;
            --ecount;
        }
    } else {
        long long y = p1.y >> kXyShift;
        while (ecount >= 0) {
            put((int)(p1.x >> kXyShift), (int)y);
            p1.x += x_step;
            ++y;
            --ecount;
        }
    }
    (void)x_step;
    (void)y_step;
}

// FillConvexPoly(img, v, 4, color, LINE_8, shift = XY_SHIFT)
template <typename Plot, typename HLine>
VKB_HD void draw_fill_convex_quad(int width, int height, const DrawPoint* v, Plot plot, HLine hline) {
    constexpr int npts = 4;
    const long long delta = kXyOne >> 1;
    DrawPoint p0 = v[npts - 1];
    long long xmin = v[0].x, xmax = v[0].x, ymin = v[0].y, ymax = v[0].y;
    int imin = 0;
    for (int i = 0; i < npts; ++i) {
        const DrawPoint p = v[i];
        if (p.y < ymin) {
            ymin = p.y;
            imin = i;
        }
        ymax = p.y > ymax ? p.y : ymax;
        xmax = p.x > xmax ? p.x : xmax;
        xmin = p.x < xmin ? p.x : xmin;
        draw_line2(width, height, p0, p, plot);
        p0 = p;
    }
    xmin = (xmin + delta) >> kXyShift;
    xmax = (xmax + delta) >> kXyShift;
    ymin = (ymin + delta) >> kXyShift;
    ymax = (ymax + delta) >> kXyShift;
    if ((int)xmax < 0 || (int)ymax < 0 || (int)xmin >= width || (int)ymin >= height) return;
    if (ymax > height - 1) ymax = height - 1;
    int e_idx[2] = {imin, imin}, e_di[2] = {1, npts - 1}, e_ye[2] = {(int)ymin, (int)ymin};
    long long e_x[2] = {-kXyOne, -kXyOne}, e_dx[2] = {0, 0};
    int edges = npts;
    int y = (int)ymin;
    do {
        for (int i = 0; i < 2; ++i) {
            if (y >= e_ye[i]) {
                int idx0 = e_idx[i];
                const int di = e_di[i];
                int idx = idx0 + di;
                if (idx >= npts) idx -= npts;
                for (; edges-- > 0;) {
                    const int ty = (int)((v[idx].y + delta) >> kXyShift);
                    if (ty > y) {
                        const long long xs = v[idx0].x, xe = v[idx].x;
                        e_ye[i] = ty;
                        e_dx[i] = ((xe - xs) * 2 + (ty - y)) / (2 * (ty - y));
                        e_x[i] = xs;
                        e_idx[i] = idx;
                        break;
                    }
                    idx0 = idx;
                    idx += di;
                    if (idx >= npts) idx -= npts;
                }
            }
        }
        if (edges < 0) break;
        if (y >= 0) {
            const int left = e_x[0] > e_x[1] ? 1 : 0, right = 1 - left;
            int xx1 = (int)((e_x[left] + delta) >> kXyShift);
            int xx2 = (int)((e_x[right] + delta) >> kXyShift);
            if (xx2 >= 0 && xx1 < width) {
                if (xx1 < 0) xx1 = 0;
                if (xx2 >= width) xx2 = width - 1;
                if (xx1 <= xx2) hline(y, xx1, xx2);
            }
        }
        e_x[0] += e_dx[0];
        e_x[1] += e_dx[1];
    } while (++y <= (int)ymax);
}

// Circle(img, center, radius, color, fill = 1)
template <typename HLine>
VKB_HD void draw_circle_fill(int width, int height, int cx, int cy, int radius, HLine hline) {
    int err = 0, dx = radius, dy = 0, plus = 1, minus = (radius << 1) - 1;
    auto span = [&](int y, int xa, int xb) {
        if ((unsigned)y >= (unsigned)height) return;
        if (xa < 0) xa = 0;
        if (xb > width - 1) xb = width - 1;
        if (xa <= xb) hline(y, xa, xb);
    };
    while (dx >= dy) {
        const int y11 = cy - dy, y12 = cy + dy, y21 = cy - dx, y22 = cy + dx;
        const int x11 = cx - dx, x12 = cx + dx, x21 = cx - dy, x22 = cx + dy;
        if (x11 < width && x12 >= 0 && y21 < height && y22 >= 0) {
            span(y11, x11, x12);
            span(y12, x11, x12);
            if (x21 < width && x22 >= 0) {
                span(y21, x21, x22);
                span(y22, x21, x22);
            }
        }
        ++dy;
        err += plus;
        plus += 2;
        const int mask = (err <= 0) - 1;
        err -= minus & mask;
        dx += mask;
        minus -= mask & 2;
    }
}

// ThickLine(img, p0, p1, color, thickness, LINE_8, flags, shift = XY_SHIFT); flags: bit 0 / bit 1 =
// round cap at p0 / p1
template <typename Plot, typename HLine>
VKB_HD void draw_thick_segment(int width, int height, DrawPoint p0, DrawPoint p1, int thickness,
                               int flags, Plot plot, HLine hline) {
    const long long half = kXyOne >> 1;
    if (thickness <= 1) {
        // thin LINE_8 lines: end points rounded to pixels, plain Bresenham
        const DrawPoint a = {(p0.x + half) >> kXyShift, (p0.y + half) >> kXyShift};
        const DrawPoint b = {(p1.x + half) >> kXyShift, (p1.y + half) >> kXyShift};
        draw_line_bresenham(width, height, a, b, plot);
        return;
    }
    const double inv = 1.0 / (double)kXyOne;
    const double dx = (double)(p0.x - p1.x) * inv, dy = (double)(p1.y - p0.y) * inv;
    double r = dx * dx + dy * dy;
    const int odd = thickness & 1;
    const long long th = (long long)thickness << (kXyShift - 1);
    if (fabs(r) > 2.220446049250313e-16) {
        r = ((double)th + (double)odd * (double)kXyOne * 0.5) / sqrt(r);
        const long long dpx = (long long)rint(dy * r), dpy = (long long)rint(dx * r);
        const DrawPoint quad[4] = {{p0.x + dpx, p0.y + dpy}, {p0.x - dpx, p0.y - dpy},
                                   {p1.x - dpx, p1.y - dpy}, {p1.x + dpx, p1.y + dpy}};
        draw_fill_convex_quad(width, height, quad, plot, hline);
    }
    for (int i = 0; i < 2; ++i) {
        if (flags & (i + 1)) {
            draw_circle_fill(width, height, (int)((p0.x + half) >> kXyShift),
                             (int)((p0.y + half) >> kXyShift), (int)((th + half) >> kXyShift), hline);
        }
        p0 = p1;
    }
}

}  // namespace vkb
