// vkb_draw_host.h -- host half of cv.ellipse: the vertex list of EllipseEx / ellipse2Poly
// (imgproc/src/drawing.cpp, OpenCV 4.13) for angle = 0, a full 0..360 arc.  A few dozen scalar
// operations per ellipse (parameter preparation, like the homography of an affine op); the pixels
// are drawn on the device from the resulting segments (vkb_draw.cuh).  Shared with tests/hostsim.
#pragma once
#include <math.h>
#include <stdint.h>
#include <vector>
#include "vkb_draw.cuh"

namespace vkb {

// SinTable of drawing.cpp: sin of 0 .. 450 degrees written as float literals with 7 decimals
inline const float* draw_sin_table() {
    static float table[451];
    static bool ready = false;
    if (!ready) {
        for (int d = 0; d <= 450; ++d) {
            const double s = sin((double)d * 3.14159265358979323846 / 180.0);
            table[d] = (float)(round(s * 1e7) / 1e7);
        }
        ready = true;
    }
    return table;
}

// center / axes in pixels; vertices in 16.16 fixed point, consecutive duplicates removed
inline void ellipse_vertices(int cx_px, int cy_px, int ax_px, int ay_px, std::vector<DrawPoint>& out) {
    const float* S = draw_sin_table();
    const long long cx = (long long)cx_px << kXyShift, cy = (long long)cy_px << kXyShift;
    const long long ax = (long long)(ax_px < 0 ? -ax_px : ax_px) << kXyShift;
    const long long ay = (long long)(ay_px < 0 ? -ay_px : ay_px) << kXyShift;
    int delta = (int)(((ax > ay ? ax : ay) + (kXyOne >> 1)) >> kXyShift);
    delta = delta < 3 ? 90 : delta < 10 ? 30 : delta < 15 ? 18 : 5;
    const double alpha = (double)S[450], beta = (double)S[0];
    out.clear();
    bool have_prev = false;
    DrawPoint prev = {0, 0};
    for (int i = 0; i < 360 + delta; i += delta) {
        const int angle = i > 360 ? 360 : i;
        const double x = (double)ax * (double)S[450 - angle];
        const double y = (double)ay * (double)S[angle];
        const double vx = (double)cx + x * alpha - y * beta;
        const double vy = (double)cy + x * beta + y * alpha;
        DrawPoint pt;
        pt.x = (long long)lrint(vx / (double)kXyOne) << kXyShift;
        pt.y = (long long)lrint(vy / (double)kXyOne) << kXyShift;
        pt.x += (long long)lrint(vx - (double)pt.x);
        pt.y += (long long)lrint(vy - (double)pt.y);
        if (!have_prev || pt.x != prev.x || pt.y != prev.y) {
            out.push_back(pt);
            prev = pt;
            have_prev = true;
        }
    }
    if (out.size() <= 1) {
        out.assign(2, DrawPoint{cx, cy});
    }
}

struct DrawSegment {
    DrawPoint p0, p1;
    int32_t flags;  // bit 0 / bit 1: round cap at p0 / p1 (thick lines)
    int32_t pad;
};

// PolyLine(img, v, count, is_closed = false, ...): the segments of one ellipse, appended
inline void ellipse_segments(int cx_px, int cy_px, int ax_px, int ay_px, std::vector<DrawSegment>& segs) {
    std::vector<DrawPoint> v;
    ellipse_vertices(cx_px, cy_px, ax_px, ay_px, v);
    int flags = 3;
    for (size_t i = 1; i < v.size(); ++i) {
        segs.push_back(DrawSegment{v[i - 1], v[i], flags, 0});
        flags = 2;
    }
}

}  // namespace vkb
