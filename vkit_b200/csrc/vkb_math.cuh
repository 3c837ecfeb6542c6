// vkb_math.cuh -- per-element numerics of the distortion / blend path.
//
// Every function here is `__host__ __device__`: the CUDA kernels in this directory call them
// per pixel / per lattice point / per cell-row, and `tests/hostsim/` compiles the very same
// header with g++ so the arithmetic can be checked against the oracle on a machine without a
// GPU.  The host build is a test harness only; nothing in the product path runs on the CPU.
//
// The arithmetic restates what the reference (vkit-x/vkit @ 98ada2d) obtains from OpenCV
// 4.13: 1/32-pixel coordinate quantisation, uint8 bilinear with integer
// weights summing to 2^15, float32 bilinear without FMA contraction, the fixed-point affine
// coordinate pipeline of cv::warpAffine, cv::fillPoly coverage (8-connected Bresenham outline
// plus 16.16 scan fill) and closed-form 4-point homographies.
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define VKB_HD __host__ __device__ __forceinline__
#else
#define VKB_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define VKB_FMUL(a, b) __fmul_rn((a), (b))
#define VKB_FADD(a, b) __fadd_rn((a), (b))
#define VKB_FSUB(a, b) __fsub_rn((a), (b))
#define VKB_DMUL(a, b) __dmul_rn((a), (b))
#define VKB_DADD(a, b) __dadd_rn((a), (b))
#define VKB_FRCP_APPROX(x) __fdividef(1.0f, (x))
#else
#define VKB_FMUL(a, b) ((a) * (b))
#define VKB_FADD(a, b) ((a) + (b))
#define VKB_FSUB(a, b) ((a) - (b))
#define VKB_DMUL(a, b) ((a) * (b))
#define VKB_DADD(a, b) ((a) + (b))
#define VKB_FRCP_APPROX(x) (1.0f / (x))
#endif

namespace vkb {

constexpr int kInterBits = 5;
constexpr int kInterTab = 32;

// ---------------------------------------------------------------------------------------
// Rounding helpers.
// ---------------------------------------------------------------------------------------
// cvRound / saturate_cast<int>(double): round half to even, out-of-range and NaN -> INT_MIN
// (what cvtsd2si returns), which lands outside every image and therefore samples the border.
VKB_HD int cv_round_d(double v) {
    if (!(v == v)) return INT32_MIN;
    double r = rint(v);
    if (r >= 2147483648.0 || r < -2147483648.0) return INT32_MIN;
    return (int)r;
}

// float map value -> 1/32 px fixed point exactly like cv::remap (value * 32 is exact in float).
VKB_HD int map_to_fixed(float m) {
    return cv_round_d((double)m * 32.0);
}

VKB_HD int clamp_short(int v) {
    return v < -32768 ? -32768 : (v > 32767 ? 32767 : v);
}

// ---------------------------------------------------------------------------------------
// Bilinear taps, BORDER_CONSTANT(0).  X, Y are 1/32 px fixed point source coordinates.
// ---------------------------------------------------------------------------------------
template <int C>
VKB_HD void bilinear_u8(const uint8_t* __restrict__ src, int h, int w, long long pitch, int X,
                        int Y, uint8_t* out) {
    const int x0 = clamp_short(X >> kInterBits);
    const int y0 = clamp_short(Y >> kInterBits);
    const int fx = X & (kInterTab - 1);
    const int fy = Y & (kInterTab - 1);
    const bool in_x0 = (unsigned)x0 < (unsigned)w;
    const bool in_x1 = (unsigned)(x0 + 1) < (unsigned)w;
    const bool in_y0 = (unsigned)y0 < (unsigned)h;
    const bool in_y1 = (unsigned)(y0 + 1) < (unsigned)h;
    const uint8_t* r0 = src + (long long)y0 * pitch + (long long)x0 * C;
    const uint8_t* r1 = r0 + pitch;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const int p00 = (in_y0 && in_x0) ? r0[c] : 0;
        const int p01 = (in_y0 && in_x1) ? r0[C + c] : 0;
        const int p10 = (in_y1 && in_x0) ? r1[c] : 0;
        const int p11 = (in_y1 && in_x1) ? r1[C + c] : 0;
        // Weights (32-fy)(32-fx)*32 ... share the factor 32, so
        // (sum p*w + 2^14) >> 15 == (t + 512) >> 10 with t below (max 255 * 1024).
        const int a = (kInterTab - fx) * p00 + fx * p01;
        const int b = (kInterTab - fx) * p10 + fx * p11;
        const int t = (kInterTab - fy) * a + fy * b;
        out[c] = (uint8_t)((t + 512) >> 10);
    }
}

VKB_HD float bilinear_f32(const float* __restrict__ src, int h, int w, long long pitch_elems,
                          int X, int Y) {
    const int x0 = clamp_short(X >> kInterBits);
    const int y0 = clamp_short(Y >> kInterBits);
    const int fx = X & (kInterTab - 1);
    const int fy = Y & (kInterTab - 1);
    const bool in_x0 = (unsigned)x0 < (unsigned)w;
    const bool in_x1 = (unsigned)(x0 + 1) < (unsigned)w;
    const bool in_y0 = (unsigned)y0 < (unsigned)h;
    const bool in_y1 = (unsigned)(y0 + 1) < (unsigned)h;
    const float* r0 = src + (long long)y0 * pitch_elems + x0;
    const float* r1 = r0 + pitch_elems;
    const float p00 = (in_y0 && in_x0) ? r0[0] : 0.f;
    const float p01 = (in_y0 && in_x1) ? r0[1] : 0.f;
    const float p10 = (in_y1 && in_x0) ? r1[0] : 0.f;
    const float p11 = (in_y1 && in_x1) ? r1[1] : 0.f;
    const float ax = (float)fx * 0.03125f;  // exact
    const float ay = (float)fy * 0.03125f;
    const float bx = 1.0f - ax;  // exact
    const float by = 1.0f - ay;
    const float w00 = VKB_FMUL(by, bx);
    const float w01 = VKB_FMUL(by, ax);
    const float w10 = VKB_FMUL(ay, bx);
    const float w11 = VKB_FMUL(ay, ax);
    float acc = VKB_FMUL(p00, w00);
    acc = VKB_FADD(acc, VKB_FMUL(p01, w01));
    acc = VKB_FADD(acc, VKB_FMUL(p10, w10));
    acc = VKB_FADD(acc, VKB_FMUL(p11, w11));
    return acc;
}

// ---------------------------------------------------------------------------------------
// Coordinate pipelines.
// ---------------------------------------------------------------------------------------
// cv::warpAffine: Mi = inverse map (2x3, double, row major).  AB_BITS = 10 fixed point.
VKB_HD void affine_coord(const double* __restrict__ Mi, int x, int y, int& X, int& Y) {
    const int adelta = cv_round_d(VKB_DMUL(VKB_DMUL(Mi[0], (double)x), 1024.0));
    const int bdelta = cv_round_d(VKB_DMUL(VKB_DMUL(Mi[3], (double)x), 1024.0));
    const int X0 = cv_round_d(VKB_DMUL(VKB_DADD(VKB_DMUL(Mi[1], (double)y), Mi[2]), 1024.0)) + 16;
    const int Y0 = cv_round_d(VKB_DMUL(VKB_DADD(VKB_DMUL(Mi[4], (double)y), Mi[5]), 1024.0)) + 16;
    X = (X0 + adelta) >> 5;
    Y = (Y0 + bdelta) >> 5;
}

// cv::warpPerspective: Mi = inverse map (3x3, double).
VKB_HD void perspective_coord(const double* __restrict__ Mi, int x, int y, int& X, int& Y) {
    const double xd = (double)x, yd = (double)y;
    double W = VKB_DADD(VKB_DADD(VKB_DMUL(Mi[6], xd), VKB_DMUL(Mi[7], yd)), Mi[8]);
    W = (W != 0.0) ? 32.0 / W : 0.0;
    double fX = VKB_DMUL(VKB_DADD(VKB_DADD(VKB_DMUL(Mi[0], xd), VKB_DMUL(Mi[1], yd)), Mi[2]), W);
    double fY = VKB_DMUL(VKB_DADD(VKB_DADD(VKB_DMUL(Mi[3], xd), VKB_DMUL(Mi[4], yd)), Mi[5]), W);
    fX = fmax(-2147483648.0, fmin(2147483647.0, fX));
    fY = fmax(-2147483648.0, fmin(2147483647.0, fY));
    X = cv_round_d(fX);
    Y = cv_round_d(fY);
}

// Grid remap: source coordinate of dst pixel (x, y) under a cell's inverse homography H
// (double), stored as float32 like the reference's map_x/map_y, then quantised like cv::remap.
// (vkit grid_rendering/type.py:231-256 followed by grid_blender.py:60.)
VKB_HD int map_to_fixed_f(float m) {
    // float * 32 is exact (or overflows to inf); round half to even; NaN / out of range ->
    // a coordinate far outside every image, like cvRound's INT_MIN.
    const float t = m * 32.0f;
    if (!(fabsf(t) < 2.0e9f)) return INT32_MIN;
    return (int)rintf(t);
}

VKB_HD void cell_coord(const double* __restrict__ H, int x, int y, int& X, int& Y) {
    const double xd = (double)x, yd = (double)y;
    const double den = fma(H[6], xd, fma(H[7], yd, H[8]));
    const double nx = fma(H[0], xd, fma(H[1], yd, H[2]));
    const double ny = fma(H[3], xd, fma(H[4], yd, H[5]));
    if (den == 0.0) {  // the reference skips such pixels; they keep map value 0
        X = 0;
        Y = 0;
        return;
    }
    const float mx = (float)(nx / den);
    const float my = (float)(ny / den);
    X = map_to_fixed_f(mx);
    Y = map_to_fixed_f(my);
}

// ---------------------------------------------------------------------------------------
// Error-bounded float32 evaluation of cell_coord.
//
// The source lattice is integer, so inside a cell the source coordinate is  u = sx0 + du  with
// du in roughly [-2, g+2].  f = 32*du is evaluated in float32 from the homography re-centred on
// the origin of the 32 x 32 dst tile the pixel lies in (tile origin -> offset from the cell's
// src corner): magnitudes stay in the low thousands, so the absolute error of f is far below the
// 1/32 px quantum.
//
// The reference rounds twice: u (float64) -> float32 map value m, then X = rint(32 * m), half
// to even.  With t = 32 * u, K an integer and T = K + 0.5: every t within h = 32 * halfulp32(u)
// of T rounds to the float32 value T itself, which rint() then sends to the EVEN neighbour.
// So whatever the parity of the neighbours, a t farther than h from every tie rounds like
// rint(t), and an estimate of t with error below eps (kFastSlack) gives the reference's X
// whenever it keeps a distance of h + eps from every tie; otherwise the caller takes the
// float64 path.
//
// The test runs on integers.  The numerators are pre-scaled by 32 * 2^kFastBits, the last FMA of
// an axis adds 1.5 * 2^23 to numerator * reciprocal, so the low mantissa bits of the sum are
// I = rint(2^kFastBits * f) (the rounding, at most half a unit = 1.2e-4, is part of eps).  With
// the cell's base  B = 2^kFastBits * 32 * s0 + 2^(kFastBits-1) - margin - (bits of the magic)
// t = bits + B  has  X = t >> kFastBits  and  p = t & (2^kFastBits - 1)  = distance of the estimate
// from the tie below minus `margin`: accepted when p <= 2^kFastBits - 1 - 2 * margin.  One FMA,
// one add, one shift and one AND per axis, one max, two logic ops and a subtraction per pixel
// whose sign is the verdict (the float form of the same test -- two parities, two thresholds,
// a range test -- cost twenty per pixel).  margin = ceil((h_max + eps) * 2^kFastBits) for the
// largest source coordinate of the page.  tests/test_hostsim.py audits all of this on every
// golden grid case (observed evaluation error incl. the rounding is > 3x below kFastSlack;
// 0 wrong results among accepted pixels).
// ---------------------------------------------------------------------------------------
struct CellLocal {
    float a0, a1, a2, g;  // kFastUnits * 32 * numerator of du_x = a0*x' + a1*y' + a2 ; denominator g*x' + h*y' + 1
    float b0, b1, b2, h;  // kFastUnits * 32 * numerator of du_y          (x', y') = pixel - tile origin
};

constexpr float kFastSlack = 1.0e-3f;          // in units of 1/32 px
constexpr float kFastRange = 1024.0f;          // |32*du| accepted by the fast path (32 px)
constexpr int kFastBits = 12;                  // fraction bits of the fast path's fixed point
constexpr int kFastOne = 1 << kFastBits;
constexpr float kFastUnits = (float)kFastOne;
constexpr float kRoundMagic = 12582912.0f;     // 1.5 * 2^23: (v + magic) rounds v half to even
constexpr int kRoundMagicBits = 0x4B400000;    // bit pattern of kRoundMagic
constexpr int kFastMaxExtent = 16000;          // 32 * extent * 2^kFastBits must fit an int32
static_assert(kFastRange * kFastUnits <= 4194304.0f, "the magic-number rounding holds below 2^22");

// 2^(e-19) for |u| in [2^e, 2^(e+1)): 32 * half ulp of float32(u); 0 for tiny |u|.
VKB_HD float half_ulp_times_32(float u) {
    union { float f; uint32_t i; } v;
    v.f = u;
    const uint32_t e = v.i & 0x7f800000u;
    v.i = e > (19u << 23) ? e - (19u << 23) : 0u;
    return v.f;
}

// Acceptance window for source coordinates that start below `extent` pixels (a page's size, or
// the largest source corner among the cells of one dst tile -- the half ulp of the float32 map
// value grows with the coordinate, so tiles near the origin get a narrower window): distance
// (in fast-path units) an estimate keeps from every tie.  < kFastOne / 4 for every extent
// below kFastMaxExtent; pages beyond that never take the fast path (fast_page_ok).
VKB_HD int fast_margin(int extent) {
    const float reach = (float)(extent < kFastMaxExtent ? extent : kFastMaxExtent) + kFastRange / 32.0f + 1.0f;
    const float w = (half_ulp_times_32(reach) + kFastSlack) * kFastUnits;
    const int m = (int)w;
    return (float)m < w ? m + 1 : m;
}
VKB_HD int fast_limit(int margin) { return kFastOne - 1 - 2 * margin; }
VKB_HD bool fast_page_ok(int extent) { return extent < kFastMaxExtent; }

// base of a cell whose source corner is s0 (pixels), for a page with acceptance margin `margin`
VKB_HD int fast_base(int s0, int margin) {
    return s0 * 32 * kFastOne + kFastOne / 2 - margin - kRoundMagicBits;
}
// base of the zero map (uncovered pixels): accepted under every margin, X = 0
constexpr int kFastBaseZero = kFastOne / 2 - kRoundMagicBits;

// H: inverse homography of the cell (dst -> src); (sx0, sy0): its src corner; (ox, oy): origin
// of the dst tile the form is valid for.
VKB_HD void make_cell_local(const double* __restrict__ H, int sx0, int sy0, int ox, int oy,
                            CellLocal& L) {
    const double ax0 = H[0] - sx0 * H[6], ax1 = H[1] - sx0 * H[7], ax2 = H[2] - sx0 * H[8];
    const double ay0 = H[3] - sy0 * H[6], ay1 = H[4] - sy0 * H[7], ay2 = H[5] - sy0 * H[8];
    const double d0 = H[6] * ox + H[7] * oy + H[8];
    const double inv = 1.0 / d0;  // inf / nan when degenerate: the fast path then always fails
    const double invs = (32.0 * kFastOne) * inv;
    L.a0 = (float)(ax0 * invs);
    L.a1 = (float)(ax1 * invs);
    L.a2 = (float)((ax0 * ox + ax1 * oy + ax2) * invs);
    L.b0 = (float)(ay0 * invs);
    L.b1 = (float)(ay1 * invs);
    L.b2 = (float)((ay0 * ox + ay1 * oy + ay2) * invs);
    L.g = (float)(H[6] * inv);
    L.h = (float)(H[7] * inv);
}

VKB_HD float fma_rn_f32(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
    return __fmaf_rn(a, b, c);
#else
    return fmaf(a, b, c);  // correctly rounded on the host as well
#endif
}

// one axis: n * r = float32 estimate of kFastUnits * 32 * du, base = fast_base(s0, margin).
// X = 32 * s0 + rint(32 * du) when the axis is accepted; returns p, the position inside the
// acceptance window (accepted: p <= limit); bits = the sum itself for the range test.
VKB_HD int fast_axis(float n, float r, int base, int& X, int& bits) {
    union { float f; int i; } v;
    v.f = fma_rn_f32(n, r, kRoundMagic);  // low mantissa bits = rint(n * r), half to even
    bits = v.i;
    const int t = v.i + base;
    X = t >> kFastBits;
    return t & (kFastOne - 1);
}
// Verdict of a pixel from both axes: >= 0 accepted, < 0 rejected.  A sum whose exponent is not
// the magic number's (|32*du| >= kFastRange, inf, nan) sets a bit above the window positions.
VKB_HD int fast_verdict(int px, int py, int bits_x, int bits_y, int limit) {
    const int p = px > py ? px : py;
    const int c = p | (((bits_x ^ 0x4B000000) | (bits_y ^ 0x4B000000)) & 0x7F800000);  // < 2^31
    return limit - c;
}

// The three linear forms are evaluated column part first: the part that depends on the pixel's
// column only (CellColumn) is shared by every row a thread visits inside one cell, a row then
// costs three FMAs.  Two roundings per form, like the row-first order.
struct CellColumn {
    float nx, ny, d;  // a0*x' + a2, b0*x' + b2, g*x' + 1
};

VKB_HD void cell_column(const CellLocal& L, float xr, CellColumn& c) {
    c.nx = fma_rn_f32(L.a0, xr, L.a2);
    c.ny = fma_rn_f32(L.b0, xr, L.b2);
    c.d = fma_rn_f32(L.g, xr, 1.0f);
}

// one row of a prepared column: (a1, b1, h) = L.a1, L.b1, L.h.  Returns the verdict (>= 0:
// accepted, the sign bit marks a pixel for the float64 path).
VKB_HD int cell_coord_fast_row(const CellColumn& c, float a1, float b1, float h, float yr, int xb,
                               int yb, int limit, int& X, int& Y) {
    const float d = fma_rn_f32(h, yr, c.d);
    const float nx = fma_rn_f32(a1, yr, c.nx);
    const float ny = fma_rn_f32(b1, yr, c.ny);
#if defined(__CUDA_ARCH__)
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));  // one MUFU.RCP, <= 1 ulp
#else
    const float r = 1.0f / d;
#endif
    int bx, by;
    const int px = fast_axis(nx, r, xb, X, bx);
    const int py = fast_axis(ny, r, yb, Y, by);
    return fast_verdict(px, py, bx, by, limit);
}

// (xr, yr): pixel - tile origin as floats (exact small integers).  Verdict as above.
VKB_HD int cell_coord_fast(const CellLocal& L, float xr, float yr, int xb, int yb, int limit,
                           int& X, int& Y) {
    CellColumn c;
    cell_column(L, xr, c);
    return cell_coord_fast_row(c, L.a1, L.b1, L.h, yr, xb, yb, limit, X, Y);
}

// collects the sign bit of a verdict: acc = acc << 1 | (verdict < 0)
VKB_HD uint32_t fast_collect(uint32_t acc, int verdict) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_l((uint32_t)verdict, acc, 1);
#else
    return (acc << 1) | ((uint32_t)verdict >> 31);
#endif
}

// ---------------------------------------------------------------------------------------
// Homography through 4 point pairs, closed form (Heckbert), double precision.
// q = [x0,y0,x1,y1,x2,y2,x3,y3].  Result maps src quad -> dst quad, H[8] = 1.
// Stands in for cv.getPerspectiveTransform(src, dst, DECOMP_SVD) (type.py:172,189).
// ---------------------------------------------------------------------------------------
VKB_HD void unit_square_to_quad(const double* q, double* A) {
    const double x0 = q[0], y0 = q[1], x1 = q[2], y1 = q[3], x2 = q[4], y2 = q[5], x3 = q[6],
                 y3 = q[7];
    const double dx1 = x1 - x2, dx2 = x3 - x2, sx = x0 - x1 + x2 - x3;
    const double dy1 = y1 - y2, dy2 = y3 - y2, sy = y0 - y1 + y2 - y3;
    const double den = dx1 * dy2 - dx2 * dy1;
    const double g = (sx * dy2 - dx2 * sy) / den;
    const double h = (dx1 * sy - sx * dy1) / den;
    A[0] = x1 - x0 + g * x1;
    A[1] = x3 - x0 + h * x3;
    A[2] = x0;
    A[3] = y1 - y0 + g * y1;
    A[4] = y3 - y0 + h * y3;
    A[5] = y0;
    A[6] = g;
    A[7] = h;
    A[8] = 1.0;
}

VKB_HD void homography_4pt(const double* src_quad, const double* dst_quad, double* H) {
    double A[9], B[9], J[9];
    unit_square_to_quad(src_quad, A);
    unit_square_to_quad(dst_quad, B);
    // adj(A): maps src quad -> unit square up to scale.
    J[0] = A[4] * A[8] - A[5] * A[7];
    J[1] = A[2] * A[7] - A[1] * A[8];
    J[2] = A[1] * A[5] - A[2] * A[4];
    J[3] = A[5] * A[6] - A[3] * A[8];
    J[4] = A[0] * A[8] - A[2] * A[6];
    J[5] = A[2] * A[3] - A[0] * A[5];
    J[6] = A[3] * A[7] - A[4] * A[6];
    J[7] = A[1] * A[6] - A[0] * A[7];
    J[8] = A[0] * A[4] - A[1] * A[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            H[r * 3 + c] = B[r * 3 + 0] * J[0 * 3 + c] + B[r * 3 + 1] * J[1 * 3 + c]
                           + B[r * 3 + 2] * J[2 * 3 + c];
        }
    }
    // cv2 fixes H[8] = 1.  When the cell's vanishing line passes (numerically) through the
    // canvas origin H[8] is ~0 and that normalisation blows up, so fall back to the largest
    // entry; a homography is scale free, N/D per pixel is unaffected.
    double big = 0.0;
#pragma unroll
    for (int i = 0; i < 9; ++i) big = fmax(big, fabs(H[i]));
    if (fabs(H[8]) > 1e-9 * big) {
        const double inv = 1.0 / H[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) H[i] *= inv;
        H[8] = 1.0;
    } else if (big > 0.0) {
        const double inv = 1.0 / big;
#pragma unroll
        for (int i = 0; i < 9; ++i) H[i] *= inv;
    }
}

// ---------------------------------------------------------------------------------------
// cv::fillPoly coverage of one polygon restricted to one row, as a bit mask.
//
// Coverage = outline (cv::LineIterator, 8-connected, left-to-right) UNION scan fill
// (16.16 fixed-point edges active on [ymin, ymax), spans [ceil(xl), floor(xr)]).
// The Bresenham walk is evaluated in closed form per row: with major extent D, minor extent d,
// the minor coordinate after j major steps is  floor((2*d*j + D - 1) / (2*D)).
// Bit i of the mask = pixel (bx0 + i, y).  `words` must hold nwords zero-initialised words.
// ---------------------------------------------------------------------------------------
VKB_HD void set_bits(uint32_t* words, int nwords, int lo, int hi) {
    // set bits [lo, hi] (inclusive), clipped to [0, 32*nwords)
    if (lo < 0) lo = 0;
    const int top = nwords * 32 - 1;
    if (hi > top) hi = top;
    if (lo > hi) return;
    const int w0 = lo >> 5, w1 = hi >> 5;
    for (int w = w0; w <= w1; ++w) {
        const int a = (w == w0) ? (lo & 31) : 0;
        const int b = (w == w1) ? (hi & 31) : 31;
        const uint32_t m = (b == 31 ? 0xFFFFFFFFu : ((1u << (b + 1)) - 1u)) & ~((1u << a) - 1u);
        words[w] |= m;
    }
}

VKB_HD void set_bits1(uint32_t& word, int lo, int hi) {  // bits [lo, hi] of ONE word
    lo = lo < 0 ? 0 : lo;
    hi = hi > 31 ? 31 : hi;
    if (lo > hi) return;
    word |= (0xFFFFFFFFu >> (31 - hi)) & (0xFFFFFFFFu << lo);
}

VKB_HD int ceil_div_pos(int num, int den) {  // den > 0, any num; ceil(num/den)
    return (num >= 0) ? (num + den - 1) / den : -((-num) / den);
}

// floor(num / den) for 0 <= num < 2^23, 0 < den, QUOTIENT < 2^15, through an (approximate)
// float reciprocal and one correction step: the estimate is off by at most one, so the result
// is exact whatever the reciprocal's last bits are (an integer division costs ~5x more issue
// slots on the GPU).
VKB_HD int floor_div_small(int num, int den, float rcp) {
    int q = (int)((float)num * rcp);
    const int r = num - q * den;
    q += (r >= den) ? 1 : 0;
    q -= (r < 0) ? 1 : 0;
    return q;
}

// pixels of the Bresenham line (ax,ay)-(bx,by) that lie on row y: [xlo, xhi] or empty.
// cv::LineIterator, 8-connected, left to right: after j major steps the minor coordinate is
// floor((2*minor*j + major - 1) / (2*major)).
VKB_HD bool line_row_run(int ax, int ay, int bx, int by, int y, int& xlo, int& xhi) {
    if (ax > bx) {  // left to right
        int t = ax; ax = bx; bx = t;
        t = ay; ay = by; by = t;
    }
    const int dx = bx - ax;
    const int dys = by - ay;
    const int sy = dys >= 0 ? 1 : -1;
    const int dy = dys >= 0 ? dys : -dys;
    const int k = (y - ay) * sy;  // rows travelled from the start point
    if (k < 0 || k > dy) return false;
    if (dy == 0) {  // horizontal (or a single point)
        xlo = ax;
        xhi = bx;
        return true;
    }
    const int den = 2 * dy;
    const bool small = dx < 2048 && dy < 2048;  // numerators below 2^23
    const float rcp = VKB_FRCP_APPROX((float)den);
    if (dy > dx) {  // steep: one pixel per row
        const int num = 2 * dx * k + dy - 1;
        const int kx = small ? floor_div_small(num, den, rcp) : num / den;
        xlo = xhi = ax + kx;
        return true;
    }
    // shallow: ceil(n / den) = floor((n + den - 1) / den) for n >= 0
    int jlo = 0;
    if (k > 0) {
        const int n = 2 * dx * k - dx + 1;  // > 0 because dx >= dy >= 1 and k >= 1
        jlo = small ? floor_div_small(n + den - 1, den, rcp) : ceil_div_pos(n, den);
    }
    const int n2 = 2 * dx * (k + 1) - dx + 1;
    int jhi = (small ? floor_div_small(n2 + den - 1, den, rcp) : ceil_div_pos(n2, den)) - 1;
    if (jhi > dx) jhi = dx;
    if (jlo > jhi) return false;
    xlo = ax + jlo;
    xhi = ax + jhi;
    return true;
}

// 16.16 fixed-point x of the scan edge (x0,y0)-(x1,y1) at row y (cv::FillEdgeCollection):
// x_at_ymin << 16 plus trunc(((x1 - x0) << 16) / (y1 - y0)) per row.
VKB_HD long long scan_edge_x(int x0, int y0, int x1, int y1, int y) {
    const int ya = y0 < y1 ? y0 : y1;
    const int xs = y0 < y1 ? x0 : x1;
    const int ddx = x1 - x0, ddy = y1 - y0;
    const int adx = ddx < 0 ? -ddx : ddx, ady = ddy < 0 ? -ddy : ddy;
    if (adx < 128 && ady < 128) {
        // trunc((adx << 16) / ady) in two 8-bit steps so every partial quotient stays < 2^15
        const float rcp = VKB_FRCP_APPROX((float)ady);
        const int q1 = floor_div_small(adx * 256, ady, rcp);
        const int r1 = adx * 256 - q1 * ady;
        const int q = q1 * 256 + floor_div_small(r1 * 256, ady, rcp);
        const int dxf = ((ddx < 0) != (ddy < 0)) ? -q : q;
        return (long long)xs * 65536 + (long long)(dxf * (y - ya));
    }
    if (adx < 16384) {
        const int dxf = (ddx * 65536) / ddy;  // 32-bit; |dxf * (y - ya)| <= |ddx| << 16
        return (long long)xs * 65536 + (long long)(dxf * (y - ya));
    }
    const long long dxf = ((long long)ddx * 65536) / (long long)ddy;
    return (long long)xs * 65536 + dxf * (long long)(y - ya);
}

// ---------------------------------------------------------------------------------------
// The same two edge walks split into a per-edge setup and a per-row evaluation, so that code
// which visits many rows of one edge (grid_masks_kernel: lane = row) pays the divisions and the
// endpoint bookkeeping once per edge.  edge_row_run == line_row_run and edge_row_cross ==
// scan_edge_x for every input (tests/test_hostsim.py exercises both through poly_row_mask).
// ---------------------------------------------------------------------------------------
struct alignas(16) EdgeConst {
    int ax, ay, dx, dy;   // left end point, |dx|, |dy| of the outline (left to right)
    int sy;               // row direction of the outline walk (+1 / -1)
    int small;            // both deltas below 2048: divisions go through floor_div_small
    float rcp;            // ~ 1 / (2 * dy)
    int scan;             // the edge takes part in the scan fill (not horizontal)
    int ya, yb;           // scan edge active on rows [ya, yb)
    int pad0, pad1;       // pad0: the slope arithmetic fits 32 bits
    long long base, dxf;  // 16.16 x at row ya, increment per row
};

VKB_HD void edge_setup(int x0, int y0, int x1, int y1, EdgeConst& E) {
    // scan part (direction independent): start at the upper end point; the slope is
    // trunc(((x1 - x0) << 16) / (y1 - y0)), through the cheapest exact route (see scan_edge_x)
    E.scan = y0 != y1;
    E.ya = y0 < y1 ? y0 : y1;
    E.yb = y0 < y1 ? y1 : y0;
    const int xs = y0 < y1 ? x0 : x1;
    E.base = (long long)xs * 65536;
    {
        const int ddx = x1 - x0, ddy = y1 - y0;
        const int adx = ddx < 0 ? -ddx : ddx, ady = ddy < 0 ? -ddy : ddy;
        E.pad0 = adx < 16384;  // slope and slope * rows fit 32 bits
        if (!E.scan) {
            E.dxf = 0;
        } else if (adx < 128 && ady < 128) {
            const float rcp = VKB_FRCP_APPROX((float)ady);
            const int q1 = floor_div_small(adx * 256, ady, rcp);
            const int r1 = adx * 256 - q1 * ady;
            const int q = q1 * 256 + floor_div_small(r1 * 256, ady, rcp);
            E.dxf = ((ddx < 0) != (ddy < 0)) ? -q : q;
        } else if (adx < 16384) {
            E.dxf = (ddx * 65536) / ddy;
        } else {
            E.dxf = ((long long)ddx * 65536) / (long long)ddy;
        }
    }
    // outline part: cv::LineIterator walks from the left end point
    if (x0 > x1) {
        int t = x0; x0 = x1; x1 = t;
        t = y0; y0 = y1; y1 = t;
    }
    E.ax = x0;
    E.ay = y0;
    E.dx = x1 - x0;
    const int dys = y1 - y0;
    E.sy = dys >= 0 ? 1 : -1;
    E.dy = dys >= 0 ? dys : -dys;
    E.small = E.dx < 2048 && E.dy < 2048;
    E.rcp = E.dy ? VKB_FRCP_APPROX((float)(2 * E.dy)) : 0.f;
    E.pad1 = 0;
}

VKB_HD bool edge_row_run(const EdgeConst& E, int y, int& xlo, int& xhi) {
    const int dx = E.dx, dy = E.dy;
    const int k = (y - E.ay) * E.sy;  // rows travelled from the start point
    if (k < 0 || k > dy) return false;
    if (dy == 0) {  // horizontal (or a single point)
        xlo = E.ax;
        xhi = E.ax + dx;
        return true;
    }
    const int den = 2 * dy;
    if (dy > dx) {  // steep: one pixel per row
        const int num = 2 * dx * k + dy - 1;
        const int kx = E.small ? floor_div_small(num, den, E.rcp) : num / den;
        xlo = xhi = E.ax + kx;
        return true;
    }
    int jlo = 0;
    if (k > 0) {
        const int n = 2 * dx * k - dx + 1;
        jlo = E.small ? floor_div_small(n + den - 1, den, E.rcp) : ceil_div_pos(n, den);
    }
    const int n2 = 2 * dx * (k + 1) - dx + 1;
    int jhi = (E.small ? floor_div_small(n2 + den - 1, den, E.rcp) : ceil_div_pos(n2, den)) - 1;
    if (jhi > dx) jhi = dx;
    if (jlo > jhi) return false;
    xlo = E.ax + jlo;
    xhi = E.ax + jhi;
    return true;
}

VKB_HD bool edge_row_cross(const EdgeConst& E, int y, long long& x) {
    if (!E.scan || y < E.ya || y >= E.yb) return false;
    if (E.pad0) x = E.base + (long long)((int)E.dxf * (y - E.ya));
    else x = E.base + E.dxf * (long long)(y - E.ya);
    return true;
}

// coverage of row y from prepared edges; N edges, bits relative to column bx0
template <int N>
VKB_HD void poly_row_mask_edges(const EdgeConst* E, int y, int bx0, uint32_t* words, int nwords) {
    long long cross[N];
    int ncross = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        int lo, hi;
        if (edge_row_run(E[i], y, lo, hi)) {
            if (nwords == 1) set_bits1(words[0], lo - bx0, hi - bx0);
            else set_bits(words, nwords, lo - bx0, hi - bx0);
        }
        long long cx;
        if (edge_row_cross(E[i], y, cx)) cross[ncross++] = cx;
    }
    // insertion sort (N is 4 for lattice cells)
    for (int i = 1; i < ncross; ++i) {
        const long long v = cross[i];
        int j = i - 1;
        while (j >= 0 && cross[j] > v) { cross[j + 1] = cross[j]; --j; }
        cross[j + 1] = v;
    }
    for (int k = 0; k + 1 < ncross; k += 2) {
        const long long xl = (cross[k] + 65535) >> 16;
        const long long xr = cross[k + 1] >> 16;
        if (xl <= xr) {
            const long long lo = xl - bx0, hi = xr - bx0;
            const int lo_i = lo < -1 ? -1 : (lo > 1 << 20 ? 1 << 20 : (int)lo);
            const int hi_i = hi < -1 ? -1 : (hi > 1 << 20 ? 1 << 20 : (int)hi);
            if (nwords == 1) set_bits1(words[0], lo_i, hi_i);
            else set_bits(words, nwords, lo_i, hi_i);
        }
    }
}

// ---------------------------------------------------------------------------------------
// Branch-free row coverage of a QUAD whose edges are all "small" (deltas below 2048, slope
// arithmetic in 32 bits, coordinates in [0, 32768)) -- every lattice cell.  Same results as
// poly_row_mask_edges<4> (cross-checked on the host); no early returns, no divergent loops:
// invalid runs / crossings are carried as empty masks / sentinels, the four crossings go through
// a 5-exchange sorting network on 32-bit keys.
// ---------------------------------------------------------------------------------------
VKB_HD bool edges_fast_ok(const EdgeConst* E) {
    return E[0].small && E[1].small && E[2].small && E[3].small && E[0].pad0 && E[1].pad0
           && E[2].pad0 && E[3].pad0;
}

VKB_HD uint32_t bits_lo_hi(int lo, int hi) {  // bits [lo, hi] clipped to one word; empty if lo > hi
    const int l = lo < 0 ? 0 : lo, h = hi > 31 ? 31 : hi;
    const uint32_t m = (0xFFFFFFFFu >> (31 - (h & 31))) & (0xFFFFFFFFu << (l & 31));
    return (l <= h) ? m : 0u;
}

VKB_HD uint32_t edge_row_bits_fast(const EdgeConst& E, int y, int bx0) {
    const int dx = E.dx, dy = E.dy;
    const int k = (y - E.ay) * E.sy;
    bool valid = (unsigned)k <= (unsigned)dy;
    const int kk = k < 0 ? 0 : (k > dy ? dy : k);
    int lo, hi;
    if (dy == 0) {  // uniform per edge
        lo = E.ax;
        hi = E.ax + dx;
    } else {
        const int den = 2 * dy;
        if (dy > dx) {
            lo = hi = E.ax + floor_div_small(2 * dx * kk + dy - 1, den, E.rcp);
        } else {
            const int n = 2 * dx * kk - dx + den;  // (2*dx*k - dx + 1) + den - 1
            const int jl = floor_div_small(n < 0 ? 0 : n, den, E.rcp);
            const int jlo = kk > 0 ? jl : 0;
            int jhi = floor_div_small(2 * dx * (kk + 1) - dx + den, den, E.rcp) - 1;
            jhi = jhi > dx ? dx : jhi;
            valid = valid && jlo <= jhi;
            lo = E.ax + jlo;
            hi = E.ax + jhi;
        }
    }
    const uint32_t m = bits_lo_hi(lo - bx0, hi - bx0);
    return valid ? m : 0u;
}

constexpr int kNoCross = 0x7fffffff;

VKB_HD int edge_row_cross_fast(const EdgeConst& E, int y) {
    const bool valid = E.scan && y >= E.ya && y < E.yb;
    const int x = (int)E.base + (int)E.dxf * (y - E.ya);
    return valid ? x : kNoCross;
}

VKB_HD uint32_t quad_row_mask_fast(const EdgeConst* E, int y, int bx0) {
    uint32_t word = edge_row_bits_fast(E[0], y, bx0) | edge_row_bits_fast(E[1], y, bx0)
                    | edge_row_bits_fast(E[2], y, bx0) | edge_row_bits_fast(E[3], y, bx0);
    int c0 = edge_row_cross_fast(E[0], y), c1 = edge_row_cross_fast(E[1], y);
    int c2 = edge_row_cross_fast(E[2], y), c3 = edge_row_cross_fast(E[3], y);
#define VKB_CSWAP(a, b) { const int lo_ = a < b ? a : b, hi_ = a < b ? b : a; a = lo_; b = hi_; }
    VKB_CSWAP(c0, c1) VKB_CSWAP(c2, c3) VKB_CSWAP(c0, c2) VKB_CSWAP(c1, c3) VKB_CSWAP(c1, c2)
#undef VKB_CSWAP
    // crossings come in pairs; sentinels sort last
    {
        const int xl = (int)(((long long)c0 + 65535) >> 16), xr = c1 >> 16;
        const uint32_t m = bits_lo_hi(xl - bx0, xr - bx0);
        word |= (c1 != kNoCross && xl <= xr) ? m : 0u;
    }
    {
        const int xl = (int)(((long long)c2 + 65535) >> 16), xr = c3 >> 16;
        const uint32_t m = bits_lo_hi(xl - bx0, xr - bx0);
        word |= (c3 != kNoCross && xl <= xr) ? m : 0u;
    }
    return word;
}

// ---------------------------------------------------------------------------------------
// The same coverage split the other way round (grid_masks_kernel, second generation): the outline
// is WALKED once per edge (cv::LineIterator step by step, a few integer operations per pixel)
// instead of being solved per row in closed form, and only the scan fill is evaluated per row.
//
//   edge_walk    visits the pixels of the 8-connected line from the LEFT end point: after j major
//                steps the minor coordinate is floor((2*minor*j + major - 1) / (2*major)), kept
//                incrementally as (quotient, remainder) -- one add, one compare per step;
//   EdgeScan     the 16.16 scan edge of cv::FillEdgeCollection for rows [ya, yb);
//   quad_fill_row  the spans of one row from the four scan edges (crossings sorted by a
//                5-exchange network, pairs filled from ceil(xl) to floor(xr)).
// coverage(row) = outline bits | quad_fill_row: same bits as poly_row_mask<4> (checked on the host
// against cv.fillPoly's model, tests/test_hostsim.py).
// ---------------------------------------------------------------------------------------
template <typename Plot>
VKB_HD void edge_walk(int x0, int y0, int x1, int y1, Plot plot) {
    if (x0 > x1) {  // left to right
        int t = x0; x0 = x1; x1 = t;
        t = y0; y0 = y1; y1 = t;
    }
    const int dx = x1 - x0;
    const int dys = y1 - y0;
    const int sy = dys >= 0 ? 1 : -1;
    const int dy = dys >= 0 ? dys : -dys;
    const bool steep = dy > dx;
    const int major = steep ? dy : dx, minor_ext = steep ? dx : dy;
    int rem = major - 1, minor = 0;  // (2*minor_ext*j + major - 1) = 2*major*minor + rem
    for (int j = 0; j <= major; ++j) {
        const int x = steep ? x0 + minor : x0 + j;
        const int y = steep ? y0 + sy * j : y0 + sy * minor;
        plot(x, y);
        rem += 2 * minor_ext;
        if (rem >= 2 * major) {  // minor_ext <= major: at most one step
            rem -= 2 * major;
            ++minor;
        }
    }
}

struct alignas(16) EdgeScan {
    int ya, yb;  // active rows [ya, yb); ya == yb: horizontal, takes no part in the fill
    int base;    // x << 16 at row ya
    int dxf;     // trunc(((x1 - x0) << 16) / (y1 - y0)) per row
};

// |x1 - x0| < 16384 and 0 <= x < 32768 (every lattice cell the masks kernel accepts)
VKB_HD void edge_scan_setup(int x0, int y0, int x1, int y1, EdgeScan& e) {
    e.ya = y0 < y1 ? y0 : y1;
    e.yb = y0 < y1 ? y1 : y0;
    e.base = (y0 < y1 ? x0 : x1) * 65536;
    e.dxf = y0 != y1 ? ((x1 - x0) * 65536) / (y1 - y0) : 0;
}

VKB_HD uint32_t quad_fill_row(const EdgeScan* e, int y, int bx0) {
    int c[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
        c[i] = (y >= e[i].ya && y < e[i].yb) ? e[i].base + e[i].dxf * (y - e[i].ya) : kNoCross;
#define VKB_CSWAP(a, b) { const int lo_ = a < b ? a : b, hi_ = a < b ? b : a; a = lo_; b = hi_; }
    VKB_CSWAP(c[0], c[1]) VKB_CSWAP(c[2], c[3]) VKB_CSWAP(c[0], c[2]) VKB_CSWAP(c[1], c[3]) VKB_CSWAP(c[1], c[2])
#undef VKB_CSWAP
    uint32_t word = 0;
    {
        const int xl = (int)(((long long)c[0] + 65535) >> 16), xr = c[1] >> 16;
        const uint32_t m = bits_lo_hi(xl - bx0, xr - bx0);
        word |= (c[1] != kNoCross && xl <= xr) ? m : 0u;
    }
    {
        const int xl = (int)(((long long)c[2] + 65535) >> 16), xr = c[3] >> 16;
        const uint32_t m = bits_lo_hi(xl - bx0, xr - bx0);
        word |= (c[3] != kNoCross && xl <= xr) ? m : 0u;
    }
    return word;
}

// Direct form (no per-edge state): used where a single row of a polygon is needed once
// (over-budget cells in the remap's slow path); same results as poly_row_mask_edges.
template <int N>
VKB_HD void poly_row_mask(const int* px, const int* py, int y, int bx0, uint32_t* words,
                          int nwords) {
    long long cross[N];
    int ncross = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const int j = (i + N - 1) % N;
        const int x0 = px[j], y0 = py[j], x1 = px[i], y1 = py[i];
        int lo, hi;
        if (line_row_run(x0, y0, x1, y1, y, lo, hi)) {
            if (nwords == 1) set_bits1(words[0], lo - bx0, hi - bx0);
            else set_bits(words, nwords, lo - bx0, hi - bx0);
        }
        if (y0 == y1) continue;
        const int ya = y0 < y1 ? y0 : y1, yb = y0 < y1 ? y1 : y0;
        if (ya <= y && y < yb) cross[ncross++] = scan_edge_x(x0, y0, x1, y1, y);
    }
    // insertion sort (N is 4 for lattice cells)
    for (int i = 1; i < ncross; ++i) {
        const long long v = cross[i];
        int j = i - 1;
        while (j >= 0 && cross[j] > v) { cross[j + 1] = cross[j]; --j; }
        cross[j + 1] = v;
    }
    for (int k = 0; k + 1 < ncross; k += 2) {
        const long long xl = (cross[k] + 65535) >> 16;
        const long long xr = cross[k + 1] >> 16;
        if (xl <= xr) {
            const long long lo = xl - bx0, hi = xr - bx0;
            const int lo_i = lo < -1 ? -1 : (lo > 1 << 20 ? 1 << 20 : (int)lo);
            const int hi_i = hi < -1 ? -1 : (hi > 1 << 20 ? 1 << 20 : (int)hi);
            if (nwords == 1) set_bits1(words[0], lo_i, hi_i);
            else set_bits(words, nwords, lo_i, hi_i);
        }
    }
}

// ---------------------------------------------------------------------------------------
// Camera lattice projection (cv.projectPoints with zero distortion, camera.py:188-196).
// R row major, t, f: all double.  Returns image coordinates in double; caller casts to the
// strategy's dtype (float32 for plane/line strategies, float64 for cubic curve).
// ---------------------------------------------------------------------------------------
VKB_HD void project_point(const double* R, const double* t, double f, double X, double Y,
                          double Z, double& u, double& v) {
    double x = VKB_DADD(VKB_DADD(VKB_DADD(VKB_DMUL(R[0], X), VKB_DMUL(R[1], Y)), VKB_DMUL(R[2], Z)), t[0]);
    double y = VKB_DADD(VKB_DADD(VKB_DADD(VKB_DMUL(R[3], X), VKB_DMUL(R[4], Y)), VKB_DMUL(R[5], Z)), t[1]);
    double z = VKB_DADD(VKB_DADD(VKB_DADD(VKB_DMUL(R[6], X), VKB_DMUL(R[7], Y)), VKB_DMUL(R[8], Z)), t[2]);
    z = (z != 0.0) ? 1.0 / z : 1.0;
    x = VKB_DMUL(x, z);
    y = VKB_DMUL(y, z);
    u = VKB_DADD(VKB_DMUL(x, f), 0.0);
    v = VKB_DADD(VKB_DMUL(y, f), 0.0);
}

}  // namespace vkb
