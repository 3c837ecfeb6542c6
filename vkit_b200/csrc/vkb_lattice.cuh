// vkb_lattice.cuh -- lattice projection numerics (camera strategies, similarity MLS).
//
// `__host__ __device__` like vkb_math.cuh, for the same reason.  Each step keeps the dtype the
// reference's NumPy code has at that point (float32 vs float64, fused vs unfused), because the
// projected lattice is rounded to integers afterwards and a flipped corner moves four cells.
#pragma once
#include "vkb_math.cuh"
#include "../../include/vkit_b200.h"

namespace vkb {

// create_src_image_grid (grid_creator.py:22-41): range(0, size, g) plus size-1 if missing.
VKB_HD int lattice_range_count(int size, int g) { return (size + g - 1) / g; }
VKB_HD int lattice_point_count(int size, int g) {
    const int n = lattice_range_count(size, g);
    return n + (((n - 1) * g != size - 1) ? 1 : 0);
}
VKB_HD int lattice_coord(int idx, int size, int g) {
    const int n = lattice_range_count(size, g);
    return idx < n ? idx * g : size - 1;
}

VKB_HD float fma_f32(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
    return __fmaf_rn(a, b, c);
#else
    return (float)((double)a * (double)b + (double)c);  // exact: 24x24-bit product fits double
#endif
}

// ----- cubic curve: z before mean subtraction (camera.py:375-398) -----------------------
// plane_projected_xs = (rotation_mat @ pts.T)[0] in float32 through BLAS sgemm, which on
// FMA hardware evaluates fl(r0*x) then fma(r1, y, .) (verified against numpy here);
// np.polyval promotes to float64 from its first step because the coefficients are float64.
VKB_HD double cubic_z(const vkb_grid_page& p, float x, float y) {
    const float pp = fma_f32(p.rot2[1], y, VKB_FMUL(p.rot2[0], x));
    const float ratio = VKB_FSUB(pp, p.proj_min) / p.proj_range;
    const double r = (double)ratio;
    double v = p.poly[0];
    v = VKB_DADD(VKB_DMUL(v, r), p.poly[1]);
    v = VKB_DADD(VKB_DMUL(v, r), p.poly[2]);
    v = VKB_DADD(VKB_DMUL(v, r), p.poly[3]);
    v = VKB_DMUL(VKB_DMUL(v, (double)p.proj_range), p.curve_scale);
    return v;
}

// ----- plane line fold / curve: weight of the perturbation vector (camera.py:464-480) ----
VKB_HD double line_weight(const vkb_grid_page& p, float x, float y) {
    const float s = VKB_FADD(VKB_FMUL(x, p.line_ab[0]), VKB_FMUL(y, p.line_ab[1]));
    const double dist = fabs(VKB_DADD((double)s, p.line_c));
    const double nd = dist / p.dist_max;
    if (p.strategy == VKB_CAM_LINE_FOLD) return p.line_alpha / VKB_DADD(nd, p.line_alpha);
    return 1.0 - pow(nd, p.line_alpha);
}

// ----- similarity MLS (mls.py:52-135), one lattice point, sequential float32 -----------
// Used by the host harness and as the single-lane fallback; the kernel distributes the
// handle loop over a warp and reduces with shuffles.
VKB_HD void mls_point_seq(const float* hs, const float* hd, int n, float vx, float vy,
                          double& out_x, double& out_y) {
    float sum_inv = 0.f;
    for (int i = 0; i < n; ++i) {
        const float dx = hs[2 * i] - vx, dy = hs[2 * i + 1] - vy;
        const float d2 = VKB_FADD(VKB_FMUL(dx, dx), VKB_FMUL(dy, dy));
        if (d2 == 0.f) {  // exact handle hit: identity to the dst handle (mls.py:57-61)
            out_x = (double)hd[2 * i];
            out_y = (double)hd[2 * i + 1];
            return;
        }
        sum_inv = VKB_FADD(sum_inv, 1.0f / d2);
    }
    float pcx = 0.f, pcy = 0.f, qcx = 0.f, qcy = 0.f;
    for (int i = 0; i < n; ++i) {
        const float dx = hs[2 * i] - vx, dy = hs[2 * i + 1] - vy;
        const float inv = 1.0f / VKB_FADD(VKB_FMUL(dx, dx), VKB_FMUL(dy, dy));
        const float w = inv / sum_inv;
        pcx = fma_f32(w, hs[2 * i], pcx);
        pcy = fma_f32(w, hs[2 * i + 1], pcy);
        qcx = fma_f32(w, hd[2 * i], qcx);
        qcy = fma_f32(w, hd[2 * i + 1], qcy);
    }
    const float ax = VKB_FSUB(vx, pcx), ay = VKB_FSUB(vy, pcy);
    float mu = 0.f, sx = 0.f, sy = 0.f;
    for (int i = 0; i < n; ++i) {
        const float dx = hs[2 * i] - vx, dy = hs[2 * i + 1] - vy;
        const float inv = 1.0f / VKB_FADD(VKB_FMUL(dx, dx), VKB_FMUL(dy, dy));
        const float hx = VKB_FSUB(hs[2 * i], pcx), hy = VKB_FSUB(hs[2 * i + 1], pcy);
        const float qx = VKB_FSUB(hd[2 * i], qcx), qy = VKB_FSUB(hd[2 * i + 1], qcy);
        // rows of the 2x2 block through sgemm: fl(h0*a0) then fma(h1, a1, .)
        const float r00 = fma_f32(hy, ay, VKB_FMUL(hx, ax));
        const float r01 = fma_f32(hy, -ax, VKB_FMUL(hx, ay));
        const float r10 = fma_f32(-hx, ay, VKB_FMUL(hy, ax));
        const float r11 = fma_f32(-hx, -ax, VKB_FMUL(hy, ay));
        const float m00 = VKB_FMUL(inv, r00), m01 = VKB_FMUL(inv, r01);
        const float m10 = VKB_FMUL(inv, r10), m11 = VKB_FMUL(inv, r11);
        const float px = VKB_FADD(VKB_FMUL(qx, m00), VKB_FMUL(qy, m10));
        const float py = VKB_FADD(VKB_FMUL(qx, m01), VKB_FMUL(qy, m11));
        sx = VKB_FADD(sx, px);
        sy = VKB_FADD(sy, py);
        mu = VKB_FADD(mu, VKB_FMUL(inv, VKB_FADD(VKB_FMUL(hx, hx), VKB_FMUL(hy, hy))));
    }
    out_x = (double)VKB_FADD(sx / mu, qcx);
    out_y = (double)VKB_FADD(sy / mu, qcy);
}

}  // namespace vkb
