// hostsim.cpp -- TEST HARNESS ONLY.  Compiles vkit_b200/csrc/vkb_math.cuh and vkb_lattice.cuh
// (the per-element numerics the CUDA kernels call) with g++ and walks them sequentially over a
// page, so the arithmetic can be compared with the oracle on a machine that has no GPU.
// Nothing in vkit_b200/ links or loads this file.
#include <algorithm>
#include <cstring>
#include <vector>
#include "../../vkit_b200/csrc/vkb_math.cuh"
#include "../../vkit_b200/csrc/vkb_lattice.cuh"
#include "../../vkit_b200/csrc/vkb_draw_host.h"

using namespace vkb;

static void sample_and_store(const vkb_planes& pl, int x, int y, int X, int Y) {
    const long long di = (long long)y * pl.dst_w + x;
    if (pl.image_channels == 3) bilinear_u8<3>(pl.src_image, pl.src_h, pl.src_w, (long long)pl.src_w * 3, X, Y, pl.dst_image + di * 3);
    else if (pl.image_channels == 1) bilinear_u8<1>(pl.src_image, pl.src_h, pl.src_w, pl.src_w, X, Y, pl.dst_image + di);
    else if (pl.image_channels == 4) bilinear_u8<4>(pl.src_image, pl.src_h, pl.src_w, (long long)pl.src_w * 4, X, Y, pl.dst_image + di * 4);
    if (pl.src_mask) bilinear_u8<1>(pl.src_mask, pl.src_h, pl.src_w, pl.src_w, X, Y, pl.dst_mask + di);
    if (pl.src_score) pl.dst_score[di] = bilinear_f32(pl.src_score, pl.src_h, pl.src_w, pl.src_w, X, Y);
}

extern "C" void hs_warp(const vkb_warp_page* pg) {
    const vkb_planes& pl = pg->planes;
    for (int y = 0; y < pl.dst_h; ++y)
        for (int x = 0; x < pl.dst_w; ++x) {
            int X, Y;
            if (pg->kind == VKB_WARP_AFFINE) affine_coord(pg->inv, x, y, X, Y);
            else perspective_coord(pg->inv, x, y, X, Y);
            sample_and_store(pl, x, y, X, Y);
        }
}

// lattice_f: P x 2 doubles (x, y)
extern "C" void hs_grid_project(const vkb_grid_page* pg, double* out) {
    const int P = pg->rows * pg->cols;
    auto cx = [&](int i) { return lattice_coord(i % pg->cols, pg->src_w, pg->grid_size); };
    auto cy = [&](int i) { return lattice_coord(i / pg->cols, pg->src_h, pg->grid_size); };
    if (pg->projector == VKB_PROJ_MLS) {
        for (int i = 0; i < P; ++i)
            mls_point_seq(pg->handles_src, pg->handles_dst, pg->n_handles, (float)cx(i), (float)cy(i), out[2 * i], out[2 * i + 1]);
        return;
    }
    if (pg->projector != VKB_PROJ_CAMERA) return;
    if (pg->strategy == VKB_CAM_PLANE) {
        for (int i = 0; i < P; ++i) {
            double u, v;
            project_point(pg->R, pg->t, pg->focal, cx(i), cy(i), 0.0, u, v);
            out[2 * i] = (double)(float)u;
            out[2 * i + 1] = (double)(float)v;
        }
    } else if (pg->strategy == VKB_CAM_CUBIC) {
        std::vector<double> z(P);
        double sum = 0;
        for (int i = 0; i < P; ++i) { z[i] = cubic_z(*pg, (float)cx(i), (float)cy(i)); sum += z[i]; }
        const double mean = sum / P;
        for (int i = 0; i < P; ++i)
            project_point(pg->R, pg->t, pg->focal, cx(i), cy(i), z[i] - mean, out[2 * i], out[2 * i + 1]);
    } else {
        std::vector<double> w(P);
        double s[3] = {0, 0, 0};
        for (int i = 0; i < P; ++i) {
            w[i] = line_weight(*pg, (float)cx(i), (float)cy(i));
            for (int k = 0; k < 3; ++k) s[k] += w[i] * (double)pg->perturb[k];
        }
        for (int k = 0; k < 3; ++k) s[k] /= P;
        for (int i = 0; i < P; ++i) {
            const float X = (float)((double)cx(i) + (w[i] * (double)pg->perturb[0] - s[0]));
            const float Y = (float)((double)cy(i) + (w[i] * (double)pg->perturb[1] - s[1]));
            const float Z = (float)(0.0 + (w[i] * (double)pg->perturb[2] - s[2]));
            double u, v;
            project_point(pg->R, pg->t, pg->focal, X, Y, Z, u, v);
            out[2 * i] = (double)(float)u;
            out[2 * i + 1] = (double)(float)v;
        }
    }
}

extern "C" void hs_grid_finalize(const vkb_grid_page* pg, const double* in, int32_t* out, vkb_grid_meta* m) {
    const int P = pg->rows * pg->cols;
    int mnx = INT32_MAX, mny = INT32_MAX;
    for (int i = 0; i < P; ++i) { mnx = std::min(mnx, cv_round_d(in[2 * i])); mny = std::min(mny, cv_round_d(in[2 * i + 1])); }
    int mxx = INT32_MIN, mxy = INT32_MIN;
    for (int i = 0; i < P; ++i) {
        out[2 * i] = cv_round_d(in[2 * i] - (double)mnx);
        out[2 * i + 1] = cv_round_d(in[2 * i + 1] - (double)mny);
        mxx = std::max(mxx, out[2 * i]); mxy = std::max(mxy, out[2 * i + 1]);
    }
    m->dst_w = mxx + 1; m->dst_h = mxy + 1; m->shift_x = mnx; m->shift_y = mny;
    m->resize_ratio_x = m->resize_ratio_y = 1.0; m->status = 0; m->n_flagged_cells = 0;
    if (pg->resize_as_src) {
        const int raw_w = m->dst_w, raw_h = m->dst_h;
        m->resize_ratio_y = (double)pg->src_h / raw_h; m->resize_ratio_x = (double)pg->src_w / raw_w;
        for (int i = 0; i < P; ++i) {
            double sx = (in[2 * i] - (double)mnx) * (double)pg->src_w / (double)raw_w;
            double sy = (in[2 * i + 1] - (double)mny) * (double)pg->src_h / (double)raw_h;
            sx = std::max(0.0, std::min(sx, (double)(pg->src_w - 1)));
            sy = std::max(0.0, std::min(sy, (double)(pg->src_h - 1)));
            out[2 * i] = cv_round_d(sx); out[2 * i + 1] = cv_round_d(sy);
        }
        m->dst_h = pg->src_h; m->dst_w = pg->src_w;
    }
}

// owner map (int32, -1 = uncovered) + inverse homographies + remap of the planes
extern "C" void hs_grid_remap(const vkb_grid_page* pg, const int32_t* lat, const vkb_planes* pl,
                              int32_t* owner, double* hinv_out) {
    const int ccols = pg->cols - 1, C = (pg->rows - 1) * ccols;
    const int W = pl->dst_w, Hh = pl->dst_h;
    std::fill(owner, owner + (size_t)W * Hh, -1);
    std::vector<double> hinv((size_t)C * 9);
    for (int cell = 0; cell < C; ++cell) {
        const int r = cell / ccols, c = cell % ccols;
        const int i00 = r * pg->cols + c, i01 = i00 + 1, i11 = i00 + pg->cols + 1, i10 = i00 + pg->cols;
        const int px[4] = {lat[2 * i00], lat[2 * i01], lat[2 * i11], lat[2 * i10]};
        const int py[4] = {lat[2 * i00 + 1], lat[2 * i01 + 1], lat[2 * i11 + 1], lat[2 * i10 + 1]};
        const double sx0 = lattice_coord(c, pg->src_w, pg->grid_size), sx1 = lattice_coord(c + 1, pg->src_w, pg->grid_size);
        const double sy0 = lattice_coord(r, pg->src_h, pg->grid_size), sy1 = lattice_coord(r + 1, pg->src_h, pg->grid_size);
        const double sq[8] = {sx0, sy0, sx1, sy0, sx1, sy1, sx0, sy1};
        const double dq[8] = {(double)px[0], (double)py[0], (double)px[1], (double)py[1], (double)px[2], (double)py[2], (double)px[3], (double)py[3]};
        homography_4pt(dq, sq, &hinv[(size_t)cell * 9]);
        const int x0 = *std::min_element(px, px + 4), x1 = *std::max_element(px, px + 4);
        const int y0 = *std::min_element(py, py + 4), y1 = *std::max_element(py, py + 4);
        const int nwords = (x1 - x0 + 32) / 32;
        std::vector<uint32_t> words(nwords);
        for (int y = y0; y <= y1; ++y) {
            std::fill(words.begin(), words.end(), 0u);
            poly_row_mask<4>(px, py, y, x0, words.data(), nwords);
            for (int x = x0; x <= x1; ++x)
                if ((words[(x - x0) >> 5] >> ((x - x0) & 31)) & 1u)
                    if (y >= 0 && y < Hh && x >= 0 && x < W) owner[(size_t)y * W + x] = std::max(owner[(size_t)y * W + x], cell);
        }
    }
    if (hinv_out) std::memcpy(hinv_out, hinv.data(), sizeof(double) * hinv.size());
    for (int y = 0; y < Hh; ++y)
        for (int x = 0; x < W; ++x) {
            int X = 0, Y = 0;
            const int o = owner[(size_t)y * W + x];
            if (o >= 0) cell_coord(&hinv[(size_t)o * 9], x, y, X, Y);
            sample_and_store(*pl, x, y, X, Y);
        }
}

extern "C" void hs_fill_poly4(const int32_t* pts, int h, int w, uint8_t* out) {
    const int px[4] = {pts[0], pts[2], pts[4], pts[6]}, py[4] = {pts[1], pts[3], pts[5], pts[7]};
    const int nwords = (w + 31) / 32;
    std::vector<uint32_t> words(nwords), words2(nwords);
    // both forms of the row coverage: the direct one and the per-edge setup + row evaluation the
    // masks kernel uses; a pixel on which they disagree is reported as 7
    EdgeConst E[4];
    for (int i = 0; i < 4; ++i) edge_setup(px[(i + 3) & 3], py[(i + 3) & 3], px[i], py[i], E[i]);
    // third form (grid_masks_kernel): outline walked per edge + per-row scan fill, for polygons
    // that fit one 32-bit window per row
    std::vector<uint32_t> walked(h, 0u);
    EdgeScan ES[4];
    const bool one_word = w <= 32;
    if (one_word) {
        for (int i = 0; i < 4; ++i) {
            const int ax = px[(i + 3) & 3], ay = py[(i + 3) & 3], bx = px[i], by = py[i];
            edge_scan_setup(ax, ay, bx, by, ES[i]);
            edge_walk(ax, ay, bx, by, [&](int x, int y) {
                if (y >= 0 && y < h && x >= 0 && x < 32) walked[y] |= 1u << x;
            });
        }
    }
    for (int y = 0; y < h; ++y) {
        std::fill(words.begin(), words.end(), 0u);
        std::fill(words2.begin(), words2.end(), 0u);
        poly_row_mask<4>(px, py, y, 0, words.data(), nwords);
        poly_row_mask_edges<4>(E, y, 0, words2.data(), nwords);
        for (int x = 0; x < w; ++x) {
            const uint32_t a = (words[x >> 5] >> (x & 31)) & 1u, b = (words2[x >> 5] >> (x & 31)) & 1u;
            out[(size_t)y * w + x] = a == b ? a : 7;
        }
        // the branch-free form works on one 32-bit window at a time
        if (edges_fast_ok(E)) {
            for (int wd = 0; wd < nwords; ++wd)
                if (quad_row_mask_fast(E, y, 32 * wd) != words[wd])
                    for (int x = 32 * wd; x < w && x < 32 * wd + 32; ++x) out[(size_t)y * w + x] = 9;
        }
        if (one_word && (walked[y] | quad_fill_row(ES, y, 0)) != words[0])
            for (int x = 0; x < w; ++x) out[(size_t)y * w + x] = 11;
    }
}

extern "C" void hs_homography(const double* src_quad, const double* dst_quad, double* H) { homography_4pt(src_quad, dst_quad, H); }
extern "C" int hs_sizeof_grid_page() { return (int)sizeof(vkb_grid_page); }
extern "C" int hs_sizeof_warp_page() { return (int)sizeof(vkb_warp_page); }
extern "C" int hs_sizeof_planes() { return (int)sizeof(vkb_planes); }
extern "C" int hs_sizeof_grid_meta() { return (int)sizeof(vkb_grid_meta); }

// Fast-path audit: for every covered pixel compare the float32 fast path with the float64 path.
// stats[0] pixels, [1] fast ok, [2] ok but X/Y differ from exact (must be 0),
// [3] max |32*du_fast - 32*du_exact| * 1e9 (x or y), over pixels of covered cells.
extern "C" void hs_fast_path_stats(const vkb_grid_page* pg, const int32_t* lat, int W, int Hh,
                                   const int32_t* owner, long long* stats) {
    const int ccols = pg->cols - 1, C = (pg->rows - 1) * ccols;
    std::vector<double> hinv((size_t)C * 9);
    std::vector<int> sx(C), sy(C), bx0(C), by0(C), big(C);
    for (int cell = 0; cell < C; ++cell) {
        const int r = cell / ccols, c = cell % ccols;
        const int i00 = r * pg->cols + c, i01 = i00 + 1, i11 = i00 + pg->cols + 1, i10 = i00 + pg->cols;
        const int px[4] = {lat[2 * i00], lat[2 * i01], lat[2 * i11], lat[2 * i10]};
        const int py[4] = {lat[2 * i00 + 1], lat[2 * i01 + 1], lat[2 * i11 + 1], lat[2 * i10 + 1]};
        const int sx0 = lattice_coord(c, pg->src_w, pg->grid_size), sx1 = lattice_coord(c + 1, pg->src_w, pg->grid_size);
        const int sy0 = lattice_coord(r, pg->src_h, pg->grid_size), sy1 = lattice_coord(r + 1, pg->src_h, pg->grid_size);
        const double sq[8] = {(double)sx0, (double)sy0, (double)sx1, (double)sy0, (double)sx1, (double)sy1, (double)sx0, (double)sy1};
        const double dq[8] = {(double)px[0], (double)py[0], (double)px[1], (double)py[1], (double)px[2], (double)py[2], (double)px[3], (double)py[3]};
        homography_4pt(dq, sq, &hinv[(size_t)cell * 9]);
        sx[cell] = sx0;
        sy[cell] = sy0;
        // bbox of the dst cell: the origin its float32 form is re-centred on (grid_cells_kernel)
        int x0 = px[0], x1 = px[0], y0 = py[0], y1 = py[0];
        for (int k = 1; k < 4; ++k) {
            x0 = px[k] < x0 ? px[k] : x0; x1 = px[k] > x1 ? px[k] : x1;
            y0 = py[k] < y0 ? py[k] : y0; y1 = py[k] > y1 ? py[k] : y1;
        }
        bx0[cell] = x0;
        by0[cell] = y0;
        big[cell] = (x1 - x0 + 32) / 32 != 1 || y1 - y0 + 1 > VKB_CELL_MASK_WORDS;  // float64 path only
    }
    const int extent = pg->src_w > pg->src_h ? pg->src_w : pg->src_h;
    stats[0] = stats[1] = stats[2] = stats[3] = 0;
    for (int y = 0; y < Hh; ++y)
        for (int x = 0; x < W; ++x) {
            const int o = owner[(size_t)y * W + x];
            if (o < 0) continue;
            stats[0]++;
            // the kernel's form: re-centred on the origin of the owner's bbox
            if (big[o]) continue;
            const int ox = bx0[o], oy = by0[o];
            CellLocal L;
            make_cell_local(&hinv[(size_t)o * 9], sx[o], sy[o], ox, oy, L);
            int Xe, Ye, Xf, Yf;
            cell_coord(&hinv[(size_t)o * 9], x, y, Xe, Ye);
            const float xr = (float)(x - ox), yr = (float)(y - oy);
            // The kernel takes the margin from the largest source corner among the cells that
            // touch the pixel's tile; the owner is one of them, so the owner's own corner gives
            // the narrowest margin (= the largest accepted set) any tile can use for this pixel.
            const int margin = fast_margin(sx[o] > sy[o] ? sx[o] : sy[o]), limit = fast_limit(margin);
            const bool ok = fast_page_ok(extent)
                            && cell_coord_fast(L, xr, yr, fast_base(sx[o], margin),
                                               fast_base(sy[o], margin), limit, Xf, Yf) >= 0;
            if (ok) { stats[1]++; if (Xf != Xe || Yf != Ye) stats[2]++; }
            // error of the float32 evaluation itself
            const double* H = &hinv[(size_t)o * 9];
            const double den = H[6] * x + H[7] * y + H[8];
            const double ux = (H[0] * x + H[1] * y + H[2]) / den, uy = (H[3] * x + H[4] * y + H[5]) / den;
            CellColumn col;
            cell_column(L, xr, col);
            const float d = fma_rn_f32(L.h, yr, col.d);
            const float nx = fma_rn_f32(L.a1, yr, col.nx);
            const float ny = fma_rn_f32(L.b1, yr, col.ny);
            const float r = 1.0f / d;
            // what the kernel's test sees: the sums rounded to whole fast-path units
            const double fx = rint((double)fma_rn_f32(nx, r, kRoundMagic) - (double)kRoundMagic) / kFastUnits;
            const double fy = rint((double)fma_rn_f32(ny, r, kRoundMagic) - (double)kRoundMagic) / kFastUnits;
            const double ex = fabs(fx - 32.0 * (ux - sx[o]));
            const double ey = fabs(fy - 32.0 * (uy - sy[o]));
            if (!(fmax(fabs((double)nx * r), fabs((double)ny * r)) < kFastRange * kFastUnits)) continue;
            const long long e = (long long)(fmax(ex, ey) * 1e9);
            if (e > stats[3]) stats[3] = e;
        }
}


// cv.ellipse(img, center, axes, 0, 0, 360, 1, thickness) through the product's drawing code
// (vkb_draw_host.h vertices + vkb_draw.cuh primitives), one segment after the other
extern "C" void hs_ellipse(int h, int w, int cx, int cy, int ax, int ay, int thickness, uint8_t* out) {
    std::vector<DrawSegment> segs;
    ellipse_segments(cx, cy, ax, ay, segs);
    auto plot = [&](int x, int y) { out[(size_t)y * w + x] = 1; };
    auto hline = [&](int y, int xa, int xb) { for (int x = xa; x <= xb; ++x) out[(size_t)y * w + x] = 1; };
    for (const DrawSegment& sg : segs) draw_thick_segment(w, h, sg.p0, sg.p1, thickness, sg.flags, plot, hline);
}
