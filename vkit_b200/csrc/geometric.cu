// geometric.cu -- sm_100a kernels + C ABI for the geometric distortions.
//
//   warp_fused_kernel      rotate / shear / skew        (affine.py:38-43, 416-456)
//   grid_project_*         lattice projection            (camera.py, mls.py, point_projector.py)
//   grid_finalize_kernel   round, shift, result shape    (grid_creator.py:44-115)
//   grid_cells_kernel      per-cell homographies, bins   (type.py:165-197)
//   grid_masks_kernel      cv.fillPoly coverage per cell (type.py:199-207, polygon.py:70-77)
//   grid_remap_kernel      owner + map + bilinear gather (type.py:209-261, grid_blender.py:54-81)
//
// All kernels are HBM/issue bound integer + fp64 work; no tensor cores (no contraction here).
#include <stdarg.h>
#include <string.h>
#include "common.cuh"
#include "vkb_math.cuh"
#include "vkb_lattice.cuh"

namespace vkb {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

// ============================================================================================
// Shared per-pixel output stage: sample every present container at (X, Y) and store.
// ============================================================================================
__device__ __forceinline__ void sample_and_store(const vkb_planes& pl, int x, int y, int X, int Y) {
    const long long dst_idx = (long long)y * pl.dst_w + x;
    if (pl.image_channels == 3) {
        uint8_t px[3];
        bilinear_u8<3>(pl.src_image, pl.src_h, pl.src_w, (long long)pl.src_w * 3, X, Y, px);
        uint8_t* d = pl.dst_image + dst_idx * 3;
        d[0] = px[0];
        d[1] = px[1];
        d[2] = px[2];
    } else if (pl.image_channels == 1) {
        uint8_t px[1];
        bilinear_u8<1>(pl.src_image, pl.src_h, pl.src_w, (long long)pl.src_w, X, Y, px);
        pl.dst_image[dst_idx] = px[0];
    } else if (pl.image_channels == 4) {
        uint8_t px[4];
        bilinear_u8<4>(pl.src_image, pl.src_h, pl.src_w, (long long)pl.src_w * 4, X, Y, px);
        *reinterpret_cast<uchar4*>(pl.dst_image + dst_idx * 4) =
            make_uchar4(px[0], px[1], px[2], px[3]);
    }
    if (pl.src_mask) {
        uint8_t m[1];
        bilinear_u8<1>(pl.src_mask, pl.src_h, pl.src_w, (long long)pl.src_w, X, Y, m);
        pl.dst_mask[dst_idx] = m[0];
    }
    if (pl.src_score) {
        pl.dst_score[dst_idx] = bilinear_f32(pl.src_score, pl.src_h, pl.src_w, pl.src_w, X, Y);
    }
}

// ============================================================================================
// Affine / perspective warp.  Block (32, 8) covers a 32 x 32 dst tile, 4 rows per thread.
// ============================================================================================
__global__ void __launch_bounds__(256) warp_fused_kernel(const vkb_warp_page* __restrict__ pages) {
    __shared__ vkb_warp_page pg;
    {
        const int tid = threadIdx.y * 32 + threadIdx.x;
        const int* src = reinterpret_cast<const int*>(pages + blockIdx.z);
        int* dst = reinterpret_cast<int*>(&pg);
        for (int i = tid; i < (int)(sizeof(vkb_warp_page) / 4); i += 256) dst[i] = src[i];
    }
    __syncthreads();
    const vkb_planes& pl = pg.planes;
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y_base = blockIdx.y * 32 + threadIdx.y;
    if (x >= pl.dst_w || y_base >= pl.dst_h) return;

    if (pg.kind == VKB_WARP_AFFINE) {
        const int adelta = cv_round_d(__dmul_rn(__dmul_rn(pg.inv[0], (double)x), 1024.0));
        const int bdelta = cv_round_d(__dmul_rn(__dmul_rn(pg.inv[3], (double)x), 1024.0));
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int y = y_base + 8 * k;
            if (y >= pl.dst_h) break;
            const int X0 = cv_round_d(__dmul_rn(__dadd_rn(__dmul_rn(pg.inv[1], (double)y), pg.inv[2]), 1024.0)) + 16;
            const int Y0 = cv_round_d(__dmul_rn(__dadd_rn(__dmul_rn(pg.inv[4], (double)y), pg.inv[5]), 1024.0)) + 16;
            sample_and_store(pl, x, y, (X0 + adelta) >> 5, (Y0 + bdelta) >> 5);
        }
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int y = y_base + 8 * k;
            if (y >= pl.dst_h) break;
            int X, Y;
            perspective_coord(pg.inv, x, y, X, Y);
            sample_and_store(pl, x, y, X, Y);
        }
    }
}

// ============================================================================================
// Lattice projection.
// ============================================================================================
__device__ __forceinline__ double block_sum(double v, double* scratch) {
    // blockDim.x threads (multiple of 32, <= 1024); result broadcast to all threads.
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    double total = 0.0;
    const int nw = blockDim.x >> 5;
    for (int i = 0; i < nw; ++i) total += scratch[i];
    return total;
}

__global__ void __launch_bounds__(1024) grid_project_camera_kernel(
    const vkb_grid_page* __restrict__ pages, int p_max, double* __restrict__ lattice_f) {
    __shared__ double scratch[32];
    const vkb_grid_page& pg = pages[blockIdx.x];
    if (pg.projector != VKB_PROJ_CAMERA) return;
    const int P = pg.rows * pg.cols;
    double* out = lattice_f + (size_t)blockIdx.x * p_max * 2;

    if (pg.strategy == VKB_CAM_PLANE) {
        for (int i = threadIdx.x; i < P; i += blockDim.x) {
            const int r = i / pg.cols, c = i - r * pg.cols;
            const double x = lattice_coord(c, pg.src_w, pg.grid_size);
            const double y = lattice_coord(r, pg.src_h, pg.grid_size);
            double u, v;
            project_point(pg.R, pg.t, pg.focal, x, y, 0.0, u, v);
            out[2 * i] = (double)(float)u;  // float32 in -> float32 out in cv.projectPoints
            out[2 * i + 1] = (double)(float)v;
        }
        return;
    }

    if (pg.strategy == VKB_CAM_CUBIC) {
        double local = 0.0;
        for (int i = threadIdx.x; i < P; i += blockDim.x) {
            const int r = i / pg.cols, c = i - r * pg.cols;
            const float x = (float)lattice_coord(c, pg.src_w, pg.grid_size);
            const float y = (float)lattice_coord(r, pg.src_h, pg.grid_size);
            const double z = cubic_z(pg, x, y);
            out[2 * i] = z;
            local += z;
        }
        const double mean = block_sum(local, scratch) / (double)P;
        for (int i = threadIdx.x; i < P; i += blockDim.x) {
            const int r = i / pg.cols, c = i - r * pg.cols;
            const double x = lattice_coord(c, pg.src_w, pg.grid_size);
            const double y = lattice_coord(r, pg.src_h, pg.grid_size);
            const double z = __dsub_rn(out[2 * i], mean);
            double u, v;
            project_point(pg.R, pg.t, pg.focal, x, y, z, u, v);
            out[2 * i] = u;  // float64 in -> float64 out
            out[2 * i + 1] = v;
        }
        return;
    }

    // plane line fold / curve
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        const int r = i / pg.cols, c = i - r * pg.cols;
        const float x = (float)lattice_coord(c, pg.src_w, pg.grid_size);
        const float y = (float)lattice_coord(r, pg.src_h, pg.grid_size);
        const double w = line_weight(pg, x, y);
        out[2 * i] = w;
        s0 += __dmul_rn(w, (double)pg.perturb[0]);
        s1 += __dmul_rn(w, (double)pg.perturb[1]);
        s2 += __dmul_rn(w, (double)pg.perturb[2]);
    }
    const double m0 = block_sum(s0, scratch) / (double)P;
    const double m1 = block_sum(s1, scratch) / (double)P;
    const double m2 = block_sum(s2, scratch) / (double)P;
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        const int r = i / pg.cols, c = i - r * pg.cols;
        const double x = lattice_coord(c, pg.src_w, pg.grid_size);
        const double y = lattice_coord(r, pg.src_h, pg.grid_size);
        const double w = out[2 * i];
        // np_3d_points (float32) += np_perturb (float64): add in double, store float32.
        const float X = (float)__dadd_rn(x, __dsub_rn(__dmul_rn(w, (double)pg.perturb[0]), m0));
        const float Y = (float)__dadd_rn(y, __dsub_rn(__dmul_rn(w, (double)pg.perturb[1]), m1));
        const float Z = (float)__dadd_rn(0.0, __dsub_rn(__dmul_rn(w, (double)pg.perturb[2]), m2));
        double u, v;
        project_point(pg.R, pg.t, pg.focal, (double)X, (double)Y, (double)Z, u, v);
        out[2 * i] = (double)(float)u;
        out[2 * i + 1] = (double)(float)v;
    }
}

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Similarity MLS: one warp per lattice point, lanes stride over the handles, shuffle
// reductions for the weight normalisation, centroids, mu and the 2-vector sum.
__global__ void __launch_bounds__(256) grid_project_mls_kernel(
    const vkb_grid_page* __restrict__ pages, int p_max, double* __restrict__ lattice_f) {
    const vkb_grid_page& pg = pages[blockIdx.y];
    if (pg.projector != VKB_PROJ_MLS) return;
    const int P = pg.rows * pg.cols;
    const int pt = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (pt >= P) return;
    const int lane = threadIdx.x & 31;
    const int n = pg.n_handles;
    const float* __restrict__ hs = pg.handles_src;
    const float* __restrict__ hd = pg.handles_dst;
    const int r = pt / pg.cols, c = pt - r * pg.cols;
    const float vx = (float)lattice_coord(c, pg.src_w, pg.grid_size);
    const float vy = (float)lattice_coord(r, pg.src_h, pg.grid_size);
    double* out = lattice_f + ((size_t)blockIdx.y * p_max + pt) * 2;

    float sum_inv = 0.f;
    int hit = -1;
    for (int i = lane; i < n; i += 32) {
        const float dx = hs[2 * i] - vx, dy = hs[2 * i + 1] - vy;
        const float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
        if (d2 == 0.f) hit = i;
        else sum_inv += 1.0f / d2;
    }
    const unsigned hit_mask = __ballot_sync(0xffffffffu, hit >= 0);
    if (hit_mask) {
        const int src_lane = __ffs(hit_mask) - 1;
        const int h = __shfl_sync(0xffffffffu, hit, src_lane);
        if (lane == 0) {
            out[0] = (double)hd[2 * h];
            out[1] = (double)hd[2 * h + 1];
        }
        return;
    }
    sum_inv = warp_sum_f(sum_inv);
    float pcx = 0.f, pcy = 0.f, qcx = 0.f, qcy = 0.f;
    for (int i = lane; i < n; i += 32) {
        const float dx = hs[2 * i] - vx, dy = hs[2 * i + 1] - vy;
        const float inv = 1.0f / __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
        const float w = inv / sum_inv;
        pcx = __fmaf_rn(w, hs[2 * i], pcx);
        pcy = __fmaf_rn(w, hs[2 * i + 1], pcy);
        qcx = __fmaf_rn(w, hd[2 * i], qcx);
        qcy = __fmaf_rn(w, hd[2 * i + 1], qcy);
    }
    pcx = warp_sum_f(pcx);
    pcy = warp_sum_f(pcy);
    qcx = warp_sum_f(qcx);
    qcy = warp_sum_f(qcy);
    const float ax = __fsub_rn(vx, pcx), ay = __fsub_rn(vy, pcy);
    float mu = 0.f, sx = 0.f, sy = 0.f;
    for (int i = lane; i < n; i += 32) {
        const float dx = hs[2 * i] - vx, dy = hs[2 * i + 1] - vy;
        const float inv = 1.0f / __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
        const float hx = __fsub_rn(hs[2 * i], pcx), hy = __fsub_rn(hs[2 * i + 1], pcy);
        const float qx = __fsub_rn(hd[2 * i], qcx), qy = __fsub_rn(hd[2 * i + 1], qcy);
        const float r00 = __fmaf_rn(hy, ay, __fmul_rn(hx, ax));
        const float r01 = __fmaf_rn(hy, -ax, __fmul_rn(hx, ay));
        const float r10 = __fmaf_rn(-hx, ay, __fmul_rn(hy, ax));
        const float r11 = __fmaf_rn(-hx, -ax, __fmul_rn(hy, ay));
        const float m00 = __fmul_rn(inv, r00), m01 = __fmul_rn(inv, r01);
        const float m10 = __fmul_rn(inv, r10), m11 = __fmul_rn(inv, r11);
        sx += __fadd_rn(__fmul_rn(qx, m00), __fmul_rn(qy, m10));
        sy += __fadd_rn(__fmul_rn(qx, m01), __fmul_rn(qy, m11));
        mu += __fmul_rn(inv, __fadd_rn(__fmul_rn(hx, hx), __fmul_rn(hy, hy)));
    }
    sx = warp_sum_f(sx);
    sy = warp_sum_f(sy);
    mu = warp_sum_f(mu);
    if (lane == 0) {
        out[0] = (double)__fadd_rn(sx / mu, qcx);
        out[1] = (double)__fadd_rn(sy / mu, qcy);
    }
}

// ============================================================================================
// Finalise: Point rounding (Python round = half to even), shift to origin, optional
// resize_as_src, result shape.  One block per page.
// ============================================================================================
__device__ __forceinline__ int block_reduce_int(int v, int* scratch, bool is_min) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const int other = __shfl_xor_sync(0xffffffffu, v, o);
        v = is_min ? min(v, other) : max(v, other);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    int total = scratch[0];
    const int nw = blockDim.x >> 5;
    for (int i = 1; i < nw; ++i) total = is_min ? min(total, scratch[i]) : max(total, scratch[i]);
    return total;
}

__global__ void __launch_bounds__(1024) grid_finalize_kernel(
    const vkb_grid_page* __restrict__ pages, int p_max, const double* __restrict__ lattice_f,
    int32_t* __restrict__ lattice_i, vkb_grid_meta* __restrict__ meta) {
    __shared__ int scratch[32];
    const vkb_grid_page& pg = pages[blockIdx.x];
    const int P = pg.rows * pg.cols;
    const double* in = lattice_f + (size_t)blockIdx.x * p_max * 2;
    int32_t* out = lattice_i + (size_t)blockIdx.x * p_max * 2;

    int mnx = INT32_MAX, mny = INT32_MAX;
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        mnx = min(mnx, cv_round_d(in[2 * i]));
        mny = min(mny, cv_round_d(in[2 * i + 1]));
    }
    const int shift_x = block_reduce_int(mnx, scratch, true);
    const int shift_y = block_reduce_int(mny, scratch, true);

    // to_shifted_point: round(smooth - shift) (point.py:80-84), NOT round(smooth) - shift.
    int mxx = INT32_MIN, mxy = INT32_MIN;
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        const int xi = cv_round_d(__dsub_rn(in[2 * i], (double)shift_x));
        const int yi = cv_round_d(__dsub_rn(in[2 * i + 1], (double)shift_y));
        out[2 * i] = xi;
        out[2 * i + 1] = yi;
        mxx = max(mxx, xi);
        mxy = max(mxy, yi);
    }
    int dst_w = block_reduce_int(mxx, scratch, false) + 1;
    int dst_h = block_reduce_int(mxy, scratch, false) + 1;
    double ratio_y = 1.0, ratio_x = 1.0;

    if (pg.resize_as_src) {
        // grid_creator.py:89-105: resize_val(v, size, resized) = clip(v * resized / size).
        ratio_y = (double)pg.src_h / (double)dst_h;
        ratio_x = (double)pg.src_w / (double)dst_w;
        const int raw_h = dst_h, raw_w = dst_w;
        for (int i = threadIdx.x; i < P; i += blockDim.x) {
            double sx = __dsub_rn(in[2 * i], (double)shift_x);
            double sy = __dsub_rn(in[2 * i + 1], (double)shift_y);
            sx = __ddiv_rn(__dmul_rn(sx, (double)pg.src_w), (double)raw_w);
            sy = __ddiv_rn(__dmul_rn(sy, (double)pg.src_h), (double)raw_h);
            sx = fmax(0.0, fmin(sx, (double)(pg.src_w - 1)));
            sy = fmax(0.0, fmin(sy, (double)(pg.src_h - 1)));
            out[2 * i] = cv_round_d(sx);
            out[2 * i + 1] = cv_round_d(sy);
        }
        dst_h = pg.src_h;
        dst_w = pg.src_w;
    }
    if (threadIdx.x == 0) {
        vkb_grid_meta m;
        m.dst_h = dst_h;
        m.dst_w = dst_w;
        m.shift_y = shift_y;
        m.shift_x = shift_x;
        m.resize_ratio_y = ratio_y;
        m.resize_ratio_x = ratio_x;
        m.status = 0;
        m.n_flagged_cells = 0;
        meta[blockIdx.x] = m;
    }
}

// ============================================================================================
// Cells: inverse (and optionally forward) homography, bbox, tile binning.  Thread per cell.
// ============================================================================================
__global__ void __launch_bounds__(128) grid_cells_kernel(
    const vkb_grid_page* __restrict__ pages, int p_max, int c_max, int t_max,
    const int32_t* __restrict__ lattice_i, vkb_grid_meta* __restrict__ meta,
    double* __restrict__ hinv, double* __restrict__ hfwd, int32_t* __restrict__ cell_box,
    CellLocal* __restrict__ cell_local, int32_t* __restrict__ tile_count,
    uint16_t* __restrict__ tile_cells) {
    const int page = blockIdx.y;
    const vkb_grid_page& pg = pages[page];
    const int ccols = pg.cols - 1;
    const int C = (pg.rows - 1) * ccols;
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= C) return;
    const int r = cell / ccols, c = cell - r * ccols;
    const int32_t* lat = lattice_i + (size_t)page * p_max * 2;
    const int i00 = r * pg.cols + c, i01 = i00 + 1, i11 = i00 + pg.cols + 1, i10 = i00 + pg.cols;
    // clockwise: (r,c) (r,c+1) (r+1,c+1) (r+1,c)   (type.py:107-116)
    int dx[4] = {lat[2 * i00], lat[2 * i01], lat[2 * i11], lat[2 * i10]};
    int dy[4] = {lat[2 * i00 + 1], lat[2 * i01 + 1], lat[2 * i11 + 1], lat[2 * i10 + 1]};
    const double sx0 = lattice_coord(c, pg.src_w, pg.grid_size);
    const double sx1 = lattice_coord(c + 1, pg.src_w, pg.grid_size);
    const double sy0 = lattice_coord(r, pg.src_h, pg.grid_size);
    const double sy1 = lattice_coord(r + 1, pg.src_h, pg.grid_size);
    const double sq[8] = {sx0, sy0, sx1, sy0, sx1, sy1, sx0, sy1};
    const double dq[8] = {(double)dx[0], (double)dy[0], (double)dx[1], (double)dy[1],
                          (double)dx[2], (double)dy[2], (double)dx[3], (double)dy[3]};
    double H[9];
    homography_4pt(dq, sq, H);
    double* ho = hinv + ((size_t)page * c_max + cell) * 9;
#pragma unroll
    for (int i = 0; i < 9; ++i) ho[i] = H[i];
    if (hfwd) {
        homography_4pt(sq, dq, H);
        double* hf = hfwd + ((size_t)page * c_max + cell) * 9;
#pragma unroll
        for (int i = 0; i < 9; ++i) hf[i] = H[i];
    }
    const int x0 = min(min(dx[0], dx[1]), min(dx[2], dx[3]));
    const int x1 = max(max(dx[0], dx[1]), max(dx[2], dx[3]));
    const int y0 = min(min(dy[0], dy[1]), min(dy[2], dy[3]));
    const int y1 = max(max(dy[0], dy[1]), max(dy[2], dy[3]));
    int32_t* box = cell_box + ((size_t)page * c_max + cell) * 4;
    box[0] = x0;
    box[1] = y0;
    box[2] = x1;
    box[3] = y1;
    // float32 re-centred form of the inverse map for the remap kernel's fast path
    {
        double Hi[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) Hi[i] = ho[i];
        CellLocal L;
        make_cell_local(Hi, (int)sx0, (int)sy0, x0, y0, L);
        cell_local[(size_t)page * c_max + cell] = L;
    }

    // bin into dst tiles
    const int tiles_x = (meta[page].dst_w + VKB_TILE - 1) / VKB_TILE;
    const int tx0 = x0 / VKB_TILE, tx1 = x1 / VKB_TILE, ty0 = y0 / VKB_TILE, ty1 = y1 / VKB_TILE;
    for (int ty = ty0; ty <= ty1; ++ty) {
        for (int tx = tx0; tx <= tx1; ++tx) {
            const int tile = ty * tiles_x + tx;
            if (tile >= t_max) continue;
            const int slot = atomicAdd(&tile_count[(size_t)page * t_max + tile], 1);
            if (slot < VKB_TILE_CAP) {
                tile_cells[((size_t)page * t_max + tile) * VKB_TILE_CAP + slot] = (uint16_t)cell;
            }
        }
    }
}

// ============================================================================================
// Coverage masks: warp per cell, lane per row of the cell's bounding box.
// ============================================================================================
__global__ void __launch_bounds__(128) grid_masks_kernel(
    const vkb_grid_page* __restrict__ pages, int p_max, int c_max,
    const int32_t* __restrict__ lattice_i, int32_t* __restrict__ cell_box,
    uint32_t* __restrict__ cell_masks) {
    const int page = blockIdx.y;
    const vkb_grid_page& pg = pages[page];
    const int ccols = pg.cols - 1;
    const int C = (pg.rows - 1) * ccols;
    const int cell = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (cell >= C) return;
    const int lane = threadIdx.x & 31;
    const int r = floor_div_small(cell, ccols, __fdividef(1.0f, (float)ccols));
    const int c = cell - r * ccols;
    const int32_t* lat = lattice_i + (size_t)page * p_max * 2;
    const int i00 = r * pg.cols + c, i01 = i00 + 1, i11 = i00 + pg.cols + 1, i10 = i00 + pg.cols;
    const int px[4] = {lat[2 * i00], lat[2 * i01], lat[2 * i11], lat[2 * i10]};
    const int py[4] = {lat[2 * i00 + 1], lat[2 * i01 + 1], lat[2 * i11 + 1], lat[2 * i10 + 1]};
    const int x0 = min(min(px[0], px[1]), min(px[2], px[3]));
    const int x1 = max(max(px[0], px[1]), max(px[2], px[3]));
    const int y0 = min(min(py[0], py[1]), min(py[2], py[3]));
    const int y1 = max(max(py[0], py[1]), max(py[2], py[3]));
    const int nrows = y1 - y0 + 1;
    const int nwords = (x1 - x0 + 32) / 32;
    uint32_t* out = cell_masks + ((size_t)page * c_max + cell) * VKB_CELL_MASK_WORDS;
    if (nwords != 1 || nrows > VKB_CELL_MASK_WORDS) {
        // too large for the fixed budget: the remap kernel rasterises this cell on the fly.
        if (lane == 0) cell_box[((size_t)page * c_max + cell) * 4 + 2] |= 0x40000000;
        return;
    }
    if (lane < nrows) {
        uint32_t word = 0;
        poly_row_mask<4>(px, py, y0 + lane, x0, &word, 1);
        out[lane] = word;
    }
}

// ============================================================================================
// Fused remap.  Block = 256 threads = 8 warps on one 32 x 32 dst tile; warp w owns rows
// 4w..4w+3, lane = column.
//
//   prologue  the tile's candidate cells (<= VKB_TILE_CAP) are staged in shared memory: bbox,
//             float64 inverse homography, and its float32 re-centred form (CellLocal);
//   owner     lane-parallel: lane i fetches candidate i's coverage words for the warp's 4 rows
//             (shifted to the tile's columns), a ballot keeps the candidates that touch the
//             band, and each survivor is broadcast with shuffles; every pixel keeps the
//             maximum cell index whose coverage bit is set (= last writer in row-major order);
//   coords    float32 fast path with a proven error band, float64 path for the <1% of pixels
//             that sit next to a rounding boundary (vkb_math.cuh);
//   gather    cv::remap's fixed-point bilinear for Image (C channels), Mask and ScoreMap.
// ============================================================================================
template <int C>
__device__ __forceinline__ void sample_u8(const uint8_t* __restrict__ src, int h, int w, int X, int Y,
                                          uint8_t* __restrict__ dst) {
    const int x0 = clamp_short(X >> kInterBits);
    const int y0 = clamp_short(Y >> kInterBits);
    const int fx = X & (kInterTab - 1);
    const int fy = Y & (kInterTab - 1);
    const int pitch = w * C;  // a page plane is < 2 GiB (checked by the launcher)
    int p00[C], p01[C], p10[C], p11[C];
    if ((unsigned)x0 < (unsigned)(w - 1) && (unsigned)y0 < (unsigned)(h - 1)) {  // interior
        const uint8_t* r0 = src + (y0 * pitch + x0 * C);
        const uint8_t* r1 = r0 + pitch;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            p00[c] = __ldg(r0 + c);
            p01[c] = __ldg(r0 + C + c);
            p10[c] = __ldg(r1 + c);
            p11[c] = __ldg(r1 + C + c);
        }
    } else {
        const bool in_x0 = (unsigned)x0 < (unsigned)w, in_x1 = (unsigned)(x0 + 1) < (unsigned)w;
        const bool in_y0 = (unsigned)y0 < (unsigned)h, in_y1 = (unsigned)(y0 + 1) < (unsigned)h;
        const uint8_t* r0 = src + ((long long)y0 * pitch + (long long)x0 * C);
        const uint8_t* r1 = r0 + pitch;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            p00[c] = (in_y0 && in_x0) ? r0[c] : 0;
            p01[c] = (in_y0 && in_x1) ? r0[C + c] : 0;
            p10[c] = (in_y1 && in_x0) ? r1[c] : 0;
            p11[c] = (in_y1 && in_x1) ? r1[C + c] : 0;
        }
    }
    const int gx = kInterTab - fx, gy = kInterTab - fy;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const int a = gx * p00[c] + fx * p01[c];
        const int b = gx * p10[c] + fx * p11[c];
        dst[c] = (uint8_t)((gy * a + fy * b + 512) >> 10);
    }
}

static_assert(sizeof(CellLocal) == VKB_CELL_LOCAL_BYTES, "CellLocal layout is part of the ABI");

struct RemapShared {
    CellLocal loc[VKB_TILE_CAP];
    int box[VKB_TILE_CAP][4];
    int cell[VKB_TILE_CAP];
};

// Rare paths are kept out of line so the hot loop stays small (instruction cache).
__device__ __noinline__ void cell_coord_exact(const double* __restrict__ H, int x, int y, int* X,
                                              int* Y) {
    cell_coord(H, x, y, *X, *Y);
}

// coverage of one over-budget cell on row y, restricted to the 32 columns starting at tx0
__device__ __noinline__ uint32_t cell_row_window_slow(const int32_t* __restrict__ lat, int cols,
                                                      int ccols, int cell, int y, int tx0) {
    const int r = cell / ccols, c = cell - r * ccols;
    const int i00 = r * cols + c, i01 = i00 + 1, i11 = i00 + cols + 1, i10 = i00 + cols;
    const int px[4] = {lat[2 * i00], lat[2 * i01], lat[2 * i11], lat[2 * i10]};
    const int py[4] = {lat[2 * i00 + 1], lat[2 * i01 + 1], lat[2 * i11 + 1], lat[2 * i10 + 1]};
    uint32_t bits = 0;
    poly_row_mask<4>(px, py, y, tx0, &bits, 1);
    return bits;
}

__device__ __forceinline__ uint32_t cell_row_window(const uint32_t* __restrict__ cell_masks,
                                                    const int32_t* __restrict__ lat, int cols,
                                                    int ccols, int cell, bool flagged, int bx0,
                                                    int by0, int y, int tx0) {
    if (flagged) return cell_row_window_slow(lat, cols, ccols, cell, y, tx0);
    const uint32_t w = cell_masks[cell * VKB_CELL_MASK_WORDS + (y - by0)];
    const int rel = tx0 - bx0;  // in (-32, 32) because the boxes overlap
    return rel >= 0 ? (w >> rel) : (w << (-rel));
}

template <int C, bool MASK, bool SCORE>
__global__ void __launch_bounds__(256, (C == 3 && !MASK && !SCORE) ? 5 : 4) grid_remap_kernel(
    const vkb_grid_page* __restrict__ pages, const vkb_planes* __restrict__ planes,
    int c_max, int t_max, const vkb_grid_meta* __restrict__ meta, const double* __restrict__ hinv,
    const int32_t* __restrict__ cell_box, const CellLocal* __restrict__ cell_local,
    const uint32_t* __restrict__ cell_masks,
    const int32_t* __restrict__ tile_count, const uint16_t* __restrict__ tile_cells,
    const int32_t* __restrict__ lattice_i, int p_max) {
    const int page = blockIdx.z;
    const int dst_h = meta[page].dst_h, dst_w = meta[page].dst_w;
    const int tiles_x = (dst_w + VKB_TILE - 1) / VKB_TILE;
    const int tiles_y = (dst_h + VKB_TILE - 1) / VKB_TILE;
    if ((int)blockIdx.x >= tiles_x || (int)blockIdx.y >= tiles_y) return;

    __shared__ RemapShared sm;
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int tile = blockIdx.y * tiles_x + blockIdx.x;
    const int count = tile_count[(size_t)page * t_max + tile];
    const bool fast = count <= VKB_TILE_CAP;
    const vkb_planes pl = planes[page];
    const int cols = pages[page].cols, rows = pages[page].rows;
    const int src_h = pl.src_h, src_w = pl.src_w;
    const int ccols = cols - 1;
    const int n_cells = (rows - 1) * ccols;
    const int32_t* lat = lattice_i + (size_t)page * p_max * 2;
    const uint32_t* page_masks = cell_masks + (size_t)page * c_max * VKB_CELL_MASK_WORDS;
    const int32_t* page_box = cell_box + (size_t)page * c_max * 4;
    const double* page_hinv = hinv + (size_t)page * c_max * 9;

    if (fast && tid < count) {
        const int cell = tile_cells[((size_t)page * t_max + tile) * VKB_TILE_CAP + tid];
        const int4 b = *reinterpret_cast<const int4*>(page_box + (size_t)cell * 4);
        sm.box[tid][0] = b.x; sm.box[tid][1] = b.y; sm.box[tid][2] = b.z; sm.box[tid][3] = b.w;
        sm.cell[tid] = cell;
        const int4* lsrc = reinterpret_cast<const int4*>(cell_local + (size_t)page * c_max + cell);
        int4* ldst = reinterpret_cast<int4*>(&sm.loc[tid]);
        ldst[0] = lsrc[0];
        ldst[1] = lsrc[1];
        ldst[2] = lsrc[2];
    }
    __syncthreads();

    const int tx0 = blockIdx.x * VKB_TILE;
    const int x = tx0 + lane;
    const int ry0 = blockIdx.y * VKB_TILE + warp * 4;
    if (ry0 >= dst_h) return;

    // ---- owner ---------------------------------------------------------------------------
    int key[4] = {-1, -1, -1, -1};
    const int n_cand = fast ? count : n_cells;
    for (int base = 0; base < n_cand; base += 32) {
        const int s = base + lane;
        uint32_t win[4] = {0u, 0u, 0u, 0u};
        int my_key = -1;
        if (s < n_cand) {
            int bx0, by0, bx1, by1, cell;
            if (fast) {
                bx0 = sm.box[s][0]; by0 = sm.box[s][1]; bx1 = sm.box[s][2]; by1 = sm.box[s][3];
                cell = sm.cell[s];
            } else {
                const int4 b = *reinterpret_cast<const int4*>(page_box + (size_t)s * 4);
                bx0 = b.x; by0 = b.y; bx1 = b.z; by1 = b.w;
                cell = s;
            }
            const bool flagged = (bx1 & 0x40000000) != 0;
            bx1 &= 0x3FFFFFFF;
            if (!(by1 < ry0 || by0 > ry0 + 3 || bx1 < tx0 || bx0 > tx0 + 31)) {
                my_key = fast ? ((cell << 6) | s) : (cell << 6);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int y = ry0 + j;
                    if (y >= by0 && y <= by1)
                        win[j] = cell_row_window(page_masks, lat, cols, ccols, cell, flagged, bx0,
                                                 by0, y, tx0);
                }
            }
        }
        unsigned active = __ballot_sync(0xffffffffu, (win[0] | win[1] | win[2] | win[3]) != 0u);
        while (active) {
            const int src_lane = __ffs(active) - 1;
            active &= active - 1;
            const int k = __shfl_sync(0xffffffffu, my_key, src_lane);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t w = __shfl_sync(0xffffffffu, win[j], src_lane);
                key[j] = max(key[j], ((w >> lane) & 1u) ? k : -1);
            }
        }
    }
    if (x >= dst_w) return;

    // ---- coordinates + gather ------------------------------------------------------------
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int y = ry0 + j;
        if (y >= dst_h) break;
        int X = 0, Y = 0;
        if (key[j] >= 0) {
            if (fast) {
                const int slot = key[j] & 63;
                if (!cell_coord_fast(sm.loc[slot], x, y, X, Y))
                    cell_coord_exact(page_hinv + (size_t)(key[j] >> 6) * 9, x, y, &X, &Y);
            } else {
                cell_coord_exact(page_hinv + (size_t)(key[j] >> 6) * 9, x, y, &X, &Y);
            }
        }
        const int di = y * dst_w + x;
        if (C > 0) {
            uint8_t px[C > 0 ? C : 1];
            sample_u8<(C > 0 ? C : 1)>(pl.src_image, src_h, src_w, X, Y, px);
            uint8_t* d = pl.dst_image + di * C;
            if (C == 4) {
                *reinterpret_cast<uchar4*>(d) = make_uchar4(px[0], px[1], px[2], px[3 % (C > 0 ? C : 1)]);
            } else {
#pragma unroll
                for (int c = 0; c < C; ++c) d[c] = px[c];
            }
        }
        if (MASK) {
            uint8_t m[1];
            sample_u8<1>(pl.src_mask, src_h, src_w, X, Y, m);
            pl.dst_mask[di] = m[0];
        }
        if (SCORE) pl.dst_score[di] = bilinear_f32(pl.src_score, src_h, src_w, src_w, X, Y);
    }
}

// ============================================================================================
// Points.
// ============================================================================================
__global__ void grid_points_kernel(const double* __restrict__ hfwd, int ccols,
                                   const double* __restrict__ xy_in,
                                   const int32_t* __restrict__ cell_rc, double* __restrict__ xy_out,
                                   int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* H = hfwd + ((size_t)cell_rc[2 * i] * ccols + cell_rc[2 * i + 1]) * 9;
    const double x = xy_in[2 * i], y = xy_in[2 * i + 1];
    // np.matmul(trans_mat, (x, y, 1.0)) in double (interface.py:211-216)
    const double tx = __dadd_rn(__dadd_rn(__dmul_rn(H[0], x), __dmul_rn(H[1], y)), H[2]);
    const double ty = __dadd_rn(__dadd_rn(__dmul_rn(H[3], x), __dmul_rn(H[4], y)), H[5]);
    const double t = __dadd_rn(__dadd_rn(__dmul_rn(H[6], x), __dmul_rn(H[7], y)), H[8]);
    xy_out[2 * i] = __ddiv_rn(tx, t);
    xy_out[2 * i + 1] = __ddiv_rn(ty, t);
}

struct Mat9 {
    double m[9];
};

__global__ void affine_points_kernel(Mat9 M, int rows, const double* __restrict__ xy_in,
                                     double* __restrict__ xy_out, int n, int f32_math) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (f32_math) {
        // affine_np_points with a float32 2x3 matrix and float32 points: sgemm, K = 3.
        const float x = (float)xy_in[2 * i], y = (float)xy_in[2 * i + 1];
        const float ox = __fmaf_rn((float)M.m[2], 1.0f, __fmaf_rn((float)M.m[1], y, __fmul_rn((float)M.m[0], x)));
        const float oy = __fmaf_rn((float)M.m[5], 1.0f, __fmaf_rn((float)M.m[4], y, __fmul_rn((float)M.m[3], x)));
        xy_out[2 * i] = (double)ox;
        xy_out[2 * i + 1] = (double)oy;
        return;
    }
    const double x = xy_in[2 * i], y = xy_in[2 * i + 1];
    const double tx = fma(M.m[2], 1.0, fma(M.m[1], y, __dmul_rn(M.m[0], x)));
    const double ty = fma(M.m[5], 1.0, fma(M.m[4], y, __dmul_rn(M.m[3], x)));
    if (rows == 2) {
        xy_out[2 * i] = tx;
        xy_out[2 * i + 1] = ty;
        return;
    }
    const double t = fma(M.m[8], 1.0, fma(M.m[7], y, __dmul_rn(M.m[6], x)));
    xy_out[2 * i] = __ddiv_rn(tx, t);
    xy_out[2 * i + 1] = __ddiv_rn(ty, t);
}

// ============================================================================================
// cv.fillPoly of one arbitrary polygon (active mask = dst lattice border polygon).
// Row kernel: crossings of all edges, sorted, spans filled.  Edge kernel: Bresenham outline.
// ============================================================================================
constexpr int kMaxCross = 32;

__global__ void fill_polygon_rows_kernel(uint8_t* __restrict__ mask, int h, int w,
                                         const int32_t* __restrict__ poly, int n, uint8_t value) {
    const int y = blockIdx.x * blockDim.x + threadIdx.x;
    if (y >= h) return;
    long long cross[kMaxCross];
    int nc = 0;
    for (int i = 0; i < n; ++i) {
        const int j = (i + n - 1) % n;
        const int x0 = poly[2 * j], y0 = poly[2 * j + 1], x1 = poly[2 * i], y1 = poly[2 * i + 1];
        if (y0 == y1) continue;
        int ya, yb;
        long long xa;
        if (y0 < y1) { ya = y0; yb = y1; xa = (long long)x0 * 65536; }
        else { ya = y1; yb = y0; xa = (long long)x1 * 65536; }
        if (ya <= y && y < yb) {
            const long long dxf = ((long long)(x1 - x0) * 65536) / (long long)(y1 - y0);
            if (nc < kMaxCross) cross[nc] = xa + dxf * (long long)(y - ya);
            ++nc;
        }
    }
    if (nc > kMaxCross) nc = kMaxCross;  // pathological polygons only
    for (int i = 1; i < nc; ++i) {
        const long long v = cross[i];
        int j = i - 1;
        while (j >= 0 && cross[j] > v) { cross[j + 1] = cross[j]; --j; }
        cross[j + 1] = v;
    }
    uint8_t* row = mask + (size_t)y * w;
    for (int k = 0; k + 1 < nc; k += 2) {
        long long xl = (cross[k] + 65535) >> 16;
        long long xr = cross[k + 1] >> 16;
        if (xl < 0) xl = 0;
        if (xr > w - 1) xr = w - 1;
        for (long long xx = xl; xx <= xr; ++xx) row[xx] = value;
    }
}

__global__ void fill_polygon_edges_kernel(uint8_t* __restrict__ mask, int h, int w,
                                          const int32_t* __restrict__ poly, int n, uint8_t value) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int j = (i + n - 1) % n;
    int ax = poly[2 * j], ay = poly[2 * j + 1], bx = poly[2 * i], by = poly[2 * i + 1];
    if (ax > bx) { int t = ax; ax = bx; bx = t; t = ay; ay = by; by = t; }
    int dx = bx - ax, dy = by - ay;
    const int sy = dy >= 0 ? 1 : -1;
    dy = dy >= 0 ? dy : -dy;
    const bool steep = dy > dx;
    if (steep) { const int t = dx; dx = dy; dy = t; }
    int err = dx - 2 * dy;
    int x = ax, y = ay;
    for (int s = 0; s <= dx; ++s) {
        if ((unsigned)x < (unsigned)w && (unsigned)y < (unsigned)h) mask[(size_t)y * w + x] = value;
        const bool m = err < 0;
        err += -2 * dy + (m ? 2 * dx : 0);
        if (steep) { y += sy; if (m) x += 1; }
        else { x += 1; if (m) y += sy; }
    }
}

}  // namespace vkb

// ============================================================================================
// C ABI
// ============================================================================================
using namespace vkb;

extern "C" int vkb_version(void) { return 1; }
extern "C" const char* vkb_last_error(void) { return g_error; }

extern "C" int vkb_warp_fused(const vkb_warp_page* pages, int32_t n_pages, int32_t max_dst_h,
                              int32_t max_dst_w, void* stream) {
    VKB_REQUIRE(pages != nullptr && n_pages > 0, "no pages");
    VKB_REQUIRE(n_pages <= 65535, "at most 65535 pages per launch");
    VKB_REQUIRE(max_dst_h > 0 && max_dst_w > 0, "empty destination");
    dim3 grid((max_dst_w + 31) / 32, (max_dst_h + 31) / 32, n_pages);
    warp_fused_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(pages);
    return check_launch("warp_fused_kernel");
}

extern "C" int vkb_affine_points(const double* mat_host, int32_t rows, const double* xy_in,
                                 double* xy_out, int32_t n, int32_t f32_math, void* stream) {
    VKB_REQUIRE(rows == 2 || rows == 3, "rows must be 2 or 3");
    if (n <= 0) return VKB_OK;
    Mat9 M;
    for (int i = 0; i < 9; ++i) M.m[i] = i < rows * 3 ? mat_host[i] : (i == 8 ? 1.0 : 0.0);
    affine_points_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(M, rows, xy_in, xy_out,
                                                                            n, f32_math);
    return check_launch("affine_points_kernel");
}

extern "C" int vkb_grid_project(const vkb_grid_page* pages, int32_t n_pages, int32_t p_max,
                                double* lattice_f, void* stream) {
    VKB_REQUIRE(pages && lattice_f && n_pages > 0 && p_max > 0, "bad arguments");
    VKB_REQUIRE(n_pages <= 65535, "at most 65535 pages per launch");
    grid_project_camera_kernel<<<n_pages, 1024, 0, (cudaStream_t)stream>>>(pages, p_max, lattice_f);
    int rc = check_launch("grid_project_camera_kernel");
    if (rc) return rc;
    grid_project_mls_kernel<<<dim3((p_max + 7) / 8, n_pages), 256, 0, (cudaStream_t)stream>>>(
        pages, p_max, lattice_f);
    return check_launch("grid_project_mls_kernel");
}

extern "C" int vkb_grid_finalize(const vkb_grid_page* pages, int32_t n_pages, int32_t p_max,
                                 const double* lattice_f, int32_t* lattice_i, vkb_grid_meta* meta,
                                 void* stream) {
    VKB_REQUIRE(pages && lattice_f && lattice_i && meta && n_pages > 0, "bad arguments");
    grid_finalize_kernel<<<n_pages, 1024, 0, (cudaStream_t)stream>>>(pages, p_max, lattice_f,
                                                                    lattice_i, meta);
    return check_launch("grid_finalize_kernel");
}

extern "C" int vkb_grid_build(const vkb_grid_page* pages, int32_t n_pages, int32_t p_max,
                              int32_t c_max, int32_t t_max, const int32_t* lattice_i,
                              vkb_grid_meta* meta, double* hinv, double* hfwd, int32_t* cell_box,
                              void* cell_local, uint32_t* cell_masks, int32_t* tile_count,
                              uint16_t* tile_cells, void* stream) {
    VKB_REQUIRE(pages && lattice_i && meta && hinv && cell_box && cell_local && cell_masks
                    && tile_count && tile_cells, "bad arguments");
    VKB_REQUIRE(n_pages > 0 && n_pages <= 65535, "1..65535 pages per launch");
    VKB_REQUIRE(c_max > 0 && c_max <= 65535, "at most 65535 cells per page");
    cudaStream_t st = (cudaStream_t)stream;
    VKB_CUDA(cudaMemsetAsync(tile_count, 0, sizeof(int32_t) * (size_t)n_pages * t_max, st));
    grid_cells_kernel<<<dim3((c_max + 127) / 128, n_pages), 128, 0, st>>>(
        pages, p_max, c_max, t_max, lattice_i, meta, hinv, hfwd, cell_box,
        reinterpret_cast<CellLocal*>(cell_local), tile_count, tile_cells);
    int rc = check_launch("grid_cells_kernel");
    if (rc) return rc;
    grid_masks_kernel<<<dim3((c_max + 3) / 4, n_pages), 128, 0, st>>>(pages, p_max, c_max, lattice_i,
                                                                     cell_box, cell_masks);
    return check_launch("grid_masks_kernel");
}

extern "C" int vkb_grid_remap(const vkb_grid_page* pages, const vkb_planes* planes, int32_t n_pages,
                              int32_t p_max, int32_t c_max, int32_t t_max,
                              const int32_t* lattice_i, const vkb_grid_meta* meta,
                              const double* hinv, const int32_t* cell_box,
                              const void* cell_local, const uint32_t* cell_masks,
                              const int32_t* tile_count, const uint16_t* tile_cells,
                              int32_t max_dst_h, int32_t max_dst_w, int32_t image_channels,
                              int32_t has_mask, int32_t has_score, void* stream) {
    VKB_REQUIRE(pages && planes && lattice_i && meta && hinv && cell_box && cell_local && cell_masks
                    && tile_count && tile_cells, "bad arguments");
    VKB_REQUIRE(n_pages > 0 && n_pages <= 65535, "1..65535 pages per launch");
    VKB_REQUIRE(max_dst_h > 0 && max_dst_w > 0, "empty destination");
    VKB_REQUIRE(image_channels == 0 || image_channels == 1 || image_channels == 3
                    || image_channels == 4, "image_channels must be 0, 1, 3 or 4");
    VKB_REQUIRE(image_channels || has_mask || has_score, "nothing to remap");
    dim3 grid((max_dst_w + VKB_TILE - 1) / VKB_TILE, (max_dst_h + VKB_TILE - 1) / VKB_TILE, n_pages);
    VKB_REQUIRE((long long)grid.x * grid.y <= t_max, "t_max smaller than the tile grid");
    cudaStream_t st = (cudaStream_t)stream;
#define VKB_LAUNCH_REMAP(CH, M, S)                                                             \
    grid_remap_kernel<CH, M, S><<<grid, 256, 0, st>>>(                                         \
        pages, planes, c_max, t_max, meta, hinv, cell_box,                                     \
        reinterpret_cast<const CellLocal*>(cell_local), cell_masks, tile_count, tile_cells,    \
        lattice_i, p_max)
    const int key = image_channels * 4 + (has_mask ? 2 : 0) + (has_score ? 1 : 0);
    switch (key) {
        case 0 * 4 + 1: VKB_LAUNCH_REMAP(0, false, true); break;
        case 0 * 4 + 2: VKB_LAUNCH_REMAP(0, true, false); break;
        case 0 * 4 + 3: VKB_LAUNCH_REMAP(0, true, true); break;
        case 1 * 4 + 0: VKB_LAUNCH_REMAP(1, false, false); break;
        case 1 * 4 + 1: VKB_LAUNCH_REMAP(1, false, true); break;
        case 1 * 4 + 2: VKB_LAUNCH_REMAP(1, true, false); break;
        case 1 * 4 + 3: VKB_LAUNCH_REMAP(1, true, true); break;
        case 3 * 4 + 0: VKB_LAUNCH_REMAP(3, false, false); break;
        case 3 * 4 + 1: VKB_LAUNCH_REMAP(3, false, true); break;
        case 3 * 4 + 2: VKB_LAUNCH_REMAP(3, true, false); break;
        case 3 * 4 + 3: VKB_LAUNCH_REMAP(3, true, true); break;
        case 4 * 4 + 0: VKB_LAUNCH_REMAP(4, false, false); break;
        case 4 * 4 + 1: VKB_LAUNCH_REMAP(4, false, true); break;
        case 4 * 4 + 2: VKB_LAUNCH_REMAP(4, true, false); break;
        case 4 * 4 + 3: VKB_LAUNCH_REMAP(4, true, true); break;
        default: VKB_REQUIRE(false, "unsupported container combination");
    }
#undef VKB_LAUNCH_REMAP
    return check_launch("grid_remap_kernel");
}

extern "C" int vkb_grid_points(const double* hfwd_page, int32_t cols_minus_1, const double* xy_in,
                               const int32_t* cell_rc, double* xy_out, int32_t n, void* stream) {
    VKB_REQUIRE(hfwd_page && xy_in && cell_rc && xy_out, "bad arguments");
    if (n <= 0) return VKB_OK;
    grid_points_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(hfwd_page, cols_minus_1,
                                                                          xy_in, cell_rc, xy_out, n);
    return check_launch("grid_points_kernel");
}

extern "C" int vkb_fill_polygon(uint8_t* mask, int32_t h, int32_t w, const int32_t* poly_xy,
                                int32_t n_pts, uint8_t value, void* stream) {
    VKB_REQUIRE(mask && poly_xy && h > 0 && w > 0 && n_pts >= 1, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    fill_polygon_rows_kernel<<<(h + 127) / 128, 128, 0, st>>>(mask, h, w, poly_xy, n_pts, value);
    int rc = check_launch("fill_polygon_rows_kernel");
    if (rc) return rc;
    fill_polygon_edges_kernel<<<(n_pts + 127) / 128, 128, 0, st>>>(mask, h, w, poly_xy, n_pts, value);
    return check_launch("fill_polygon_edges_kernel");
}
