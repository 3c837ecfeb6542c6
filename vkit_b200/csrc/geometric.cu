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
#include <stdlib.h>
#include <mutex>
#include <cstdlib>
#include "common.cuh"
#include "vkb_math.cuh"
#include "vkb_lattice.cuh"
#include "vkb_grid.cuh"
#include "vkb_gather.cuh"

namespace vkb {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
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

// The same projection with ONE THREAD per lattice point (second generation, the default).  The
// warp-per-point form above spends ~415 warp instructions per point, most of them on lanes without
// a handle and on 8 x 5 shuffle steps; here a thread walks the handles itself and reduces them in
// the SAME order as the butterfly did, so the float32 results are bit-identical
// (tests/test_gpu_parity.py::test_mls_projection_generations_agree): lane L of the old kernel
// accumulated handles L, L + 32, ... and the xor butterfly then added the 32 lane values as a
// perfect binary tree over the lanes in bit-reversed order -- that tree is evaluated with a
// five-level merge stack while the slots are visited in bit-reversed order.
constexpr int kMlsSmemHandles = 64;

template <int R>
struct MlsTree {
    float st[R][5];
    float total[R];
    // slot number `k` of the bit-reversed visit carries the values v[0..R-1]
    __device__ __forceinline__ void push(int k, const float* v) {
#pragma unroll
        for (int q = 0; q < R; ++q) {
            float x = v[q];
            bool placed = false;
#pragma unroll
            for (int lev = 0; lev < 5; ++lev) {
                if (placed) continue;
                if ((k >> lev) & 1) {
                    x = st[q][lev] + x;
                } else {
                    st[q][lev] = x;
                    placed = true;
                }
            }
            if (!placed) total[q] = x;  // k == 31
        }
    }
};

__global__ void __launch_bounds__(128) grid_project_mls_points_kernel(
    const vkb_grid_page* __restrict__ pages, int p_max, double* __restrict__ lattice_f) {
    __shared__ float s_hs[2 * kMlsSmemHandles], s_hd[2 * kMlsSmemHandles];
    const vkb_grid_page& pg = pages[blockIdx.y];
    if (pg.projector != VKB_PROJ_MLS) return;  // block uniform
    const int n = pg.n_handles;
    for (int i = threadIdx.x; i < 2 * min(n, kMlsSmemHandles); i += blockDim.x) {
        s_hs[i] = pg.handles_src[i];
        s_hd[i] = pg.handles_dst[i];
    }
    __syncthreads();
    const int P = pg.rows * pg.cols;
    const int pt = blockIdx.x * blockDim.x + threadIdx.x;
    if (pt >= P) return;
    const float* __restrict__ g_hs = pg.handles_src;
    const float* __restrict__ g_hd = pg.handles_dst;
    auto hs = [&](int j) { return j < 2 * kMlsSmemHandles ? s_hs[j] : g_hs[j]; };
    auto hd = [&](int j) { return j < 2 * kMlsSmemHandles ? s_hd[j] : g_hd[j]; };
    const int r = pt / pg.cols, c = pt - r * pg.cols;
    const float vx = (float)lattice_coord(c, pg.src_w, pg.grid_size);
    const float vy = (float)lattice_coord(r, pg.src_h, pg.grid_size);
    double* out = lattice_f + ((size_t)blockIdx.y * p_max + pt) * 2;

    // pass 0: sum of the inverse squared distances; an exact hit returns the handle's target
    // (the lowest slot with a hit wins, inside a slot the last one: the old kernel's ballot)
    int hit = -1;
    MlsTree<1> t0;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
        const int L = (int)(__brev((unsigned)k) >> 27);
        float sum_inv = 0.f;
        int slot_hit = -1;
        for (int i = L; i < n; i += 32) {
            const float dx = hs(2 * i) - vx, dy = hs(2 * i + 1) - vy;
            const float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
            if (d2 == 0.f) slot_hit = i;
            else sum_inv += 1.0f / d2;
        }
        if (slot_hit >= 0 && (hit < 0 || L < (hit & 31))) hit = slot_hit;
        t0.push(k, &sum_inv);
    }
    if (hit >= 0) {
        out[0] = (double)hd(2 * hit);
        out[1] = (double)hd(2 * hit + 1);
        return;
    }
    const float sum_inv = t0.total[0];

    // pass 1: weighted centroids of the source and target handles
    MlsTree<4> t1;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
        const int L = (int)(__brev((unsigned)k) >> 27);
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        for (int i = L; i < n; i += 32) {
            const float px = hs(2 * i), py = hs(2 * i + 1);
            const float dx = px - vx, dy = py - vy;
            const float inv = 1.0f / __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
            const float w = inv / sum_inv;
            v[0] = __fmaf_rn(w, px, v[0]);
            v[1] = __fmaf_rn(w, py, v[1]);
            v[2] = __fmaf_rn(w, hd(2 * i), v[2]);
            v[3] = __fmaf_rn(w, hd(2 * i + 1), v[3]);
        }
        t1.push(k, v);
    }
    const float pcx = t1.total[0], pcy = t1.total[1], qcx = t1.total[2], qcy = t1.total[3];
    const float ax = __fsub_rn(vx, pcx), ay = __fsub_rn(vy, pcy);

    // pass 2: mu and the 2-vector sum
    MlsTree<3> t2;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
        const int L = (int)(__brev((unsigned)k) >> 27);
        float v[3] = {0.f, 0.f, 0.f};  // sx, sy, mu
        for (int i = L; i < n; i += 32) {
            const float px = hs(2 * i), py = hs(2 * i + 1);
            const float dx = px - vx, dy = py - vy;
            const float inv = 1.0f / __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
            const float hx = __fsub_rn(px, pcx), hy = __fsub_rn(py, pcy);
            const float qx = __fsub_rn(hd(2 * i), qcx), qy = __fsub_rn(hd(2 * i + 1), qcy);
            const float r00 = __fmaf_rn(hy, ay, __fmul_rn(hx, ax));
            const float r01 = __fmaf_rn(hy, -ax, __fmul_rn(hx, ay));
            const float r10 = __fmaf_rn(-hx, ay, __fmul_rn(hy, ax));
            const float r11 = __fmaf_rn(-hx, -ax, __fmul_rn(hy, ay));
            const float m00 = __fmul_rn(inv, r00), m01 = __fmul_rn(inv, r01);
            const float m10 = __fmul_rn(inv, r10), m11 = __fmul_rn(inv, r11);
            v[0] += __fadd_rn(__fmul_rn(qx, m00), __fmul_rn(qy, m10));
            v[1] += __fadd_rn(__fmul_rn(qx, m01), __fmul_rn(qy, m11));
            v[2] += __fmul_rn(inv, __fadd_rn(__fmul_rn(hx, hx), __fmul_rn(hy, hy)));
        }
        t2.push(k, v);
    }
    const float sx = t2.total[0], sy = t2.total[1], mu = t2.total[2];
    out[0] = (double)__fadd_rn(sx / mu, qcx);
    out[1] = (double)__fadd_rn(sy / mu, qcy);
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
    int32_t* __restrict__ lattice_i, vkb_grid_meta* __restrict__ meta,
    vkb_grid_meta* __restrict__ meta_mirror) {
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
        if (meta_mirror) meta_mirror[blockIdx.x] = m;  // mapped pinned host memory: no D2H copy
    }
}

// ============================================================================================
// Output layout on the device: exclusive scan of the pages' result pixels, destination pointers
// and result shapes written into the plane records -- so a batch needs no host round trip between
// the lattice projection and the remap (the host reads shapes and offsets when the batch is done).
// Capacities are the caller's bet; when a page needs more tiles than the workspaces hold or the
// batch more pixels than the arenas, every result shape is zeroed (all later kernels then have
// nothing to do) and the status word tells the host to run again with exact sizes.
// ============================================================================================
__global__ void __launch_bounds__(1024) grid_layout_kernel(
    vkb_grid_meta* __restrict__ meta, int n_pages, vkb_planes* __restrict__ planes,
    long long cap_pixels, int t_max, long long* __restrict__ layout,
    long long* __restrict__ layout_mirror) {
    __shared__ long long warp_sums[32];
    __shared__ long long carry;
    __shared__ int bad;
    if (threadIdx.x == 0) {
        carry = 0;
        bad = 0;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < n_pages; base += 1024) {
        const int page = base + threadIdx.x;
        long long v = 0;
        if (page < n_pages) {
            v = (long long)meta[page].dst_h * meta[page].dst_w;
            if (page_tiles(meta[page]) > t_max) atomicOr(&bad, 2);
        }
        long long inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const long long o = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += o;
        }
        if (lane == 31) warp_sums[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            long long ws = warp_sums[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const long long o = __shfl_up_sync(0xffffffffu, ws, d);
                if (lane >= d) ws += o;
            }
            warp_sums[lane] = ws;
        }
        __syncthreads();
        const long long ex = carry + (warp > 0 ? warp_sums[warp - 1] : 0) + inc - v;
        const long long total = warp_sums[31];
        __syncthreads();
        if (threadIdx.x == 0) carry += total;
        if (page < n_pages) {
            layout[page] = ex;
            if (layout_mirror) layout_mirror[page] = ex;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0 && carry > cap_pixels) bad |= 1;
    __syncthreads();
    const int status = bad;
    if (threadIdx.x == 0) {
        layout[n_pages] = carry;
        layout[n_pages + 1] = status;
        if (layout_mirror) {
            layout_mirror[n_pages] = carry;
            layout_mirror[n_pages + 1] = status;
        }
    }
    for (int page = threadIdx.x; page < n_pages; page += 1024) {
        vkb_planes& pl = planes[page];
        if (status) {  // nothing downstream may touch the (too small) buffers
            meta[page].dst_h = 0;
            meta[page].dst_w = 0;
            pl.dst_h = 0;
            pl.dst_w = 0;
            continue;
        }
        const long long off = layout[page];
        pl.dst_h = meta[page].dst_h;
        pl.dst_w = meta[page].dst_w;
        if (pl.dst_image) pl.dst_image += off * pl.image_channels;
        if (pl.dst_mask) pl.dst_mask += off;
        if (pl.dst_score) pl.dst_score += off;
    }
}

// ============================================================================================
// Cells: inverse (and optionally forward) homography, bbox, tile binning.  Thread per cell.
// ============================================================================================
__global__ void __launch_bounds__(128) grid_cells_kernel(
    const vkb_grid_page* __restrict__ pages, int p_max, int c_max, int t_max,
    const int32_t* __restrict__ lattice_i, vkb_grid_meta* __restrict__ meta,
    double* __restrict__ hinv, double* __restrict__ hfwd, int32_t* __restrict__ cell_box,
    int32_t* __restrict__ tile_count, uint16_t* __restrict__ tile_cells,
    TileSlot* __restrict__ slots, int s_cap) {
    // The 72-byte homographies of a block's 128 consecutive cells are contiguous in global memory:
    // they go through shared memory and leave as coalesced 8-byte stores (one thread storing its own
    // nine doubles touches nine sectors per instruction and throttled the load / store unit).
    __shared__ double s_h[128 * 9];
    const int page = blockIdx.y;
    const vkb_grid_page& pg = pages[page];
    const int ccols = pg.cols - 1;
    const int C = (pg.rows - 1) * ccols;
    const int cell0 = blockIdx.x * blockDim.x;
    if (cell0 >= C) return;  // block uniform
    const int cell = cell0 + threadIdx.x;
    const bool valid = cell < C;
    const int n_valid = min((int)blockDim.x, C - cell0);
    int dx[4] = {0, 0, 0, 0}, dy[4] = {0, 0, 0, 0};
    double sq[8] = {}, dq[8] = {};
    if (valid) {
        const int r = cell / ccols, c = cell - r * ccols;
        const int32_t* lat = lattice_i + (size_t)page * p_max * 2;
        const int i00 = r * pg.cols + c, i01 = i00 + 1, i11 = i00 + pg.cols + 1, i10 = i00 + pg.cols;
        // clockwise: (r,c) (r,c+1) (r+1,c+1) (r+1,c)   (type.py:107-116)
        const int2 p00 = *reinterpret_cast<const int2*>(lat + 2 * i00);
        const int2 p01 = *reinterpret_cast<const int2*>(lat + 2 * i01);
        const int2 p11 = *reinterpret_cast<const int2*>(lat + 2 * i11);
        const int2 p10 = *reinterpret_cast<const int2*>(lat + 2 * i10);
        dx[0] = p00.x; dx[1] = p01.x; dx[2] = p11.x; dx[3] = p10.x;
        dy[0] = p00.y; dy[1] = p01.y; dy[2] = p11.y; dy[3] = p10.y;
        const double sx0 = lattice_coord(c, pg.src_w, pg.grid_size);
        const double sx1 = lattice_coord(c + 1, pg.src_w, pg.grid_size);
        const double sy0 = lattice_coord(r, pg.src_h, pg.grid_size);
        const double sy1 = lattice_coord(r + 1, pg.src_h, pg.grid_size);
        sq[0] = sx0; sq[1] = sy0; sq[2] = sx1; sq[3] = sy0; sq[4] = sx1; sq[5] = sy1; sq[6] = sx0; sq[7] = sy1;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            dq[2 * k] = (double)dx[k];
            dq[2 * k + 1] = (double)dy[k];
        }
        double H[9];
        homography_4pt(dq, sq, H);
#pragma unroll
        for (int i = 0; i < 9; ++i) s_h[threadIdx.x * 9 + i] = H[i];
    }
    __syncthreads();
    {
        double* __restrict__ ho = hinv + ((size_t)page * c_max + cell0) * 9;
        for (int k = threadIdx.x; k < n_valid * 9; k += blockDim.x) ho[k] = s_h[k];
    }
    // the cell's record for the remap (needs the inverse map, still in shared memory here)
    if (valid) {
        const int x0 = min(min(dx[0], dx[1]), min(dx[2], dx[3]));
        const int x1 = max(max(dx[0], dx[1]), max(dx[2], dx[3]));
        const int y0 = min(min(dy[0], dy[1]), min(dy[2], dy[3]));
        const int y1 = max(max(dy[0], dy[1]), max(dy[2], dy[3]));
        const bool big = (x1 - x0 + 32) / 32 != 1 || y1 - y0 + 1 > VKB_CELL_MASK_WORDS;
        const int r = cell / ccols, c = cell - r * ccols;
        const int gs = pg.grid_size;
        double H[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) H[i] = s_h[threadIdx.x * 9 + i];
        TileSlot rec;
        make_cell_local(H, c * gs, r * gs, x0, y0, rec.loc);
        // The float32 form is audited for arguments up to one mask window (tests/hostsim); larger
        // cells and pages whose coordinates overflow the fixed point take the float64 path (NaN).
        if (big || !fast_page_ok(max(pg.src_h, pg.src_w))) rec.loc.a2 = __int_as_float(0x7fc00000);
        const int margin = fast_margin(max(r, c) * gs);
        rec.x0 = x0;
        rec.y0 = y0;
        rec.nr = y1 - y0;
        rec.cellf = cell | (big ? (int)0x80000000 : 0);
        rec.xm = fast_base(c * gs, margin);
        rec.ym = fast_base(r * gs, margin);
        rec.nox = (float)(-x0);
        rec.noy = (float)(-y0);
        int4* __restrict__ dst = reinterpret_cast<int4*>(slots + ((size_t)page * s_cap + cell));
        const int4* src = reinterpret_cast<const int4*>(&rec);
        dst[0] = src[0];
        dst[1] = src[1];
        dst[2] = src[2];
        dst[3] = src[3];
    }
    if (hfwd) {
        __syncthreads();
        if (valid) {
            double H[9];
            homography_4pt(sq, dq, H);
#pragma unroll
            for (int i = 0; i < 9; ++i) s_h[threadIdx.x * 9 + i] = H[i];
        }
        __syncthreads();
        double* __restrict__ hf = hfwd + ((size_t)page * c_max + cell0) * 9;
        for (int k = threadIdx.x; k < n_valid * 9; k += blockDim.x) hf[k] = s_h[k];
    }
    if (!valid) return;
    const int x0 = min(min(dx[0], dx[1]), min(dx[2], dx[3]));
    const int x1 = max(max(dx[0], dx[1]), max(dx[2], dx[3]));
    const int y0 = min(min(dy[0], dy[1]), min(dy[2], dy[3]));
    const int y1 = max(max(dy[0], dy[1]), max(dy[2], dy[3]));
    // bit 30 of x1: the coverage of this cell exceeds the fixed mask budget (one 32-bit word per
    // row, VKB_CELL_MASK_WORDS rows): the masks kernel skips it, the remap rasterises it on the fly
    const bool big = (x1 - x0 + 32) / 32 != 1 || y1 - y0 + 1 > VKB_CELL_MASK_WORDS;
    *reinterpret_cast<int4*>(cell_box + ((size_t)page * c_max + cell) * 4) =
        make_int4(x0, y0, x1 | (big ? 0x40000000 : 0), y1);
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
// Coverage masks (cv.fillPoly of every lattice cell, type.py:199-207).  A warp takes EIGHT cells at
// a time:
//
//   outline   lane = (cell, edge): the lane walks its edge like cv::LineIterator (edge_walk: one
//             add and one compare per pixel) and ORs the pixels into the cell's row words in
//             shared memory; the four scan-edge records of the cell are set up on the way;
//   fill      lane = (cell, row mod 4): crossings of the four scan edges (kept in registers),
//             5-exchange sort, spans, OR with the outline word, one store per row.
//
// ~100 warp instructions per cell; the first generation (warp per cell, lane per row, closed-form
// outline per row) spent ~500 and was the second most expensive kernel of the step.
// ============================================================================================
constexpr int kMaskWarps = 4;
constexpr int kMaskCells = 8;  // cells per warp and step

__global__ void __launch_bounds__(32 * kMaskWarps) grid_masks_kernel(
    const vkb_grid_page* __restrict__ pages, int p_max, int c_max,
    const int32_t* __restrict__ lattice_i, uint32_t* __restrict__ cell_masks) {
    __shared__ uint32_t sm_rows[kMaskWarps][kMaskCells][VKB_CELL_MASK_WORDS];
    __shared__ EdgeScan sm_scan[kMaskWarps][kMaskCells][4];
    const int page = blockIdx.y;
    const vkb_grid_page& pg = pages[page];
    const int ccols = pg.cols - 1;
    const int C = (pg.rows - 1) * ccols;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cell0 = (blockIdx.x * kMaskWarps + warp) * kMaskCells;
    if (cell0 >= C) return;
#pragma unroll
    for (int k = 0; k < kMaskCells; ++k) sm_rows[warp][k][lane] = 0u;

    // ---- lane = (cell, edge): edge e runs from vertex e - 1 to vertex e of the clockwise quad
    //      (r,c) (r,c+1) (r+1,c+1) (r+1,c)   (type.py:107-116)
    const int k8 = lane >> 2, e = lane & 3;
    const int cell = cell0 + k8;
    const bool valid = cell < C;
    int vx = 0, vy = 0;
    if (valid) {
        const int r = floor_div_small(cell, ccols, __fdividef(1.0f, (float)ccols));
        const int c = cell - r * ccols;
        const int idx = (r + (e >> 1)) * pg.cols + c + ((e == 1 || e == 2) ? 1 : 0);
        const int2 v = *reinterpret_cast<const int2*>(lattice_i + ((size_t)page * p_max + idx) * 2);
        vx = v.x;
        vy = v.y;
    }
    const int prev = (lane & ~3) | ((e + 3) & 3);
    const int ux = __shfl_sync(0xffffffffu, vx, prev), uy = __shfl_sync(0xffffffffu, vy, prev);
    int x0 = min(vx, __shfl_xor_sync(0xffffffffu, vx, 1)), x1 = max(vx, __shfl_xor_sync(0xffffffffu, vx, 1));
    int y0 = min(vy, __shfl_xor_sync(0xffffffffu, vy, 1)), y1 = max(vy, __shfl_xor_sync(0xffffffffu, vy, 1));
    x0 = min(x0, __shfl_xor_sync(0xffffffffu, x0, 2));
    x1 = max(x1, __shfl_xor_sync(0xffffffffu, x1, 2));
    y0 = min(y0, __shfl_xor_sync(0xffffffffu, y0, 2));
    y1 = max(y1, __shfl_xor_sync(0xffffffffu, y1, 2));
    int nrows = y1 - y0 + 1;
    // too large for the fixed budget: flagged by grid_cells_kernel, rasterised by the remap
    const bool big = (x1 - x0 + 32) / 32 != 1 || nrows > VKB_CELL_MASK_WORDS;
    if (!valid || big) nrows = 0;
    if (nrows) {
        EdgeScan es;
        edge_scan_setup(ux, uy, vx, vy, es);
        sm_scan[warp][k8][e] = es;
    }
    __syncwarp();  // the zeroed rows are visible before the first OR
    if (nrows) {
        uint32_t* rows = sm_rows[warp][k8];
        edge_walk(ux, uy, vx, vy, [&](int x, int y) { atomicOr(rows + (y - y0), 1u << (x - x0)); });
    }
    __syncwarp();

    // ---- lane = (cell, row mod 4): the four lanes of a cell take its rows in turn; the cell's scan
    //      edges stay in registers for all of them (a linearised (cell, row) list balanced the lanes
    //      better but paid 17 % of the kernel's instructions for finding its cell per row)
    if (nrows) {
        EdgeScan es4[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) es4[i] = sm_scan[warp][k8][i];
        const uint32_t* rows = sm_rows[warp][k8];
        uint32_t* out = cell_masks + ((size_t)page * c_max + cell) * VKB_CELL_MASK_WORDS;
        for (int row = e; row < nrows; row += 4) {
            // A row of an ordinary (convex) cell meets exactly two scan edges: one span between
            // them, no sorting.  Other counts (folded cells) take the general routine.
            const int y = y0 + row;
            int lo = kNoCross, hi = -kNoCross, n_cross = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const bool active = y >= es4[i].ya && y < es4[i].yb;
                const int c = es4[i].base + es4[i].dxf * (y - es4[i].ya);
                lo = active ? min(lo, c) : lo;
                hi = active ? max(hi, c) : hi;
                n_cross += active ? 1 : 0;
            }
            uint32_t fill;
            if (n_cross == 2) {
                const int xl = (int)(((long long)lo + 65535) >> 16), xr = hi >> 16;
                fill = xl <= xr ? bits_lo_hi(xl - x0, xr - x0) : 0u;
            } else {
                fill = quad_fill_row(es4, y, x0);
            }
            out[row] = rows[row] | fill;
        }
    }
}

// ============================================================================================
// Per-tile candidate lists for the remap kernel.
//
//   grid_tile_base_kernel     prefix sum over pages of the number of 32 x 32 dst tiles: the
//                             remap's flat work list (tile_base[n_pages] = all tiles);
//   grid_tile_lists_kernel    eight lanes per tile: the tile's candidates in ascending cell order
//                             and its header.  (The 64-byte record of a candidate -- bbox and the
//                             float32 form of the cell's inverse homography, CellLocal -- exists
//                             once per CELL, written by grid_cells_kernel and re-centred on the
//                             cell's bbox origin; round 2's first version wrote one per
//                             (tile, cell) pair, re-centred on the tile: 2.2 x the float64 work
//                             and 280 MB per 256 pages, 119 us.)
// ============================================================================================
// exclusive scan of v over the block (1024 threads); returns the exclusive prefix, adds the
// block total to `carry` (shared).
__device__ __forceinline__ int block_exclusive_scan_1024(int v, int* warp_sums, int* carry) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += o;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int ws = warp_sums[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, ws, d);
            if (lane >= d) ws += o;
        }
        warp_sums[lane] = ws;
    }
    __syncthreads();
    const int base = *carry + (warp > 0 ? warp_sums[warp - 1] : 0);
    const int total = warp_sums[31];
    __syncthreads();
    if (threadIdx.x == 0) *carry += total;
    __syncthreads();
    return base + inc - v;
}

__global__ void __launch_bounds__(1024) grid_tile_base_kernel(const vkb_grid_meta* __restrict__ meta,
                                                              int n_pages, int32_t* __restrict__ tile_base) {
    __shared__ int warp_sums[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n_pages; base += 1024) {
        const int page = base + threadIdx.x;
        const int v = page < n_pages ? page_tiles(meta[page]) : 0;
        const int ex = block_exclusive_scan_1024(v, warp_sums, &carry);
        if (page < n_pages) tile_base[page] = ex;
    }
    if (threadIdx.x == 0) tile_base[n_pages] = carry;
}

// Eight lanes per tile: rank the tile's candidate cells by cell index (the remap resolves "last
// writer wins" by letting later candidates overwrite earlier ones), write the sorted list behind
// the page's cell records and the tile's header: page, origin, count, the acceptance limit of
// the fast path and the first 16 sorted candidates packed for the small-tile kernel.  The kernel
// is a chain of dependent loads per tile (count -> bin -> stores); what matters is how many tiles
// are in flight, hence the narrow groups.
#ifndef VKB_LIST_LANES
#define VKB_LIST_LANES 8
#endif
constexpr int kListLanes = VKB_LIST_LANES;
constexpr int kListTiles = 128 / kListLanes;  // tiles per block

__global__ void __launch_bounds__(128) grid_tile_lists_kernel(
    const vkb_grid_page* __restrict__ pages, const vkb_grid_meta* __restrict__ meta, int c_max,
    int t_max, int s_cap, const int32_t* __restrict__ tile_count,
    const uint16_t* __restrict__ tile_cells, const int32_t* __restrict__ tile_base,
    TileSlot* __restrict__ slots, RemapTile* __restrict__ headers, int32_t* __restrict__ large) {
    __shared__ __align__(4) uint16_t s_sorted[kListTiles][VKB_TILE_CAP];
    __shared__ __align__(4) uint16_t s_raw[kListTiles][VKB_TILE_CAP];
    const int page = blockIdx.y;
    const int group = threadIdx.x / kListLanes;
    const int t = blockIdx.x * kListTiles + group;
    const int lane = threadIdx.x % kListLanes;
    const int dst_w = meta[page].dst_w;
    const int tiles_x = (dst_w + VKB_TILE - 1) / VKB_TILE;
    if (t >= min(page_tiles(meta[page]), t_max)) return;
    const size_t pt = (size_t)page * t_max + t;
    const int count = tile_count[pt];
    const int ty = t / tiles_x, tx = t - ty * tiles_x;
    const bool usable = count <= VKB_TILE_CAP;
    const unsigned group_mask = ((1u << kListLanes) - 1u) << (threadIdx.x & 31 & ~(kListLanes - 1));
    const uint16_t* __restrict__ cells = tile_cells + pt * VKB_TILE_CAP;
    uint16_t* sorted = s_sorted[group];
    uint16_t* raw = s_raw[group];
    int reach = 0;
    if (usable) {
        // the bin goes through shared memory (one coalesced read); ranks come from broadcast reads
        for (int s = lane; s < count; s += kListLanes) raw[s] = cells[s];
        __syncwarp(group_mask);
        const int ccols = pages[page].cols - 1;
        for (int s = lane; s < count; s += kListLanes) {
            const int cell = (int)raw[s];
            int rank = 0;
            for (int j = 0; j < count; ++j) rank += (int)raw[j] < cell;
            sorted[rank] = (uint16_t)cell;
            const int r = cell / ccols, c = cell - r * ccols;
            reach = max(reach, max(r, c));
        }
#pragma unroll
        for (int d = kListLanes / 2; d >= 1; d >>= 1) reach = max(reach, __shfl_xor_sync(group_mask, reach, d));
    }
    __syncwarp(group_mask);
    RemapTile* __restrict__ hd = headers + (tile_base[page] + t);
    if (usable) {
        // the sorted list: behind the page's cell records (tile_list), two 16-bit entries per lane and step
        uint32_t* __restrict__ out = reinterpret_cast<uint32_t*>(
            const_cast<uint16_t*>(tile_list(slots, page, s_cap, c_max, t)));
        const uint32_t* sw = reinterpret_cast<const uint32_t*>(sorted);
        for (int w = lane; 2 * w < count; w += kListLanes) out[w] = sw[w];  // (an odd count copies one stale entry)
        for (int w = lane; w < 8; w += kListLanes) {
            const uint32_t lo = w < count ? sorted[w] : 0u;
            const uint32_t hi = w + 8 < count ? sorted[w + 8] : 0u;
            hd->ids[w] = lo | (hi << 16);
        }
    }
    if (lane == 0) {
        RemapTile h;
        h.page = page;
        h.tx0 = tx * VKB_TILE;
        h.ty0 = ty * VKB_TILE;
        h.count = usable ? count : -1;
        h.tile = t;
        h.lim = fast_limit(fast_margin(reach * pages[page].grid_size));
        h.pad[0] = h.pad[1] = 0;
        int4* __restrict__ dst = reinterpret_cast<int4*>(hd);
        const int4* src = reinterpret_cast<const int4*>(&h);
        dst[0] = src[0];
        dst[1] = src[1];
        // tiles outside the lane-per-row owner path go on the second launch's work list
        if ((unsigned)h.count > (unsigned)kPlaneCands) large[1 + atomicAdd(large, 1)] = tile_base[page] + t;
    }
}

// ============================================================================================
// Affine / perspective warp (rotate / shear / skew: affine.py:38-43, 416-456).  Block (32, 8)
// covers a 32 x 32 dst tile, 4 rows per thread.  The coordinates follow cv::warpAffine /
// cv::warpPerspective (double arithmetic, 10-bit / 5-bit fixed point); the gather is the remap's:
// all four pixels' aligned row loads in flight, PRMT + IDP.4A blend for RGB, tap weights shared by
// Image and Mask, out-of-image footprints fixed behind one branch per thread.
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
    constexpr int R = 4;
    const int src_h = pl.src_h, src_w = pl.src_w;

    int X[R], Y[R];
    if (pg.kind == VKB_WARP_AFFINE) {
        const int adelta = cv_round_d(__dmul_rn(__dmul_rn(pg.inv[0], (double)x), 1024.0));
        const int bdelta = cv_round_d(__dmul_rn(__dmul_rn(pg.inv[3], (double)x), 1024.0));
#pragma unroll
        for (int k = 0; k < R; ++k) {
            const int y = y_base + 8 * k;
            const int X0 = cv_round_d(__dmul_rn(__dadd_rn(__dmul_rn(pg.inv[1], (double)y), pg.inv[2]), 1024.0)) + 16;
            const int Y0 = cv_round_d(__dmul_rn(__dadd_rn(__dmul_rn(pg.inv[4], (double)y), pg.inv[5]), 1024.0)) + 16;
            X[k] = (X0 + adelta) >> 5;
            Y[k] = (Y0 + bdelta) >> 5;
        }
    } else {
#pragma unroll
        for (int k = 0; k < R; ++k) perspective_coord(pg.inv, x, y_base + 8 * k, X[k], Y[k]);
    }

    const bool tiny = src_h < 2 || src_w < 2;
    TapWeights tw[R];
    bool outside = false;
#pragma unroll
    for (int k = 0; k < R; ++k) {
        tw[k] = tap_weights_plain(X[k], Y[k]);
        outside |= tap_outside(tw[k], src_h, src_w);
    }
    if (outside && !tiny) {
#pragma unroll
        for (int k = 0; k < R; ++k)
            if (tap_outside(tw[k], src_h, src_w)) tap_border_fix(tw[k], src_h, src_w);
    }
    const int channels = pl.image_channels;
    if (channels == 3 && !tiny) {
        Taps<3> taps[R];
#pragma unroll
        for (int k = 0; k < R; ++k) {
            taps[k].t = tw[k];
            taps_load<3>(pl.src_image, src_w, taps[k]);
        }
#pragma unroll
        for (int k = 0; k < R; ++k) {
            const int y = y_base + 8 * k;
            uint8_t px[3];
            taps_blend<3>(taps[k], px);
            if (y < pl.dst_h) {
                uint8_t* d = pl.dst_image + ((long long)y * pl.dst_w + x) * 3;
                d[0] = px[0];
                d[1] = px[1];
                d[2] = px[2];
            }
        }
    } else if (channels) {
#pragma unroll
        for (int k = 0; k < R; ++k) {
            const int y = y_base + 8 * k;
            if (y >= pl.dst_h) break;
            const long long dst_idx = (long long)y * pl.dst_w + x;
            if (channels == 3) {
                uint8_t px[3];
                bilinear_u8<3>(pl.src_image, src_h, src_w, (long long)src_w * 3, X[k], Y[k], px);
                uint8_t* d = pl.dst_image + dst_idx * 3;
                d[0] = px[0];
                d[1] = px[1];
                d[2] = px[2];
            } else if (channels == 1) {
                uint8_t px[1];
                bilinear_u8<1>(pl.src_image, src_h, src_w, (long long)src_w, X[k], Y[k], px);
                pl.dst_image[dst_idx] = px[0];
            } else {
                uint8_t px[4];
                bilinear_u8<4>(pl.src_image, src_h, src_w, (long long)src_w * 4, X[k], Y[k], px);
                *reinterpret_cast<uchar4*>(pl.dst_image + dst_idx * 4) =
                    make_uchar4(px[0], px[1], px[2], px[3]);
            }
        }
    }
    if (pl.src_mask) {
        if (!tiny) {
            Taps<1> taps[R];
#pragma unroll
            for (int k = 0; k < R; ++k) {
                taps[k].t = tw[k];
                taps_load<1>(pl.src_mask, src_w, taps[k]);
            }
#pragma unroll
            for (int k = 0; k < R; ++k) {
                const int y = y_base + 8 * k;
                uint8_t m[1];
                taps_blend<1>(taps[k], m);
                if (y < pl.dst_h) pl.dst_mask[(long long)y * pl.dst_w + x] = m[0];
            }
        } else {
#pragma unroll
            for (int k = 0; k < R; ++k) {
                const int y = y_base + 8 * k;
                if (y >= pl.dst_h) break;
                uint8_t m[1];
                bilinear_u8<1>(pl.src_mask, src_h, src_w, (long long)src_w, X[k], Y[k], m);
                pl.dst_mask[(long long)y * pl.dst_w + x] = m[0];
            }
        }
    }
    if (pl.src_score) {
#pragma unroll
        for (int k = 0; k < R; ++k) {
            const int y = y_base + 8 * k;
            if (y >= pl.dst_h) break;
            pl.dst_score[(long long)y * pl.dst_w + x] =
                bilinear_f32(pl.src_score, src_h, src_w, src_w, X[k], Y[k]);
        }
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

// Batched forms for RandomDistortionBatch: every point carries its page.
__global__ void grid_points_batched_kernel(const vkb_grid_page* __restrict__ pages,
                                           const double* __restrict__ hfwd, int c_max,
                                           const double* __restrict__ xy_in,
                                           const int32_t* __restrict__ page_cell,
                                           double* __restrict__ xy_out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int page = page_cell[3 * i];
    const int ccols = pages[page].cols - 1;
    const double* H = hfwd + ((size_t)page * c_max + (size_t)page_cell[3 * i + 1] * ccols
                              + page_cell[3 * i + 2]) * 9;
    const double x = xy_in[2 * i], y = xy_in[2 * i + 1];
    const double tx = __dadd_rn(__dadd_rn(__dmul_rn(H[0], x), __dmul_rn(H[1], y)), H[2]);
    const double ty = __dadd_rn(__dadd_rn(__dmul_rn(H[3], x), __dmul_rn(H[4], y)), H[5]);
    const double t = __dadd_rn(__dadd_rn(__dmul_rn(H[6], x), __dmul_rn(H[7], y)), H[8]);
    xy_out[2 * i] = __ddiv_rn(tx, t);
    xy_out[2 * i + 1] = __ddiv_rn(ty, t);
}

// mats: per page 9 doubles (row major, 2 x 3 matrices in the first six) + kind in `rows_f32`
// (bits 0-1: rows 2 / 3, bit 2: float32 arithmetic like affine_np_points with a 2 x 3 matrix)
__global__ void affine_points_batched_kernel(const double* __restrict__ mats,
                                             const int32_t* __restrict__ rows_f32,
                                             const int32_t* __restrict__ page_of,
                                             const double* __restrict__ xy_in,
                                             double* __restrict__ xy_out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int page = page_of[i];
    const double* M = mats + (size_t)page * 9;
    const int rows = rows_f32[page] & 3;
    if (rows_f32[page] & 4) {
        const float x = (float)xy_in[2 * i], y = (float)xy_in[2 * i + 1];
        xy_out[2 * i] = (double)__fmaf_rn((float)M[2], 1.0f, __fmaf_rn((float)M[1], y, __fmul_rn((float)M[0], x)));
        xy_out[2 * i + 1] = (double)__fmaf_rn((float)M[5], 1.0f, __fmaf_rn((float)M[4], y, __fmul_rn((float)M[3], x)));
        return;
    }
    const double x = xy_in[2 * i], y = xy_in[2 * i + 1];
    const double tx = fma(M[2], 1.0, fma(M[1], y, __dmul_rn(M[0], x)));
    const double ty = fma(M[5], 1.0, fma(M[4], y, __dmul_rn(M[3], x)));
    if (rows == 2) {
        xy_out[2 * i] = tx;
        xy_out[2 * i + 1] = ty;
        return;
    }
    const double t = fma(M[8], 1.0, fma(M[7], y, __dmul_rn(M[6], x)));
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

// ============================================================================================
// Ordered polygon fills (label rasterisation after distortion): N polygons, each cv.fillPoly
// exact (scan fill with 16.16 edges + 8-connected outline), applied in list order like the
// reference's loops of Polygon.fill_mask / fill_score_map (page_distortion.py:163-314,
// engine/char_mask/default.py:44-56).  Pass 1 marks every covered pixel with atomicMax of an
// order key into an int32 canvas (assign: the polygon's list position, so the LAST polygon wins;
// keep max / keep min: an order-preserving image of the value); pass 2 resolves the keys.
// ============================================================================================
__device__ __forceinline__ void poly_mark(int32_t* __restrict__ keys, int h, int w, int x, int y, int key) {
    if ((unsigned)x < (unsigned)w && (unsigned)y < (unsigned)h) atomicMax(keys + (size_t)y * w + x, key);
}

__device__ __forceinline__ int poly_item_key(const vkb_poly_item& it, int index, int mode) {
    if (mode == 0) return index + 1;
    const int bits = __float_as_int(it.value);  // values are >= 0: the bit pattern is monotonic
    return mode == 1 ? bits + 1 : 0x7fffffff - bits;
}

__global__ void __launch_bounds__(128) poly_rows_kernel(int32_t* __restrict__ keys, int h, int w,
                                                        const int32_t* __restrict__ pts,
                                                        const vkb_poly_item* __restrict__ items,
                                                        int mode) {
    const vkb_poly_item it = items[blockIdx.y];
    const int y = it.y_min + blockIdx.x * blockDim.x + threadIdx.x;
    if (y > it.y_max || y < 0 || y >= h) return;
    const int32_t* __restrict__ poly = pts + 2 * (size_t)it.first_pt;
    const int n = it.n_pts;
    long long cross[kMaxCross];
    int nc = 0;
    for (int i = 0; i < n; ++i) {
        const int j = i == 0 ? n - 1 : i - 1;
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
    const int key = poly_item_key(it, blockIdx.y, mode);
    int32_t* __restrict__ row = keys + (size_t)y * w;
    for (int k = 0; k + 1 < nc; k += 2) {
        long long xl = (cross[k] + 65535) >> 16;
        long long xr = cross[k + 1] >> 16;
        if (xl < 0) xl = 0;
        if (xr > w - 1) xr = w - 1;
        for (long long xx = xl; xx <= xr; ++xx) atomicMax(row + xx, key);
    }
}

__global__ void __launch_bounds__(128) poly_edges_kernel(int32_t* __restrict__ keys, int h, int w,
                                                         const int32_t* __restrict__ pts,
                                                         const vkb_poly_item* __restrict__ items,
                                                         int mode) {
    const vkb_poly_item it = items[blockIdx.y];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = it.n_pts;
    if (i >= n) return;
    const int32_t* __restrict__ poly = pts + 2 * (size_t)it.first_pt;
    const int j = i == 0 ? n - 1 : i - 1;
    int ax = poly[2 * j], ay = poly[2 * j + 1], bx = poly[2 * i], by = poly[2 * i + 1];
    if (ax > bx) { int t = ax; ax = bx; bx = t; t = ay; ay = by; by = t; }
    int dx = bx - ax, dy = by - ay;
    const int sy = dy >= 0 ? 1 : -1;
    dy = dy >= 0 ? dy : -dy;
    const bool steep = dy > dx;
    if (steep) { const int t = dx; dx = dy; dy = t; }
    int err = dx - 2 * dy;
    int x = ax, y = ay;
    const int key = poly_item_key(it, blockIdx.y, mode);
    for (int s = 0; s <= dx; ++s) {
        poly_mark(keys, h, w, x, y, key);
        const bool m = err < 0;
        err += -2 * dy + (m ? 2 * dx : 0);
        if (steep) { y += sy; if (m) x += 1; }
        else { x += 1; if (m) y += sy; }
    }
}

template <typename T>
__global__ void __launch_bounds__(256) poly_resolve_kernel(T* __restrict__ dst, const int32_t* __restrict__ keys,
                                                           long long n, const vkb_poly_item* __restrict__ items,
                                                           int mode) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int key = keys[i];
    if (key <= 0) return;
    float v;
    if (mode == 0) v = items[key - 1].value;
    else if (mode == 1) v = __int_as_float(key - 1);
    else v = __int_as_float(0x7fffffff - key);
    const float old = (float)dst[i];
    if (mode == 1) v = fmaxf(old, v);
    if (mode == 2) v = fminf(old, v);
    dst[i] = (T)v;
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
    VKB_NVTX("vkb_warp_fused");
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
                                double* lattice_f, int32_t projectors, void* stream) {
    VKB_NVTX("vkb_grid_project");
    VKB_REQUIRE(pages && lattice_f && n_pages > 0 && p_max > 0, "bad arguments");
    VKB_REQUIRE(n_pages <= 65535, "at most 65535 pages per launch");
    // every kernel skips the pages of the other projector; a kernel no page needs is not launched
    if (projectors & (1 << VKB_PROJ_CAMERA)) {
        grid_project_camera_kernel<<<n_pages, 1024, 0, (cudaStream_t)stream>>>(pages, p_max, lattice_f);
        int rc = check_launch("grid_project_camera_kernel");
        if (rc) return rc;
    }
    if (projectors & (1 << VKB_PROJ_MLS)) {
        // VKB_MLS_V1=1 keeps the warp-per-point generation (A/B tests)
        const char* v1 = getenv("VKB_MLS_V1");
        if (v1 && v1[0] == '1') {
            grid_project_mls_kernel<<<dim3((p_max + 7) / 8, n_pages), 256, 0, (cudaStream_t)stream>>>(
                pages, p_max, lattice_f);
            return check_launch("grid_project_mls_kernel");
        }
        grid_project_mls_points_kernel<<<dim3((p_max + 127) / 128, n_pages), 128, 0,
                                         (cudaStream_t)stream>>>(pages, p_max, lattice_f);
        return check_launch("grid_project_mls_points_kernel");
    }
    return VKB_OK;
}

extern "C" int vkb_grid_finalize(const vkb_grid_page* pages, int32_t n_pages, int32_t p_max,
                                 const double* lattice_f, int32_t* lattice_i, vkb_grid_meta* meta,
                                 vkb_grid_meta* meta_mirror, void* stream) {
    VKB_NVTX("vkb_grid_finalize");
    VKB_REQUIRE(pages && lattice_f && lattice_i && meta && n_pages > 0, "bad arguments");
    grid_finalize_kernel<<<n_pages, 1024, 0, (cudaStream_t)stream>>>(pages, p_max, lattice_f,
                                                                    lattice_i, meta, meta_mirror);
    return check_launch("grid_finalize_kernel");
}

// Parameter blocks from (mapped, pinned) host memory into device memory by a KERNEL: a
// cudaMemcpyAsync of a few kilobytes would wait on the copy engine behind whatever bulk page
// copies the caller has queued there (the end-to-end pipeline keeps ~100 MB in flight).
__global__ void __launch_bounds__(256) stage_params_kernel(uint4* __restrict__ dst,
                                                           const uint4* __restrict__ src, int n16) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += gridDim.x * blockDim.x)
        dst[i] = src[i];
}

extern "C" int vkb_stage_params(void* dst, const void* src_host, int64_t nbytes, void* stream) {
    VKB_REQUIRE(dst && src_host && nbytes > 0 && nbytes % 16 == 0, "bad arguments (16-byte units)");
    VKB_REQUIRE(nbytes < (1ll << 30), "parameter blocks only");
    const int n16 = (int)(nbytes / 16);
    const int blocks = n16 < 256 * 64 ? (n16 + 255) / 256 : 64;
    stage_params_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<uint4*>(dst), reinterpret_cast<const uint4*>(src_host), n16);
    return check_launch("stage_params_kernel");
}

extern "C" int vkb_grid_layout(vkb_grid_meta* meta, int32_t n_pages, vkb_planes* planes,
                               int64_t cap_pixels, int32_t t_max, int64_t* layout,
                               int64_t* layout_mirror, void* stream) {
    VKB_NVTX("vkb_grid_layout");
    VKB_REQUIRE(meta && planes && layout && n_pages > 0, "bad arguments");
    VKB_REQUIRE(cap_pixels > 0 && t_max > 0, "empty capacities");
    grid_layout_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(
        meta, n_pages, planes, (long long)cap_pixels, t_max, reinterpret_cast<long long*>(layout),
        reinterpret_cast<long long*>(layout_mirror));
    return check_launch("grid_layout_kernel");
}

// One side stream + fork / join events per device for vkb_grid_build (created on first use; work
// submitted through it is ordered with the caller's stream by the two events).
struct BuildSide {
    cudaStream_t stream;
    cudaEvent_t fork, join;
};
static BuildSide* build_side() {
    static BuildSide sides[64];
    static bool ready[64] = {};
    static std::mutex mutex;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> lock(mutex);
    if (!ready[dev]) {
        BuildSide s;
        if (cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        sides[dev] = s;
        ready[dev] = true;
    }
    return &sides[dev];
}

extern "C" int vkb_grid_build(const vkb_grid_page* pages, int32_t n_pages, int32_t p_max,
                              int32_t c_max, int32_t t_max, int32_t s_cap, const int32_t* lattice_i,
                              vkb_grid_meta* meta, double* hinv, double* hfwd, int32_t* cell_box,
                              uint32_t* cell_masks, int32_t* tile_count, uint16_t* tile_cells,
                              int32_t* tile_off, int32_t* tile_base, void* tile_slots,
                              void* tile_headers, void* stream) {
    VKB_NVTX("vkb_grid_build");
    VKB_REQUIRE(pages && lattice_i && meta && hinv && cell_box && cell_masks && tile_count
                    && tile_cells && tile_base && tile_slots && tile_headers,
                "bad arguments");
    VKB_REQUIRE(n_pages > 0 && n_pages <= 65535, "1..65535 pages per launch");
    VKB_REQUIRE(c_max > 0 && c_max <= 65535, "at most 65535 cells per page");
    VKB_REQUIRE(t_max > 0 && s_cap > 0, "empty tile workspace");
    VKB_REQUIRE((long long)n_pages * t_max < (1ll << 31), "too many tiles in one launch");
    VKB_REQUIRE((long long)n_pages * s_cap < (1ll << 31), "too many tile records in one launch");
    VKB_REQUIRE(s_cap >= c_max + 2 * t_max, "s_cap holds the cell records and the tile lists: >= c_max + 2 * t_max");
    (void)tile_off;  // (kept in the signature: earlier layouts kept per-tile record offsets here)
    cudaStream_t st = (cudaStream_t)stream;
    VKB_CUDA(cudaMemsetAsync(tile_count, 0, sizeof(int32_t) * (size_t)n_pages * t_max, st));
    // the list of large tiles lives behind the headers: [count, tile index ...]
    int32_t* large = reinterpret_cast<int32_t*>(reinterpret_cast<char*>(tile_headers)
                                                + (size_t)n_pages * t_max * VKB_TILE_HEADER_BYTES);
    VKB_CUDA(cudaMemsetAsync(large, 0, sizeof(int32_t), st));
    // cells (fp64 latency bound) and masks (integer issue bound) only share their input: the
    // masks run on a side stream next to the cells
    BuildSide* side = build_side();
    static std::mutex side_mutex;  // the fork / join events are shared by all callers
    std::lock_guard<std::mutex> side_lock(side_mutex);
    if (side) {
        VKB_CUDA(cudaEventRecord(side->fork, st));
        VKB_CUDA(cudaStreamWaitEvent(side->stream, side->fork, 0));
    }
    grid_masks_kernel<<<dim3((c_max + kMaskWarps * kMaskCells - 1) / (kMaskWarps * kMaskCells), n_pages),
                        32 * kMaskWarps, 0, side ? side->stream : st>>>(pages, p_max, c_max, lattice_i,
                                                                       cell_masks);
    int rc = check_launch("grid_masks_kernel");
    if (rc) return rc;
    grid_cells_kernel<<<dim3((c_max + 127) / 128, n_pages), 128, 0, st>>>(
        pages, p_max, c_max, t_max, lattice_i, meta, hinv, hfwd, cell_box, tile_count, tile_cells,
        reinterpret_cast<TileSlot*>(tile_slots), s_cap);
    rc = check_launch("grid_cells_kernel");
    if (rc) return rc;
    // (the caller's stream joins the masks after the tile lists: only the remap reads them)
    if (side) VKB_CUDA(cudaEventRecord(side->join, side->stream));
    grid_tile_base_kernel<<<1, 1024, 0, st>>>(meta, n_pages, tile_base);
    grid_tile_lists_kernel<<<dim3((t_max + kListTiles - 1) / kListTiles, n_pages), 128, 0, st>>>(
        pages, meta, c_max, t_max, s_cap, tile_count, tile_cells, tile_base,
        reinterpret_cast<TileSlot*>(tile_slots), reinterpret_cast<RemapTile*>(tile_headers), large);
    rc = check_launch("grid_tile_lists_kernel");
    if (rc) return rc;
    if (side) VKB_CUDA(cudaStreamWaitEvent(st, side->join, 0));
    return VKB_OK;
}

extern "C" int vkb_grid_points(const double* hfwd_page, int32_t cols_minus_1, const double* xy_in,
                               const int32_t* cell_rc, double* xy_out, int32_t n, void* stream) {
    VKB_REQUIRE(hfwd_page && xy_in && cell_rc && xy_out, "bad arguments");
    if (n <= 0) return VKB_OK;
    grid_points_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(hfwd_page, cols_minus_1,
                                                                          xy_in, cell_rc, xy_out, n);
    return check_launch("grid_points_kernel");
}

extern "C" int vkb_grid_points_batched(const vkb_grid_page* pages, const double* hfwd, int32_t c_max,
                                       const double* xy_in, const int32_t* page_cell,
                                       double* xy_out, int32_t n, void* stream) {
    VKB_NVTX("vkb_grid_points_batched");
    VKB_REQUIRE(pages && hfwd && xy_in && page_cell && xy_out && c_max > 0, "bad arguments");
    if (n <= 0) return VKB_OK;
    grid_points_batched_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        pages, hfwd, c_max, xy_in, page_cell, xy_out, n);
    return check_launch("grid_points_batched_kernel");
}

extern "C" int vkb_affine_points_batched(const double* mats, const int32_t* rows_f32,
                                         const int32_t* page_of, const double* xy_in,
                                         double* xy_out, int32_t n, void* stream) {
    VKB_NVTX("vkb_affine_points_batched");
    VKB_REQUIRE(mats && rows_f32 && page_of && xy_in && xy_out, "bad arguments");
    if (n <= 0) return VKB_OK;
    affine_points_batched_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        mats, rows_f32, page_of, xy_in, xy_out, n);
    return check_launch("affine_points_batched_kernel");
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

extern "C" int vkb_fill_polygons(void* dst, int32_t dst_f32, int32_t h, int32_t w,
                                 const int32_t* pts_xy, const vkb_poly_item* items,
                                 const vkb_poly_item* items_host, int32_t n_items, int32_t mode,
                                 int32_t* keys, void* stream) {
    VKB_NVTX("vkb_fill_polygons");
    VKB_REQUIRE(dst && pts_xy && items && items_host && keys && h > 0 && w > 0, "bad arguments");
    VKB_REQUIRE(mode >= 0 && mode <= 2, "mode must be 0 (assign), 1 (keep max) or 2 (keep min)");
    VKB_REQUIRE(n_items >= 0 && n_items <= 65535, "at most 65535 polygons per launch");
    if (n_items == 0) return VKB_OK;
    int max_rows = 1, max_pts = 1;
    for (int i = 0; i < n_items; ++i) {
        const vkb_poly_item& it = items_host[i];
        VKB_REQUIRE(it.n_pts >= 1 && it.first_pt >= 0, "empty polygon");
        VKB_REQUIRE(mode == 0 || it.value >= 0.f, "keep max / keep min need values >= 0");
        VKB_REQUIRE(dst_f32 || (it.value >= 0.f && it.value <= 255.f && it.value == (float)(int)it.value),
                    "uint8 destinations take integer values 0..255");
        const int rows = it.y_max - it.y_min + 1;
        max_rows = rows > max_rows ? rows : max_rows;
        max_pts = it.n_pts > max_pts ? it.n_pts : max_pts;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const long long n = (long long)h * w;
    VKB_CUDA(cudaMemsetAsync(keys, 0, sizeof(int32_t) * (size_t)n, st));
    poly_rows_kernel<<<dim3((max_rows + 127) / 128, n_items), 128, 0, st>>>(keys, h, w, pts_xy, items, mode);
    int rc = check_launch("poly_rows_kernel");
    if (rc) return rc;
    poly_edges_kernel<<<dim3((max_pts + 127) / 128, n_items), 128, 0, st>>>(keys, h, w, pts_xy, items, mode);
    rc = check_launch("poly_edges_kernel");
    if (rc) return rc;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    if (dst_f32)
        poly_resolve_kernel<float><<<blocks, 256, 0, st>>>(reinterpret_cast<float*>(dst), keys, n, items, mode);
    else
        poly_resolve_kernel<uint8_t><<<blocks, 256, 0, st>>>(reinterpret_cast<uint8_t*>(dst), keys, n, items, mode);
    return check_launch("poly_resolve_kernel");
}
