// fog.cu -- the plasma field of the `fog` distortion on the device
// (vkit/mechanism/distortion/photometric/effect.py:89-208: generate_diamond_square_mask + the
// normalisation of fog_image).
//
// The field is a function of the caller's NumPy generator, so it has to consume that generator's
// stream: NumPy's default bit generator is PCG64 (a 128-bit LCG with an XSL-RR output), whose
// state can be advanced by any number of steps in O(log n).  Every uniform double of the field is
// ONE 64-bit output ((out >> 11) * 2^-53), so draw number i is a pure function of (state, i): the
// kernel below regenerates the reference's draws in parallel, the host only advances its
// generator by the number of draws.  The diamond-square arithmetic follows NumPy's dtype rules
// of the reference's expressions operation by operation (float32 corner sums, float32 products
// with the weak Python scalar for the square centres, float64 everywhere a float64 array takes
// part; no contraction), see the kernels.
#include <stdint.h>
#include "common.cuh"

namespace vkb {

typedef unsigned __int128 u128;

__device__ __forceinline__ u128 make_u128(uint64_t hi, uint64_t lo) { return ((u128)hi << 64) | lo; }

// PCG64's multiplier (pcg_variants.h: PCG_DEFAULT_MULTIPLIER_128)
__device__ __forceinline__ u128 pcg_mult() { return make_u128(0x2360ED051FC65DA4ull, 0x4385DF649FCCF645ull); }

// state after `delta` steps of  s -> s * M + inc  (Brown, "Random number generation with arbitrary strides")
__device__ u128 pcg_advance(u128 state, u128 inc, uint64_t delta) {
    u128 acc_mult = 1, acc_plus = 0, cur_mult = pcg_mult(), cur_plus = inc;
    while (delta > 0) {
        if (delta & 1) {
            acc_mult *= cur_mult;
            acc_plus = acc_plus * cur_mult + cur_plus;
        }
        cur_plus = (cur_mult + 1) * cur_plus;
        cur_mult *= cur_mult;
        delta >>= 1;
    }
    return acc_mult * state + acc_plus;
}

// XSL-RR 128/64 output of a state, as a double in [0, 1) (numpy: next_double)
__device__ __forceinline__ double pcg_double(u128 s) {
    const uint64_t hi = (uint64_t)(s >> 64), lo = (uint64_t)s;
    const uint64_t x = hi ^ lo;
    const unsigned r = (unsigned)(hi >> 58);
    const uint64_t o = (x >> r) | (x << ((64u - r) & 63u));
    return (double)(o >> 11) * (1.0 / 9007199254740992.0);
}

constexpr int kDrawsPerThread = 16;

// draws[i] = the (i + 1)-th double of the generator whose state is (state, inc)
__global__ void __launch_bounds__(128) pcg64_uniform_kernel(double* __restrict__ draws, int64_t n,
                                                            uint64_t state_hi, uint64_t state_lo,
                                                            uint64_t inc_hi, uint64_t inc_lo) {
    const int64_t first = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * kDrawsPerThread;
    if (first >= n) return;
    const u128 inc = make_u128(inc_hi, inc_lo), mult = pcg_mult();
    u128 s = pcg_advance(make_u128(state_hi, state_lo), inc, (uint64_t)first);
    const int64_t last = first + kDrawsPerThread < n ? first + kDrawsPerThread : n;
    for (int64_t i = first; i < last; ++i) {
        s = s * mult + inc;
        draws[i] = pcg_double(s);
    }
}

// One level of the diamond-square construction.  F: size x size float32 field; the level's
// corners are F[i * step][j * step], i, j = 0 .. m.
//
//   centres (m x m):   keep32 * ((c[i][j] + c[i+1][j]) + (c[i][j] + c[i][j+1])) / 4   in float32
//                      (the reference's `down[:, :-1] + right[:-1]`, Python-scalar product in the
//                      array's dtype), + weight * U1 in float64; kept in float64 for the edges
__global__ void __launch_bounds__(256) ds_centres_kernel(float* __restrict__ F, int size, int step, int m,
                                                         float keep32, double weight,
                                                         const double* __restrict__ u1,
                                                         double* __restrict__ centres) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= m * m) return;
    const int i = idx / m, j = idx - i * m;
    const int half = step >> 1;
    const float c00 = F[(size_t)(i * step) * size + j * step];
    const float c10 = F[(size_t)((i + 1) * step) * size + j * step];
    const float c01 = F[(size_t)(i * step) * size + (j + 1) * step];
    const float sum = __fadd_rn(__fadd_rn(c00, c10), __fadd_rn(c00, c01));
    const float t = __fmul_rn(__fmul_rn(keep32, sum), 0.25f);
    const double v = __dadd_rn((double)t, __dmul_rn(weight, u1[idx]));
    centres[idx] = v;
    F[(size_t)(i * step + half) * size + j * step + half] = (float)v;
}

//   edge midpoints, all float64 after the float32 corner sums:
//     rows (m + 1) x m:  keep * ((c[i][j] + c[i][j+1]) + ab[i][j]) / 4 + weight * U2
//                        ab[i] = centres[i] + centres[i-1], wrapping: ab[0] = ab[m] = centres[0] + centres[m-1]
//     cols m x (m + 1):  keep * ((c[i][j] + c[i+1][j]) + lr[i][j]) / 4 + weight * U3
//                        lr[:, j] = centres[:, j] + centres[:, j-1], lr[:, 0] = centres[:, 0] + centres[:, m-1],
//                        lr[i][m] = lr[0][i]  (the reference closes the wrap with the first ROW)
__global__ void __launch_bounds__(256) ds_edges_kernel(float* __restrict__ F, int size, int step, int m,
                                                       double keep, double weight,
                                                       const double* __restrict__ u2,
                                                       const double* __restrict__ u3,
                                                       const double* __restrict__ centres) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int n_rows = (m + 1) * m;
    const int half = step >> 1;
    if (idx < n_rows) {
        const int i = idx / m, j = idx - i * m;
        const int ia = (i == 0 || i == m) ? 0 : i, ib = (i == 0 || i == m) ? m - 1 : i - 1;
        const double ab = __dadd_rn(centres[ia * m + j], centres[ib * m + j]);
        const float right = __fadd_rn(F[(size_t)(i * step) * size + j * step],
                                      F[(size_t)(i * step) * size + (j + 1) * step]);
        const double t = __dmul_rn(__dmul_rn(keep, __dadd_rn((double)right, ab)), 0.25);
        F[(size_t)(i * step) * size + j * step + half] = (float)__dadd_rn(t, __dmul_rn(weight, u2[idx]));
    } else if (idx < 2 * n_rows) {
        const int k = idx - n_rows;
        const int i = k / (m + 1), j = k - i * (m + 1);
        // lr[r][c] for c < m
        auto lr = [&](int r, int c) {
            const int cb = c == 0 ? m - 1 : c - 1;
            return __dadd_rn(centres[r * m + c], centres[r * m + cb]);
        };
        const double v_lr = j < m ? lr(i, j) : lr(0, i);
        const float down = __fadd_rn(F[(size_t)(i * step) * size + j * step],
                                     F[(size_t)((i + 1) * step) * size + j * step]);
        const double t = __dmul_rn(__dmul_rn(keep, __dadd_rn((double)down, v_lr)), 0.25);
        F[(size_t)(i * step + half) * size + j * step] = (float)__dadd_rn(t, __dmul_rn(weight, u3[k]));
    }
}

// order-preserving map float -> uint32
__device__ __forceinline__ uint32_t float_key(float f) {
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_float(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}

// minmax[0] = min key, minmax[1] = max key of the crop (initialised to ~0 / 0 by the caller)
__global__ void __launch_bounds__(256) ds_minmax_kernel(const float* __restrict__ F, int size, int up, int left,
                                                        int h, int w, uint32_t* __restrict__ minmax) {
    uint32_t lo = 0xFFFFFFFFu, hi = 0u;
    const int64_t n = (int64_t)h * w;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
        const int y = (int)(p / w), x = (int)(p - (int64_t)y * w);
        const uint32_t k = float_key(F[(size_t)(up + y) * size + left + x]);
        lo = min(lo, k);
        hi = max(hi, k);
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, d));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, d));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(minmax, lo);
        atomicMax(minmax + 1, hi);
    }
}

// fog_image's normalisation, float32 in place like the reference:
//   mask -= min; mask /= max(mask); mask *= span; mask += ratio_min
__global__ void __launch_bounds__(256) ds_normalise_kernel(const float* __restrict__ F, int size, int up, int left,
                                                           int h, int w, const uint32_t* __restrict__ minmax,
                                                           float span, float ratio_min, float* __restrict__ out) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (int64_t)h * w) return;
    const int y = (int)(p / w), x = (int)(p - (int64_t)y * w);
    const float mn = key_float(minmax[0]);
    const float mx = __fsub_rn(key_float(minmax[1]), mn);  // max of (mask - min)
    float v = __fsub_rn(F[(size_t)(up + y) * size + left + x], mn);
    v = __fdiv_rn(v, mx);
    v = __fmul_rn(v, span);
    out[p] = __fadd_rn(v, ratio_min);
}

__global__ void ds_init_kernel(float* __restrict__ F, int size, float c00, float c01, float c11, float c10,
                               uint32_t* __restrict__ minmax) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        F[0] = c00;
        F[size - 1] = c01;
        F[(size_t)(size - 1) * size + size - 1] = c11;
        F[(size_t)(size - 1) * size] = c10;
        minmax[0] = 0xFFFFFFFFu;
        minmax[1] = 0u;
    }
}

}  // namespace vkb

using namespace vkb;

// number of uniform doubles the array draws of one field consume (effect.py:104-141)
static int64_t fog_draw_count(int size) {
    int64_t n = 0;
    for (int step = size - 1; step >= 2; step >>= 1) {
        const int64_t m = (size - 1) / step;
        n += m * m + 2 * m * (m + 1);
    }
    return n;
}

extern "C" int vkb_fog_draws(int32_t size, int64_t* count) {
    VKB_REQUIRE(count && size >= 3 && ((size - 1) & (size - 2)) == 0, "size must be 2^k + 1");
    *count = fog_draw_count(size);
    return VKB_OK;
}

extern "C" int vkb_fog_mask(const vkb_fog_params* p, float* field, double* centres, double* draws,
                            uint32_t* minmax, float* alpha, void* stream) {
    VKB_NVTX("vkb_fog_mask");
    VKB_REQUIRE(p && field && centres && draws && minmax && alpha, "bad arguments");
    const int size = p->size;
    VKB_REQUIRE(size >= 3 && ((size - 1) & (size - 2)) == 0, "size must be 2^k + 1");
    VKB_REQUIRE(p->height > 0 && p->width > 0 && p->up >= 0 && p->left >= 0
                    && p->up + p->height <= size && p->left + p->width <= size, "crop outside the field");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n = fog_draw_count(size);
    const int64_t gen_threads = (n + kDrawsPerThread - 1) / kDrawsPerThread;
    pcg64_uniform_kernel<<<(unsigned)((gen_threads + 127) / 128), 128, 0, st>>>(
        draws, n, p->state_hi, p->state_lo, p->inc_hi, p->inc_lo);
    ds_init_kernel<<<1, 32, 0, st>>>(field, size, p->corners[0], p->corners[1], p->corners[2],
                                     p->corners[3], minmax);
    int64_t off = 0;
    int level = 0;
    for (int step = size - 1; step >= 2; step >>= 1, ++level) {
        VKB_REQUIRE(level < VKB_FOG_MAX_LEVELS, "too many levels");
        const int m = (size - 1) / step;
        const double weight = p->weight[level], keep = 1.0 - weight;
        const double* u1 = draws + off;
        const double* u2 = u1 + (int64_t)m * m;
        const double* u3 = u2 + (int64_t)(m + 1) * m;
        off += (int64_t)m * m + 2 * (int64_t)m * (m + 1);
        ds_centres_kernel<<<(m * m + 255) / 256, 256, 0, st>>>(field, size, step, m, (float)keep, weight,
                                                               u1, centres);
        ds_edges_kernel<<<(2 * (m + 1) * m + 255) / 256, 256, 0, st>>>(field, size, step, m, keep, weight,
                                                                       u2, u3, centres);
    }
    const int64_t px = (int64_t)p->height * p->width;
    const int blocks = (int)((px + 255) / 256 < 1184 ? (px + 255) / 256 : 1184);
    ds_minmax_kernel<<<blocks, 256, 0, st>>>(field, size, p->up, p->left, p->height, p->width, minmax);
    ds_normalise_kernel<<<(unsigned)((px + 255) / 256), 256, 0, st>>>(
        field, size, p->up, p->left, p->height, p->width, minmax, p->ratio_span, p->ratio_min, alpha);
    return check_launch("vkb_fog_mask");
}

// ============================================================================================
// glass_blur's pixel permutation on the device (photometric/blur.py:232-262).  Per round every
// (2 * delta + 1)-th pixel ("centre") trades places with a random neighbour of the pixel that
// currently sits there; the shifts are NumPy bounded integers: Lemire's method on the 32-bit
// halves of the generator's 64-bit outputs (low half first; a half left over by an earlier draw
// comes first).  Half number p of the stream is again a pure function of (state, p) -- unless a
// draw is REJECTED (probability 2^-32 per draw for three values), which shifts everything behind
// it: the kernel raises a flag then and the caller repeats the call on the host.
// ============================================================================================
namespace vkb {

// 32-bit half number p (0-based) of the stream that starts with `cached` when has_cached is set
struct HalfStream {
    u128 s;       // state of the output the next half comes from (already stepped when `high`)
    bool high;    // the next half is the high half of the current output
};

__device__ __forceinline__ uint32_t half_at_start(u128 state, u128 inc, int has_cached, uint32_t cached,
                                                  int64_t p, HalfStream& hs, bool& from_cache) {
    // positions behind the cached half map onto outputs 1, 2, ...: q = p - has_cached
    from_cache = has_cached && p == 0;
    if (from_cache) {
        hs.s = state;  // next: low half of output 1
        hs.high = false;
        return cached;
    }
    const int64_t q = p - has_cached;
    u128 s = pcg_advance(state, inc, (uint64_t)(q >> 1) + 1);
    hs.s = s;
    const uint64_t hi = (uint64_t)(s >> 64), lo = (uint64_t)s;
    const uint64_t x = hi ^ lo;
    const unsigned r = (unsigned)(hi >> 58);
    const uint64_t o = (x >> r) | (x << ((64u - r) & 63u));
    hs.high = !(q & 1);  // after a low half comes the high half of the same output
    return (q & 1) ? (uint32_t)(o >> 32) : (uint32_t)o;
}

__device__ __forceinline__ uint32_t half_next(HalfStream& hs, u128 inc, bool was_cache) {
    if (!was_cache && hs.high) {  // high half of the current output
        const uint64_t hi = (uint64_t)(hs.s >> 64), lo = (uint64_t)hs.s;
        const uint64_t x = hi ^ lo;
        const unsigned r = (unsigned)(hi >> 58);
        const uint64_t o = (x >> r) | (x << ((64u - r) & 63u));
        hs.high = false;
        return (uint32_t)(o >> 32);
    }
    hs.s = hs.s * pcg_mult() + inc;
    const uint64_t hi = (uint64_t)(hs.s >> 64), lo = (uint64_t)hs.s;
    const uint64_t x = hi ^ lo;
    const unsigned r = (unsigned)(hi >> 58);
    const uint64_t o = (x >> r) | (x << ((64u - r) & 63u));
    hs.high = true;
    return (uint32_t)o;
}

constexpr int kGlassRun = 8;  // consecutive centres per thread (one jump-ahead per axis)

// bounded integer low + Lemire(u, span); flags a draw NumPy would have rejected
__device__ __forceinline__ int lemire_bounded(uint32_t u, uint32_t span, uint32_t threshold, int low,
                                              int32_t* __restrict__ flag) {
    const uint64_t m = (uint64_t)u * span;
    const uint32_t leftover = (uint32_t)m;
    if (leftover < span && leftover < threshold) *flag = 1;
    return low + (int)(m >> 32);
}

// phase 1 of a round: shifts, targets, and the values both sides hold BEFORE any write
__global__ void __launch_bounds__(128) glass_targets_kernel(
    const int32_t* __restrict__ pos_y, const int32_t* __restrict__ pos_x, int h, int w, int row0,
    int col0, int period, int delta, int ny, int nx, uint64_t state_hi, uint64_t state_lo,
    uint64_t inc_hi, uint64_t inc_lo, int has_cached, uint32_t cached, int round,
    int32_t* __restrict__ target, int32_t* __restrict__ at_centre, int32_t* __restrict__ at_target,
    int32_t* __restrict__ owner, int32_t* __restrict__ flag) {
    const int64_t n = (int64_t)ny * nx;
    const int64_t first = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * kGlassRun;
    if (first >= n) return;
    const u128 state = make_u128(state_hi, state_lo), inc = make_u128(inc_hi, inc_lo);
    const uint32_t span = 2u * (uint32_t)delta + 1u;
    const uint32_t threshold = (uint32_t)((0x100000000ull - span) % span);
    HalfStream sy, sx;
    bool cy_cache, cx_cache;
    uint32_t uy = half_at_start(state, inc, has_cached, cached, first, sy, cy_cache);
    uint32_t ux = half_at_start(state, inc, has_cached, cached, n + first, sx, cx_cache);
    const int64_t last = first + kGlassRun < n ? first + kGlassRun : n;
    for (int64_t k = first; k < last; ++k) {
        const int shift_y = lemire_bounded(uy, span, threshold, -delta, flag);
        const int shift_x = lemire_bounded(ux, span, threshold, -delta, flag);
        const int iy = (int)(k / nx), ix = (int)(k - (int64_t)iy * nx);
        const int c = (row0 + iy * period) * w + (col0 + ix * period);
        const int vy = pos_y[c], vx = pos_x[c];
        const int ty = min(max(vy + shift_y, 0), h - 1), tx = min(max(vx + shift_x, 0), w - 1);
        const int t = ty * w + tx;
        target[k] = t;
        at_centre[2 * k] = vy;
        at_centre[2 * k + 1] = vx;
        at_target[2 * k] = pos_y[t];
        at_target[2 * k + 1] = pos_x[t];
        // duplicates among the targets: the last centre in C order writes last
        atomicMax(owner + t, (round << 24) | (int)k);
        if (k + 1 < last) {
            uy = half_next(sy, inc, cy_cache);
            ux = half_next(sx, inc, cx_cache);
            cy_cache = cx_cache = false;
        }
    }
}

// phase 2: pos[centres] = at_target
__global__ void __launch_bounds__(256) glass_centres_kernel(int32_t* __restrict__ pos_y, int32_t* __restrict__ pos_x,
                                                            int w, int row0, int col0, int period, int ny, int nx,
                                                            const int32_t* __restrict__ at_target) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= ny * nx) return;
    const int iy = k / nx, ix = k - iy * nx;
    const int c = (row0 + iy * period) * w + (col0 + ix * period);
    pos_y[c] = at_target[2 * k];
    pos_x[c] = at_target[2 * k + 1];
}

// phase 3: pos[targets] = at_centre, the winner of every target only
__global__ void __launch_bounds__(256) glass_scatter_kernel(int32_t* __restrict__ pos_y, int32_t* __restrict__ pos_x,
                                                            int n, int round, const int32_t* __restrict__ target,
                                                            const int32_t* __restrict__ at_centre,
                                                            const int32_t* __restrict__ owner) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int t = target[k];
    if (owner[t] != ((round << 24) | k)) return;
    pos_y[t] = at_centre[2 * k];
    pos_x[t] = at_centre[2 * k + 1];
}

__global__ void __launch_bounds__(256) glass_init_kernel(int32_t* __restrict__ pos_y, int32_t* __restrict__ pos_x,
                                                         int32_t* __restrict__ owner, int h, int w,
                                                         int32_t* __restrict__ flag) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p == 0) *flag = 0;
    if (p >= (int64_t)h * w) return;
    const int y = (int)(p / w);
    pos_y[p] = y;
    pos_x[p] = (int)(p - (int64_t)y * w);
    owner[p] = -1;
}

}  // namespace vkb

extern "C" int vkb_glass_init(int32_t* pos_y, int32_t* pos_x, int32_t* owner, int32_t h, int32_t w,
                              int32_t* flag, void* stream) {
    VKB_REQUIRE(pos_y && pos_x && owner && flag && h > 0 && w > 0, "bad arguments");
    const int64_t n = (int64_t)h * w;
    glass_init_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(pos_y, pos_x, owner, h, w, flag);
    return check_launch("glass_init_kernel");
}

extern "C" int vkb_glass_round(int32_t* pos_y, int32_t* pos_x, int32_t* owner, int32_t h, int32_t w,
                               int32_t row0, int32_t col0, int32_t delta, int32_t round,
                               uint64_t state_hi, uint64_t state_lo, uint64_t inc_hi, uint64_t inc_lo,
                               int32_t has_cached, uint32_t cached, int32_t* target, int32_t* at_centre,
                               int32_t* at_target, int32_t* flag, void* stream) {
    VKB_NVTX("vkb_glass_round");
    VKB_REQUIRE(pos_y && pos_x && owner && target && at_centre && at_target && flag, "bad arguments");
    VKB_REQUIRE(delta >= 1 && round >= 0 && round < 127, "delta >= 1, at most 127 rounds");
    const int period = 2 * delta + 1;
    VKB_REQUIRE(row0 >= 0 && row0 < period && col0 >= 0 && col0 < period, "bad offsets");
    const int ny = h - delta > row0 ? (h - delta - row0 + period - 1) / period : 0;
    const int nx = w - delta > col0 ? (w - delta - col0 + period - 1) / period : 0;
    const int64_t n = (int64_t)ny * nx;
    if (n == 0) return VKB_OK;
    VKB_REQUIRE(n < (1 << 24), "too many centres per round");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t threads = (n + kGlassRun - 1) / kGlassRun;
    glass_targets_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, st>>>(
        pos_y, pos_x, h, w, row0, col0, period, delta, ny, nx, state_hi, state_lo, inc_hi, inc_lo,
        has_cached, cached, round, target, at_centre, at_target, owner, flag);
    glass_centres_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(pos_y, pos_x, w, row0, col0, period,
                                                                     ny, nx, at_target);
    glass_scatter_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(pos_y, pos_x, (int)n, round, target,
                                                                     at_centre, owner);
    return check_launch("vkb_glass_round");
}
