// photometric.cu -- sm_100a kernels + C ABI for blends, colour conversion, the fused per-pixel
// photometric ops, the 8.8 fixed-point Gaussian, noise and streaks.
//
//   blend_*            element/opt.py:118-209 through box.py:311-416
//   cvt_color_kernel   element/image.py:771-814
//   color_ops_kernel   photometric/color.py:32-439
//   gaussian_blur_*    photometric/blur.py:49-76   (cv.GaussianBlur uint8)
//   noise_*            photometric/noise.py:25-190
//   streak_*           photometric/streak.py:44-279
//
// All of it is byte-stream work bound by HBM bandwidth: one read and one write per pixel,
// no tensor cores.
#include <cuda.h>  // CUtensorMap and its enums only: the encoder is looked up at run time
#include <curand_kernel.h>
#include <cstdlib>
#include <mutex>
#include "common.cuh"
#include "vkb_color.cuh"

namespace vkb {

// Division tables of RGB2HSV: looked up with per-pixel indices, so they live in global memory
// (coalesced staging into shared memory per block), not in the constant bank, whose reads
// serialise when the lanes of a warp use different addresses.
__device__ HsvTables c_hsv;

// `c_hsv` exists once per device: every device gets its own upload, under a lock (first calls
// from several threads / several devices of one process).
static int ensure_tables(cudaStream_t st) {
    static std::mutex mutex;
    static bool ready[64] = {};
    static HsvTables host;
    static bool host_ready = false;
    int dev = 0;
    VKB_CUDA(cudaGetDevice(&dev));
    VKB_REQUIRE(dev >= 0 && dev < 64, "device ordinal out of range");
    std::lock_guard<std::mutex> lock(mutex);
    if (!ready[dev]) {
        if (!host_ready) {
            fill_hsv_tables(host);
            host_ready = true;
        }
        VKB_CUDA(cudaMemcpyToSymbolAsync(c_hsv, &host, sizeof(host), 0, cudaMemcpyHostToDevice, st));
        VKB_CUDA(cudaStreamSynchronize(st));
        ready[dev] = true;
    }
    return VKB_OK;
}

// ============================================================================================
// Blend
// ============================================================================================
__device__ __forceinline__ void blend_pixel(const vkb_blend_item& it, int ry, int rx) {
    bool active = true;
    float a = it.alpha;
    if (it.alpha_arr) {
        a = it.alpha_arr[(long long)ry * it.alpha_pitch + rx];
        active = a > 0.f;
    }
    if (it.mask) active = it.mask[(long long)ry * it.mask_pitch + rx] != 0;
    if (!active) return;
    const long long di = ((long long)(it.box_y + ry) * it.dst_w + (it.box_x + rx)) * it.channels;
    const long long vi = ((long long)ry * it.value_pitch + rx) * it.channels;
    const bool do_blend = it.alpha_arr != nullptr || a < 1.0f;
    for (int c = 0; c < it.channels; ++c) {
        if (it.dst_f32) {
            float* d = reinterpret_cast<float*>(it.dst) + di + c;
            const float v = it.value_arr ? reinterpret_cast<const float*>(it.value_arr)[vi + c]
                                         : it.value_const[c];
            if (do_blend) *d = blend_f32(*d, v, a);
            else if ((it.keep_mode & 3) == 1) { if (*d < v) *d = v; }
            else if ((it.keep_mode & 3) == 2) { if (*d > v) *d = v; }
            else *d = v;
        } else {
            uint8_t* d = reinterpret_cast<uint8_t*>(it.dst) + di + c;
            // a constant value is cast to the destination dtype first (prep_value's np.full_like,
            // element/opt.py:96-115) unless the caller blends with its own float constant
            // (VKB_BLEND_FLOAT_CONST: fog on a GRAYSCALE page, effect.py:194-197)
            const int v = it.value_arr ? (int)reinterpret_cast<const uint8_t*>(it.value_arr)[vi + c]
                                       : (int)it.value_const[c];
            const float vf = (!it.value_arr && (it.keep_mode & VKB_BLEND_FLOAT_CONST))
                                 ? it.value_const[c] : (float)v;
            if (do_blend) *d = (uint8_t)(int)blend_f32((float)*d, vf, a);  // truncation
            else if ((it.keep_mode & 3) == 1) { if (*d < v) *d = (uint8_t)v; }
            else if ((it.keep_mode & 3) == 2) { if (*d > v) *d = (uint8_t)v; }
            else *d = (uint8_t)v;
        }
    }
}

__global__ void __launch_bounds__(256) blend_fill_kernel(const vkb_blend_item it) {
    const int rx = blockIdx.x * 32 + threadIdx.x;
    const int ry0 = blockIdx.y * 32 + threadIdx.y;
    if (rx >= it.box_w) return;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int ry = ry0 + 8 * k;
        if (ry < it.box_h) blend_pixel(it, ry, rx);
    }
}

// Ordered draw list: every thread owns one destination pixel and walks the items in order,
// so overlapping items compose exactly like successive fill calls.
__global__ void __launch_bounds__(256) blend_draw_list_kernel(const vkb_blend_item* __restrict__ items,
                                                              int n_items, int dst_h, int dst_w) {
    extern __shared__ unsigned char smem_raw[];
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y0 = blockIdx.y * 32 + threadIdx.y;
    const int tx0 = blockIdx.x * 32, ty0 = blockIdx.y * 32;
    for (int i = 0; i < n_items; ++i) {
        const vkb_blend_item& it = items[i];
        // tile / box rejection is block uniform
        if (it.box_x > tx0 + 31 || it.box_x + it.box_w <= tx0 || it.box_y > ty0 + 31
            || it.box_y + it.box_h <= ty0)
            continue;
        const int rx = x - it.box_x;
        if (rx < 0 || rx >= it.box_w || x >= dst_w) continue;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int y = y0 + 8 * k;
            const int ry = y - it.box_y;
            if (y < dst_h && ry >= 0 && ry < it.box_h) blend_pixel(it, ry, rx);
        }
    }
    (void)smem_raw;
}

// ============================================================================================
// Colour conversion
// ============================================================================================
__global__ void __launch_bounds__(256) cvt_color_kernel(const uint8_t* __restrict__ src,
                                                        uint8_t* __restrict__ dst, long long n, int code) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int a, b, c;
    switch (code) {
        case VKB_CVT_RGB2HSV: {
            rgb2hsv_full(c_hsv, src[3 * i], src[3 * i + 1], src[3 * i + 2], a, b, c);
            dst[3 * i] = a; dst[3 * i + 1] = b; dst[3 * i + 2] = c;
        } break;
        case VKB_CVT_HSV2RGB: {
            hsv2rgb_full(src[3 * i], src[3 * i + 1], src[3 * i + 2], a, b, c);
            dst[3 * i] = a; dst[3 * i + 1] = b; dst[3 * i + 2] = c;
        } break;
        case VKB_CVT_RGB2HSL: {
            int h, l, s;
            rgb2hls_full(src[3 * i], src[3 * i + 1], src[3 * i + 2], h, l, s);
            dst[3 * i] = h; dst[3 * i + 1] = s; dst[3 * i + 2] = l;  // stored H, S, L
        } break;
        case VKB_CVT_HSL2RGB: {
            hls2rgb_full(src[3 * i], src[3 * i + 2], src[3 * i + 1], a, b, c);
            dst[3 * i] = a; dst[3 * i + 1] = b; dst[3 * i + 2] = c;
        } break;
        case VKB_CVT_RGB2GRAY:
            dst[i] = rgb2gray(src[3 * i], src[3 * i + 1], src[3 * i + 2]);
            break;
        case VKB_CVT_GRAY2RGB:
            dst[3 * i] = dst[3 * i + 1] = dst[3 * i + 2] = src[i];
            break;
        case VKB_CVT_RGBA2RGB:
            dst[3 * i] = src[4 * i]; dst[3 * i + 1] = src[4 * i + 1]; dst[3 * i + 2] = src[4 * i + 2];
            break;
        case VKB_CVT_RGB2RGBA:
            dst[4 * i] = src[3 * i]; dst[4 * i + 1] = src[3 * i + 1]; dst[4 * i + 2] = src[3 * i + 2];
            dst[4 * i + 3] = 255;
            break;
        case VKB_CVT_GRAY2RGBA:
            dst[4 * i] = dst[4 * i + 1] = dst[4 * i + 2] = src[i];
            dst[4 * i + 3] = 255;
            break;
        case VKB_CVT_RGBA2GRAY:
            dst[i] = rgb2gray(src[4 * i], src[4 * i + 1], src[4 * i + 2]);
            break;
        default: break;
    }
}

// ============================================================================================
// Fused per-pixel photometric ops.
// ============================================================================================
struct ColorOpList {
    vkb_color_op ops[VKB_MAX_COLOR_OPS];
    int n;
};

// The division tables are looked up with per-pixel indices: constant memory would serialise the
// lanes of a warp, so every block works from a shared-memory copy.
__device__ __forceinline__ void stage_hsv_tables(HsvTables& sm, int tid, int n_threads) {
    const int* __restrict__ src = c_hsv.sdiv;
    int* dst = sm.sdiv;
    for (int i = tid; i < 512; i += n_threads) dst[i] = __ldg(src + i);
}
static_assert(sizeof(HsvTables) == 512 * sizeof(int), "HsvTables: two 256-entry tables");

// photometric/noise.py:25-190 with the counter-based generator of vkb_noise_philox: keyed by
// (seed, pixel index), so a page gives the same result alone or in a batch.  Out of line: the
// Philox state and the double arithmetic stay out of the fused chain kernel (registers): noise
// is its own batched pass (noise_philox_batched_kernel).
__device__ __forceinline__ void noise_op(const vkb_color_op& op, int* px, int channels, long long i) {
    const unsigned long long seed =
        (unsigned long long)(unsigned)op.i1 | ((unsigned long long)(unsigned)op.i2 << 32);
    curandStatePhilox4_32_10_t st;
    curand_init(seed, (unsigned long long)i, 0, &st);
    const double p0 = (double)op.f0, p1 = (double)op.f1;
    if (op.i0 == 2) {
        const double u = curand_uniform_double(&st);
        const double presv = 1.0 - p0 - p1;
        for (int c = 0; c < channels; ++c) px[c] = u < presv ? px[c] : (u < presv + p0 ? 255 : 0);
        return;
    }
    for (int c = 0; c < channels; ++c) {
        const int v = px[c];
        if (op.i0 == 0) {
            px[c] = clip_u8(v + (int)rint((double)curand_normal(&st) * p0));
        } else if (op.i0 == 1) {
            px[c] = clip_u8((int)curand_poisson(&st, (double)v));
        } else {
            const double r = (double)v + (double)v * ((double)curand_normal(&st) * p0);
            px[c] = (int)fmin(fmax(r, 0.0), 255.0);
        }
    }
}

// photometric/streak.py:44-106: periodic lines with dashes, vertical mask blended first,
// horizontal second (crossings get alpha twice).  Only compiled into the POS variant of the chain
// kernel, so the common chains keep their register budget.
__device__ __forceinline__ void line_streak_op(const vkb_color_op& op, int* px, int channels, int x,
                                               int y) {
    const int thickness = op.i0, step = op.i0 + op.i1;
    const int dash_t = op.i2 & 0xFFFF, dash_g = (op.i2 >> 16) & 0xFFFF;
    const bool dashed = dash_t > 0 && dash_g > 0;
    const bool vert = (op.i3 & 1) && (x % step) < thickness
                      && !(dashed && (y % (dash_t + dash_g)) < dash_g);
    const bool hori = (op.i3 & 2) && (y % step) < thickness
                      && !(dashed && (x % (dash_t + dash_g)) < dash_g);
    const float col[4] = {op.f1, op.f2, op.f3, op.g0};
    for (int rep = 0; rep < 2; ++rep) {
        if (!(rep == 0 ? vert : hori)) continue;
        for (int c = 0; c < channels; ++c) {
            if (op.f0 >= 1.0f) px[c] = (int)col[c];
            else px[c] = (int)blend_f32((float)px[c], col[c], op.f0);
        }
    }
}

// (x, y): position inside the page, i: linear pixel index of the page (position dependent ops)
template <bool POS = false>
__device__ __forceinline__ void apply_color_op(const vkb_color_op& op, int* px, int channels,
                                               const HsvTables& tables, int x = 0, int y = 0,
                                               long long i = 0) {
    if (POS && op.kind == VKB_OP_LINE_STREAK) {
        line_streak_op(op, px, channels, x, y);
        return;
    }
    switch (op.kind) {
        case VKB_OP_MEAN_SHIFT: {
            for (int c = 0; c < channels; ++c) {
                if (!((op.i2 >> c) & 1)) continue;
                int v = px[c];
                bool hit = true;
                if (op.i1 >= 0) hit = op.i0 > 0 ? (v <= op.i1) : (op.i1 <= v);
                if (hit) v += op.i0;
                px[c] = op.i3 ? floor_mod_256(v) : clip_u8(v);
            }
        } break;
        case VKB_OP_HUE_SHIFT_RGB: {
            int h, s, v;
            rgb2hsv_full(tables, px[0], px[1], px[2], h, s, v);
            h = floor_mod_256(h + op.i0);
            hsv2rgb_full(h, s, v, px[0], px[1], px[2]);
        } break;
        case VKB_OP_LIGHT_SHIFT_RGB: {
            if (op.i1 == 0) {
                int h, l, s;
                rgb2hls_full(px[0], px[1], px[2], h, l, s);
                l = clip_u8(l + op.i0);
                hls2rgb_full(h, l, s, px[0], px[1], px[2]);
            } else {
                int h, s, v;
                rgb2hsv_full(tables, px[0], px[1], px[2], h, s, v);
                v = clip_u8(v + op.i0);
                hsv2rgb_full(h, s, v, px[0], px[1], px[2]);
            }
        } break;
        case VKB_OP_STD_SHIFT: {
            const float sub[3] = {op.f1, op.f2, op.f3};
            for (int c = 0; c < channels; ++c) {
                if (!((op.i2 >> c) & 1)) continue;
                const float v = __fsub_rn(__fmul_rn((float)px[c], op.f0), sub[c]);
                px[c] = round_u8(v);
            }
        } break;
        case VKB_OP_COMPLEMENT: {
            for (int c = 0; c < channels; ++c) {
                if (!((op.i2 >> c) & 1)) continue;
                bool hit = true;
                if (op.i1 >= 0) hit = op.i3 ? (px[c] <= op.i1) : (op.i1 <= px[c]);
                if (hit) px[c] = 255 - px[c];
            }
        } break;
        case VKB_OP_POSTERIZE: {
            for (int c = 0; c < channels; ++c)
                if ((op.i2 >> c) & 1) px[c] &= op.i0;
        } break;
        case VKB_OP_COLOR_BALANCE: {
            if (channels >= 3) {
                const float gray = (float)rgb2gray(px[0], px[1], px[2]);
                for (int c = 0; c < 3; ++c) {
                    const float v = __fadd_rn(__fmul_rn(op.f1, gray), __fmul_rn(op.f0, (float)px[c]));
                    px[c] = clip_u8((int)fminf(fmaxf(v, 0.f), 255.f));  // clip, then truncate
                }
            }
        } break;
        case VKB_OP_PERMUTE: {
            int t[4] = {px[0], px[1], px[2], px[3]};
            for (int c = 0; c < channels; ++c) px[c] = t[(op.i0 >> (4 * c)) & 0xF];
        } break;
        case VKB_OP_BOUNDARY_EQ: {
            const float mn[3] = {op.f0, op.f1, op.f2};
            const float sc[3] = {op.g0, op.g1, op.g2};
            for (int c = 0; c < channels; ++c) {
                if (!((op.i2 >> c) & 1)) continue;
                px[c] = round_u8(__fmul_rn(__fsub_rn((float)px[c], mn[c]), sc[c]));
            }
        } break;
        default: break;
    }
}

template <int C>
__global__ void __launch_bounds__(256) color_ops_kernel(const uint8_t* __restrict__ src,
                                                        uint8_t* __restrict__ dst, long long n,
                                                        const ColorOpList list) {
    __shared__ HsvTables tables;
    stage_hsv_tables(tables, threadIdx.x, 256);
    __syncthreads();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int px[4] = {0, 0, 0, 0};
#pragma unroll
    for (int c = 0; c < C; ++c) px[c] = src[i * C + c];
    for (int k = 0; k < list.n; ++k) apply_color_op(list.ops[k], px, C, tables);
#pragma unroll
    for (int c = 0; c < C; ++c) dst[i * C + c] = (uint8_t)px[c];
}

// per-channel sum/min/max
__global__ void __launch_bounds__(256) channel_stats_kernel(const uint8_t* __restrict__ src, long long n,
                                                            int channels, unsigned long long* sums,
                                                            unsigned int* mins, unsigned int* maxs) {
    unsigned long long s[3] = {0, 0, 0};
    unsigned int mn[3] = {255, 255, 255}, mx[3] = {0, 0, 0};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        for (int c = 0; c < channels; ++c) {
            const unsigned int v = src[i * channels + c];
            s[c] += v;
            mn[c] = min(mn[c], v);
            mx[c] = max(mx[c], v);
        }
    }
    for (int c = 0; c < channels; ++c) {
        for (int o = 16; o > 0; o >>= 1) {
            s[c] += __shfl_xor_sync(0xffffffffu, s[c], o);
            mn[c] = min(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
            mx[c] = max(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
        }
        if ((threadIdx.x & 31) == 0) {
            atomicAdd(&sums[c], s[c]);
            atomicMin(&mins[c], mn[c]);
            atomicMax(&maxs[c], mx[c]);
        }
    }
}

__global__ void channel_stats_init_kernel(unsigned long long* sums, unsigned int* mins, unsigned int* maxs) {
    if (threadIdx.x < 3) {
        sums[threadIdx.x] = 0;
        mins[threadIdx.x] = 255;
        maxs[threadIdx.x] = 0;
    }
}

// per-channel 256-bin histogram (cv.equalizeHist, photometric/color.py:264-298) and LUT apply
__global__ void __launch_bounds__(256) histogram_kernel(const uint8_t* __restrict__ src, long long n,
                                                        int channels, unsigned int* __restrict__ out) {
    __shared__ unsigned int sh[3 * 256];
    for (int i = threadIdx.x; i < 3 * 256; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        for (int c = 0; c < channels; ++c) atomicAdd(&sh[c * 256 + src[i * channels + c]], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < channels * 256; i += blockDim.x)
        if (sh[i]) atomicAdd(&out[i], sh[i]);
}

template <int C>
__global__ void __launch_bounds__(256) apply_lut_kernel(const uint8_t* __restrict__ src,
                                                        uint8_t* __restrict__ dst, long long n,
                                                        const uint8_t* __restrict__ lut, int bits) {
    __shared__ uint8_t sl[3 * 256];
    for (int i = threadIdx.x; i < 3 * 256; i += blockDim.x) sl[i] = lut[i];
    __syncthreads();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const uint8_t v = src[i * C + c];
        dst[i * C + c] = ((bits >> c) & 1) && c < 3 ? sl[c * 256 + v] : v;
    }
}

// ============================================================================================
// Gaussian blur, uint8, 8.8 fixed point, BORDER_REFLECT_101.  Block = 32 x 8 threads on a
// 32 x 32 output tile; the tile plus halo is staged in shared memory, the horizontal pass keeps
// its saturated 16-bit rows in shared memory, the vertical pass writes the result.
// ============================================================================================
constexpr int kGaussMaxR = 8;
struct GaussKernel {
    int k[2 * kGaussMaxR + 1];
    int r;
};

__device__ __forceinline__ int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) {
        if (i < 0) i = -i;
        else i = 2 * (n - 1) - i;
    }
    return i;
}

template <int C>
__global__ void __launch_bounds__(256) gaussian_blur_kernel(const uint8_t* __restrict__ src,
                                                            uint8_t* __restrict__ dst, int h, int w,
                                                            const GaussKernel gk) {
    extern __shared__ unsigned char smem[];
    const int r = gk.r;
    const int TW = 32 + 2 * r, TH = 32 + 2 * r;
    uint8_t* tile = smem;                                                  // TH x TW x C
    unsigned short* rows = reinterpret_cast<unsigned short*>(smem + ((TH * TW * C + 3) & ~3));  // TH x 32 x C
    const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    for (int i = tid; i < TH * TW; i += 256) {
        const int ty = i / TW, tx = i - ty * TW;
        const int sy = reflect101(y0 + ty - r, h), sx = reflect101(x0 + tx - r, w);
        const uint8_t* p = src + ((long long)sy * w + sx) * C;
#pragma unroll
        for (int c = 0; c < C; ++c) tile[i * C + c] = p[c];
    }
    __syncthreads();
    for (int i = tid; i < TH * 32; i += 256) {
        const int ty = i >> 5, tx = i & 31;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            int acc = 0;
            for (int k = 0; k <= 2 * r; ++k) acc += (int)tile[(ty * TW + tx + k) * C + c] * gk.k[k];
            rows[i * C + c] = (unsigned short)min(acc, 65535);
        }
    }
    __syncthreads();
    const int x = x0 + threadIdx.x;
    if (x >= w) return;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int ly = threadIdx.y + 8 * j;
        const int y = y0 + ly;
        if (y >= h) break;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            int acc = 0;
            for (int k = 0; k <= 2 * r; ++k) acc += (int)rows[((ly + k) * 32 + threadIdx.x) * C + c] * gk.k[k];
            dst[((long long)y * w + x) * C + c] = (uint8_t)min((acc + (1 << 15)) >> 16, 255);
        }
    }
}

// --------------------------------------------------------------------------------------------
// The same blur with the tile + halo of INTERIOR tiles staged by one TMA box load
// (cp.async.bulk.tensor.2d, SASS UTMALDG): the page is described to the copy engine as a 2-D
// byte tensor [h][w * C]; one elected thread posts the expected byte count on an mbarrier and
// issues the copy, the other 255 threads skip the per-element reflect-101 / byte-load loop
// entirely.  Tiles that touch the page border keep the loop (TMA fills out-of-range bytes with
// zeros, cv2 wants BORDER_REFLECT_101).  Needs a 16-byte aligned base and row pitch; other pages
// take gaussian_blur_kernel.  The first byte of a box must be 16-byte aligned as well (measured:
// UTMALDG raises "illegal instruction" otherwise, tools/ubench/tma_probe.cu), and a tile + halo
// starts at byte (x0 - r) * C: the box therefore starts MIS = (-r * C) mod 16 bytes earlier -- the
// same for every tile because 32 * C is a multiple of 16 -- and the passes read at that constant
// offset.  PITCH = bytes per shared-memory tile row = box width = roundup16(MIS + (32 + 2r) * C).
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

template <int C>
__global__ void __launch_bounds__(256) gaussian_blur_tma_kernel(const __grid_constant__ CUtensorMap tmap,
                                                                const uint8_t* __restrict__ src,
                                                                uint8_t* __restrict__ dst, int h, int w,
                                                                const GaussKernel gk, int PITCH, int MIS) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bar;
    const int r = gk.r;
    const int TW = 32 + 2 * r, TH = 32 + 2 * r;
    const uint8_t* tile = smem + MIS;                                                // TH x PITCH
    unsigned short* rows = reinterpret_cast<unsigned short*>(smem + ((TH * PITCH + 15) & ~15));  // TH x 32 x C
    const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    const bool interior = x0 - r >= 0 && y0 - r >= 0 && x0 + 32 + r <= w && y0 + 32 + r <= h;
    if (interior) {
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)),
                         "r"(TH * PITCH)
                         : "memory");
            asm volatile(
                "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
                " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem)),
                "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(smem_u32(&bar)), "r"((x0 - r) * C - MIS), "r"(y0 - r)
                : "memory");
        }
        __syncthreads();  // the barrier is initialised before anybody polls it
        uint32_t done = 0;
        while (!done) {
            asm volatile(
                "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
                "selp.u32 %0, 1, 0, p;\n}"
                : "=r"(done)
                : "r"(smem_u32(&bar)), "r"(0)
                : "memory");
        }
    } else {
        for (int i = tid; i < TH * TW; i += 256) {
            const int ty = i / TW, tx = i - ty * TW;
            const int sy = reflect101(y0 + ty - r, h), sx = reflect101(x0 + tx - r, w);
            const uint8_t* p = src + ((long long)sy * w + sx) * C;
#pragma unroll
            for (int c = 0; c < C; ++c) smem[MIS + ty * PITCH + tx * C + c] = p[c];
        }
        __syncthreads();
    }
    for (int i = tid; i < TH * 32; i += 256) {
        const int ty = i >> 5, tx = i & 31;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            int acc = 0;
            for (int k = 0; k <= 2 * r; ++k) acc += (int)tile[ty * PITCH + (tx + k) * C + c] * gk.k[k];
            rows[i * C + c] = (unsigned short)min(acc, 65535);
        }
    }
    __syncthreads();
    const int x = x0 + threadIdx.x;
    if (x >= w) return;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int ly = threadIdx.y + 8 * j;
        const int y = y0 + ly;
        if (y >= h) break;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            int acc = 0;
            for (int k = 0; k <= 2 * r; ++k) acc += (int)rows[((ly + k) * 32 + threadIdx.x) * C + c] * gk.k[k];
            dst[((long long)y * w + x) * C + c] = (uint8_t)min((acc + (1 << 15)) >> 16, 255);
        }
    }
}

typedef CUresult (*TensorMapEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                           const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                           const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                           CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// cuTensorMapEncodeTiled through the runtime's driver entry point table: the library keeps no link
// dependency on libcuda.so (it must load on machines without a driver for the symbol checks).
static TensorMapEncodeTiledFn tensor_map_encoder() {
    static TensorMapEncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        (void)cudaGetLastError();
        return reinterpret_cast<TensorMapEncodeTiledFn>(p);
    }();
    return fn;
}

// Returns 1 when the TMA form was launched, 0 when the page does not qualify, < 0 on error.
static int gaussian_blur_tma(const uint8_t* src, uint8_t* dst, int h, int w, int channels,
                             const GaussKernel& gk, cudaStream_t st) {
    const char* off = getenv("VKB_BLUR_NO_TMA");
    if (off && off[0] == '1') return 0;
    const long long pitch = (long long)w * channels;
    if ((pitch & 15) || (reinterpret_cast<uintptr_t>(src) & 15)) return 0;
    TensorMapEncodeTiledFn encode = tensor_map_encoder();
    if (!encode) return 0;
    const int TW = 32 + 2 * gk.r;
    const int mis = (16 - (gk.r * channels) % 16) % 16;
    const int box_w = (mis + TW * channels + 15) & ~15;
    if (box_w > 256 || TW > 256) return 0;
    CUtensorMap tmap;
    const cuuint64_t dims[2] = {(cuuint64_t)pitch, (cuuint64_t)h};
    const cuuint64_t strides[1] = {(cuuint64_t)pitch};
    const cuuint32_t box[2] = {(cuuint32_t)box_w, (cuuint32_t)TW};
    const cuuint32_t elem[2] = {1, 1};
    if (encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t*>(src), dims, strides, box, elem,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return 0;
    const size_t smem = (((size_t)TW * box_w + 15) & ~(size_t)15) + (size_t)TW * 32 * channels * 2;
    dim3 grid((w + 31) / 32, (h + 31) / 32);
    if (channels == 1) gaussian_blur_tma_kernel<1><<<grid, dim3(32, 8), smem, st>>>(tmap, src, dst, h, w, gk, box_w, mis);
    else if (channels == 3) gaussian_blur_tma_kernel<3><<<grid, dim3(32, 8), smem, st>>>(tmap, src, dst, h, w, gk, box_w, mis);
    else gaussian_blur_tma_kernel<4><<<grid, dim3(32, 8), smem, st>>>(tmap, src, dst, h, w, gk, box_w, mis);
    const int rc = check_launch("gaussian_blur_tma_kernel");
    return rc == VKB_OK ? 1 : rc;
}

// ============================================================================================
// Noise
// ============================================================================================
__global__ void __launch_bounds__(256) noise_philox_kernel(const uint8_t* __restrict__ src,
                                                           uint8_t* __restrict__ dst, long long n_pixels,
                                                           int channels, int kind, double p0, double p1,
                                                           unsigned long long seed) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pixels) return;
    curandStatePhilox4_32_10_t st;
    curand_init(seed, (unsigned long long)i, 0, &st);
    if (kind == 2) {  // impulse: one categorical per pixel, all channels (noise.py:132-142)
        const double u = curand_uniform_double(&st);
        const double presv = 1.0 - p0 - p1;
        for (int c = 0; c < channels; ++c) {
            const uint8_t v = src[i * channels + c];
            dst[i * channels + c] = u < presv ? v : (u < presv + p0 ? 255 : 0);
        }
        return;
    }
    for (int c = 0; c < channels; ++c) {
        const int v = src[i * channels + c];
        int out;
        if (kind == 0) {
            const double nz = rint((double)curand_normal(&st) * p0);
            out = clip_u8(v + (int)nz);
        } else if (kind == 1) {
            out = clip_u8((int)curand_poisson(&st, (double)v));
        } else {
            const double nz = (double)curand_normal(&st) * p0;
            const double r = (double)v + (double)v * nz;
            out = (int)fmin(fmax(r, 0.0), 255.0);  // clip then truncate (noise.py:182)
        }
        dst[i * channels + c] = (uint8_t)out;
    }
}

__global__ void __launch_bounds__(256) noise_field_kernel(const uint8_t* __restrict__ src,
                                                          uint8_t* __restrict__ dst, long long n_pixels,
                                                          int channels, int kind,
                                                          const void* __restrict__ field) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pixels) return;
    for (int c = 0; c < channels; ++c) {
        const long long e = i * channels + c;
        const int v = src[e];
        int out;
        if (kind == 0) {
            out = clip_u8(v + (int)reinterpret_cast<const short*>(field)[e]);
        } else if (kind == 1) {
            const long long s = reinterpret_cast<const long long*>(field)[e];
            out = s < 0 ? 0 : (s > 255 ? 255 : (int)s);
        } else if (kind == 2) {
            const long long cat = reinterpret_cast<const long long*>(field)[i];
            out = cat == 1 ? 255 : (cat == 2 ? 0 : v);
        } else {
            // float32 mat + float32 mat * float64 noise -> float64 (noise.py:181-182)
            const double nz = reinterpret_cast<const double*>(field)[e];
            const double r = __dadd_rn((double)v, __dmul_rn((double)v, nz));
            out = (int)fmin(fmax(r, 0.0), 255.0);
        }
        dst[e] = (uint8_t)out;
    }
}

// ============================================================================================
// Streaks
// ============================================================================================
struct StreakColor {
    float c[4];
};

__device__ __forceinline__ void streak_blend(uint8_t* p, int channels, const StreakColor& col, float alpha) {
    for (int c = 0; c < channels; ++c) {
        if (alpha >= 1.0f) p[c] = (uint8_t)(int)col.c[c];
        else p[c] = (uint8_t)(int)blend_f32((float)p[c], col.c[c], alpha);
    }
}

__device__ __forceinline__ bool dash_off(int coord, int dash_thickness, int dash_gap) {
    // fill_*_dash_gap (streak.py:24-41): positions with coord % (thickness + gap) < gap are cleared
    if (dash_thickness <= 0 || dash_gap <= 0) return false;
    return (coord % (dash_thickness + dash_gap)) < dash_gap;
}

__global__ void __launch_bounds__(256) streak_line_kernel(uint8_t* image, int h, int w, int channels,
                                                          int thickness, int gap, int dash_thickness,
                                                          int dash_gap, int enable_vert, int enable_hori,
                                                          StreakColor col, float alpha) {
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = blockIdx.y * 8 + threadIdx.y;
    if (x >= w || y >= h) return;
    const int step = thickness + gap;
    const bool vert = enable_vert && (x % step) < thickness && !dash_off(y, dash_thickness, dash_gap);
    const bool hori = enable_hori && (y % step) < thickness && !dash_off(x, dash_thickness, dash_gap);
    uint8_t* p = image + ((long long)y * w + x) * channels;
    if (vert) streak_blend(p, channels, col, alpha);
    if (hori) streak_blend(p, channels, col, alpha);
}

__global__ void __launch_bounds__(256) fill_rects_kernel(uint8_t* mask, int h, int w,
                                                         const vkb_rect* __restrict__ rects) {
    const vkb_rect rc = rects[blockIdx.x];
    const int up = max(rc.up, 0), down = min(rc.down, h - 1);
    const int left = max(rc.left, 0), right = min(rc.right, w - 1);
    const int rw = right - left + 1, rh = down - up + 1;
    if (rw <= 0 || rh <= 0) return;
    for (long long i = threadIdx.x; i < (long long)rw * rh; i += blockDim.x) {
        const int yy = (int)(i / rw), xx = (int)(i - (long long)yy * rw);
        mask[(long long)(up + yy) * w + left + xx] = 1;
    }
}

__global__ void __launch_bounds__(256) streak_masks_kernel(uint8_t* image, int h, int w, int channels,
                                                           const uint8_t* __restrict__ mask_vert,
                                                           const uint8_t* __restrict__ mask_hori,
                                                           int dash_thickness, int dash_gap,
                                                           StreakColor col, float alpha) {
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = blockIdx.y * 8 + threadIdx.y;
    if (x >= w || y >= h) return;
    const long long i = (long long)y * w + x;
    const bool vert = mask_vert && mask_vert[i] && !dash_off(y, dash_thickness, dash_gap);
    const bool hori = mask_hori && mask_hori[i] && !dash_off(x, dash_thickness, dash_gap);
    uint8_t* p = image + i * channels;
    if (vert) streak_blend(p, channels, col, alpha);
    if (hori) streak_blend(p, channels, col, alpha);
}

// ============================================================================================
// cv.filter2D(uint8, -1, float32 kernel), BORDER_REFLECT_101, anchor at the kernel centre:
// float32 accumulation over the taps in row-major order, round half to even, saturate
// (defocus_blur / motion_blur, photometric/blur.py:79-192).  32 x 32 tile + halo in shared
// memory, taps in shared memory, 4 rows per thread.
// ============================================================================================
template <int C>
__global__ void __launch_bounds__(256) filter2d_kernel(const uint8_t* __restrict__ src,
                                                       uint8_t* __restrict__ dst, int h, int w,
                                                       const float* __restrict__ taps, int kh, int kw) {
    extern __shared__ __align__(16) unsigned char smem[];
    float* ktaps = reinterpret_cast<float*>(smem);
    uint8_t* tile = smem + (((size_t)kh * kw * 4 + 15) & ~(size_t)15);
    const int ay = kh / 2, ax = kw / 2;
    const int TW = 32 + kw - 1, TH = 32 + kh - 1;
    const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    for (int i = tid; i < kh * kw; i += 256) ktaps[i] = taps[i];
    for (int i = tid; i < TH * TW; i += 256) {
        const int ty = i / TW, tx = i - ty * TW;
        const int sy = reflect101(y0 + ty - ay, h), sx = reflect101(x0 + tx - ax, w);
        const uint8_t* p = src + ((long long)sy * w + sx) * C;
#pragma unroll
        for (int c = 0; c < C; ++c) tile[i * C + c] = p[c];
    }
    __syncthreads();
    const int x = x0 + threadIdx.x;
    float acc[4][C];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int c = 0; c < C; ++c) acc[j][c] = 0.f;
    for (int ky = 0; ky < kh; ++ky) {
        for (int kx = 0; kx < kw; ++kx) {
            const float f = ktaps[ky * kw + kx];
            if (f == 0.f) continue;  // cv keeps only the non-zero taps
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint8_t* p = tile + ((threadIdx.y + 8 * j + ky) * TW + threadIdx.x + kx) * C;
#pragma unroll
                for (int c = 0; c < C; ++c) acc[j][c] = __fmaf_rn(f, (float)p[c], acc[j][c]);
            }
        }
    }
    if (x >= w) return;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int y = y0 + threadIdx.y + 8 * j;
        if (y >= h) break;
#pragma unroll
        for (int c = 0; c < C; ++c) dst[((long long)y * w + x) * C + c] = (uint8_t)round_u8(acc[j][c]);
    }
}

// ============================================================================================
// cv.resize(uint8): INTER_LINEAR (11-bit fixed-point coefficients, the 8-bit vertical pass
// ((b0*(S0>>4))>>16 + (b1*(S1>>4))>>16 + 2) >> 2) and INTER_NEAREST (floor(x / scale)).
// pixelation (photometric/effect.py:58-79) = linear down + nearest up.  Thread per dst pixel.
// ============================================================================================
// Mask.to_resized_mask (element/mask.py:454-479) fused into the resize kernels: with
// mask_thr >= 0 a source pixel reads as (v > 0) * 255 and the result is stored as (v > mask_thr),
// which removes the two threshold passes around cv.resize.  mask_thr < 0: plain pixels.
__device__ __forceinline__ int px_in(uint8_t v, int mask_thr) {
    return mask_thr >= 0 ? (v ? 255 : 0) : (int)v;
}
__device__ __forceinline__ float px_in(float v, int) { return v; }
__device__ __forceinline__ uint8_t px_out(int v, int mask_thr) {
    return (uint8_t)(mask_thr >= 0 ? (int)(v > mask_thr) : v);
}

// Columns outside the source get fraction 0; rows keep their fraction and clip the two row
// indices (cv2's vertical pass then splits one row over both coefficients, which matters to the
// last bit when the horizontal pass interpolated).
// cv::resize "area mode" (INTER_AREA with an enlarging axis): the bilinear passes with
// sx = floor(dx * scale), fx = (dx + 1) - (sx + 1) * inv_scale, 0 if <= 0, else its fractional part.
__device__ __forceinline__ void resize_area_mode_frac(int d, double scale, int dn, int sn, int& si,
                                                      float& f) {
    const double inv_scale = (double)dn / (double)sn;
    si = (int)floor(__dmul_rn((double)d, scale));
    f = (float)__dsub_rn((double)(d + 1), __dmul_rn((double)(si + 1), inv_scale));
    f = f <= 0.f ? 0.f : __fsub_rn(f, floorf(f));
}

template <bool ROWS>
__device__ __forceinline__ void resize_lin_coef(int d, double scale, int sn, int& s0, int& s1,
                                                int& a0, int& a1, int area_dn = 0) {
    float f;
    int si;
    if (area_dn > 0) {
        resize_area_mode_frac(d, scale, area_dn, sn, si, f);
    } else {
        f = (float)(((double)d + 0.5) * scale - 0.5);
        si = (int)floorf(f);
        f -= (float)si;
    }
    if (!ROWS) {
        if (si < 0) { si = 0; f = 0.f; }
        if (si >= sn - 1) { si = sn - 1; f = 0.f; }
    }
    s0 = min(max(si, 0), sn - 1);
    s1 = min(max(si + 1, 0), sn - 1);
    a0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
    a1 = __float2int_rn(__fmul_rn(f, 2048.f));
}

// cv.INTER_LINEAR_EXACT (bit-exact by design in cv2): double source position, 8-bit coefficients
// rint(frac * 256), 8.8 horizontal pass, vertical pass (v + 2^15) >> 16
__device__ __forceinline__ void resize_lin_exact_coef(int d, double scale, int sn, int& s0, int& s1,
                                                      int& c1) {
    const double pos = ((double)d + 0.5) * scale - 0.5;
    int si = (int)floor(pos);
    double fr = pos - (double)si;
    if (si < 0) { si = 0; fr = 0.0; }
    if (si >= sn - 1) { si = sn - 1; fr = 0.0; }
    s0 = si;
    s1 = min(si + 1, sn - 1);
    c1 = (int)rint(fr * 256.0);
}

template <int C>
__global__ void __launch_bounds__(256) resize_u8_kernel(const uint8_t* __restrict__ src, int sh, int sw,
                                                        uint8_t* __restrict__ dst, int dh, int dw,
                                                        double scale_x, double scale_y, int nearest,
                                                        int mask_thr) {
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = blockIdx.y * 8 + threadIdx.y;
    if (x >= dw || y >= dh) return;
    uint8_t* d = dst + ((long long)y * dw + x) * C;
    if (nearest == 2) {
        // cv.INTER_NEAREST_EXACT: 16.16 fixed point on pixel centres (resizeNN_bitexact)
        const int ifx = ((sw << 16) + dw / 2) / dw, ifx0 = ifx / 2 - (sw % 2);
        const int ify = ((sh << 16) + dh / 2) / dh, ify0 = ify / 2 - (sh % 2);
        const int sx = min((ifx0 + ifx * x) >> 16, sw - 1);
        const int sy = min((ify0 + ify * y) >> 16, sh - 1);
        const uint8_t* p = src + ((long long)sy * sw + sx) * C;
#pragma unroll
        for (int c = 0; c < C; ++c) d[c] = px_out(px_in(p[c], mask_thr), mask_thr);
        return;
    }
    if (nearest == 3) {
        int x0, x1, cx, y0, y1, cy;
        resize_lin_exact_coef(x, scale_x, sw, x0, x1, cx);
        resize_lin_exact_coef(y, scale_y, sh, y0, y1, cy);
        const uint8_t* r0 = src + (long long)y0 * sw * C;
        const uint8_t* r1 = src + (long long)y1 * sw * C;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const int h0 = px_in(r0[x0 * C + c], mask_thr) * (256 - cx) + px_in(r0[x1 * C + c], mask_thr) * cx;
            const int h1 = px_in(r1[x0 * C + c], mask_thr) * (256 - cx) + px_in(r1[x1 * C + c], mask_thr) * cx;
            d[c] = px_out(min(max((h0 * (256 - cy) + h1 * cy + 32768) >> 16, 0), 255), mask_thr);
        }
        return;
    }
    if (nearest == 1) {
        const int sx = min((int)floor((double)x * scale_x), sw - 1);
        const int sy = min((int)floor((double)y * scale_y), sh - 1);
        const uint8_t* p = src + ((long long)sy * sw + sx) * C;
#pragma unroll
        for (int c = 0; c < C; ++c) d[c] = px_out(px_in(p[c], mask_thr), mask_thr);
        return;
    }
    int x0, x1, ax0, ax1, y0, y1, by0, by1;
    const bool area_mode = nearest == 4;  // INTER_AREA with an enlarging axis
    resize_lin_coef<false>(x, scale_x, sw, x0, x1, ax0, ax1, area_mode ? dw : 0);
    resize_lin_coef<true>(y, scale_y, sh, y0, y1, by0, by1, area_mode ? dh : 0);
    const uint8_t* r0 = src + (long long)y0 * sw * C;
    const uint8_t* r1 = src + (long long)y1 * sw * C;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const int s0 = px_in(r0[x0 * C + c], mask_thr) * ax0 + px_in(r0[x1 * C + c], mask_thr) * ax1;
        const int s1 = px_in(r1[x0 * C + c], mask_thr) * ax0 + px_in(r1[x1 * C + c], mask_thr) * ax1;
        const int v = (((by0 * (s0 >> 4)) >> 16) + ((by1 * (s1 >> 4)) >> 16) + 2) >> 2;
        d[c] = px_out(min(max(v, 0), 255), mask_thr);
    }
}

// ============================================================================================
// cv.resize(float32) for ScoreMap.to_resized_score_map (element/score_map.py:616-637): cv2's
// float path -- float32 coefficients, horizontal pass then vertical pass, every product and sum
// rounded to float32 in tap order (no contraction, so the CPU restatement is bit identical),
// replicated borders; `clip01` fuses the np.clip(mat, 0, 1) of probability maps.
// mode 0 NEAREST, 1 LINEAR, 2 CUBIC.
// ============================================================================================
__device__ __forceinline__ void resize_cubic_coef_f64(int d, double scale, int& s0, double* c);

__device__ __forceinline__ void resize_cubic_coef_f32(int d, double scale, int& si, float* c) {
    float f = (float)(((double)d + 0.5) * scale - 0.5);
    si = (int)floorf(f);
    f = __fsub_rn(f, (float)si);
    const float A = -0.75f;
    const float f1 = __fadd_rn(f, 1.f), g = __fsub_rn(1.f, f);
    c[0] = __fsub_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fsub_rn(__fmul_rn(A, f1), 5.f * A), f1), 8.f * A), f1), 4.f * A);
    c[1] = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(A + 2.f, f), A + 3.f), f), f), 1.f);
    c[2] = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(A + 2.f, g), A + 3.f), g), g), 1.f);
    c[3] = __fsub_rn(__fsub_rn(__fsub_rn(1.f, c[0]), c[1]), c[2]);
}

template <int K>
__global__ void __launch_bounds__(256) resize_f32_kernel(const float* __restrict__ src, int sh, int sw,
                                                         float* __restrict__ dst, int dh, int dw,
                                                         double scale_x, double scale_y, int clip01,
                                                         float post) {
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = blockIdx.y * 8 + threadIdx.y;
    if (x >= dw || y >= dh) return;
    float v;
    // Intel IPP (the cv2 wheel's default) serves LINEAR for sources of at least 2 x 2 pixels and
    // CUBIC for 4 x 4: coordinates and taps in double.  Restated in float64, taps in order,
    // rounded to float32 once.  (clip01 & 8 keeps cv2's own float32 path.)
    const bool ipp_form = !(clip01 & (4 | 8)) && ((K == 2 && sh >= 2 && sw >= 2) || (K == 4 && sh >= 4 && sw >= 4));
    if (K > 1 && ipp_form) {
        int x0, y0;
        double cx[4], cy[4];
        if (K == 4) {
            resize_cubic_coef_f64(x, scale_x, x0, cx);
            resize_cubic_coef_f64(y, scale_y, y0, cy);
        } else {
            double fx = __dsub_rn(__dmul_rn((double)x + 0.5, scale_x), 0.5);
            double fl = floor(fx);
            x0 = (int)fl;
            fx = __dsub_rn(fx, fl);
            cx[0] = __dsub_rn(1.0, fx); cx[1] = fx;
            double fy = __dsub_rn(__dmul_rn((double)y + 0.5, scale_y), 0.5);
            fl = floor(fy);
            y0 = (int)fl;
            fy = __dsub_rn(fy, fl);
            cy[0] = __dsub_rn(1.0, fy); cy[1] = fy;
        }
        int xs[K];
#pragma unroll
        for (int i = 0; i < K; ++i) xs[i] = min(max(x0 + i, 0), sw - 1);
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const float* row = src + (long long)min(max(y0 + j, 0), sh - 1) * sw;
            double hsum = __dmul_rn((double)row[xs[0]], cx[0]);
#pragma unroll
            for (int i = 1; i < K; ++i) hsum = __dadd_rn(hsum, __dmul_rn((double)row[xs[i]], cx[i]));
            const double term = __dmul_rn(hsum, cy[j]);
            acc = j ? __dadd_rn(acc, term) : term;
        }
        v = (float)acc;
    } else if (K == 1) {
        int sx, sy;
        if (clip01 & 2) {  // cv.INTER_NEAREST_EXACT: 16.16 fixed point on pixel centres
            const int ifx = ((sw << 16) + dw / 2) / dw, ifx0 = ifx / 2 - (sw % 2);
            const int ify = ((sh << 16) + dh / 2) / dh, ify0 = ify / 2 - (sh % 2);
            sx = min((ifx0 + ifx * x) >> 16, sw - 1);
            sy = min((ify0 + ify * y) >> 16, sh - 1);
        } else {
            sx = min((int)floor((double)x * scale_x), sw - 1);
            sy = min((int)floor((double)y * scale_y), sh - 1);
        }
        v = src[(long long)sy * sw + sx];
    } else {
        int x0, y0;
        float cx[4], cy[4];
        if (K == 4) {
            resize_cubic_coef_f32(x, scale_x, x0, cx);
            resize_cubic_coef_f32(y, scale_y, y0, cy);
            x0 -= 1;
            y0 -= 1;
        } else {
            float fx, fy;
            if (clip01 & 4) {  // INTER_AREA with an enlarging axis
                resize_area_mode_frac(x, scale_x, dw, sw, x0, fx);
                resize_area_mode_frac(y, scale_y, dh, sh, y0, fy);
            } else {
                fx = (float)(((double)x + 0.5) * scale_x - 0.5);
                x0 = (int)floorf(fx);
                fx = __fsub_rn(fx, (float)x0);
                fy = (float)(((double)y + 0.5) * scale_y - 0.5);
                y0 = (int)floorf(fy);
                fy = __fsub_rn(fy, (float)y0);
            }
            if (x0 < 0) { x0 = 0; fx = 0.f; }  // columns: fraction 0 at the border; rows clip
            if (x0 >= sw - 1) { x0 = sw - 1; fx = 0.f; }
            cx[0] = __fsub_rn(1.f, fx); cx[1] = fx;
            cy[0] = __fsub_rn(1.f, fy); cy[1] = fy;
        }
        int xs[K];
#pragma unroll
        for (int i = 0; i < K; ++i) xs[i] = min(max(x0 + i, 0), sw - 1);
        float t[K];
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const float* row = src + (long long)min(max(y0 + j, 0), sh - 1) * sw;
            float hsum = __fmul_rn(row[xs[0]], cx[0]);
#pragma unroll
            for (int i = 1; i < K; ++i) hsum = __fadd_rn(hsum, __fmul_rn(row[xs[i]], cx[i]));
            t[j] = __fmul_rn(hsum, cy[j]);
        }
        // cv2's vector code (4 columns at a time) adds the rows from the last to the first, its
        // scalar tail (the last dw % 4 columns) from the first to the last
        if (x < (dw & ~3)) {
            v = t[K - 1];
#pragma unroll
            for (int j = K - 2; j >= 0; --j) v = __fadd_rn(t[j], v);
        } else {
            v = t[0];
#pragma unroll
            for (int j = 1; j < K; ++j) v = __fadd_rn(v, t[j]);
        }
    }
    if (clip01 & 1) v = fminf(fmaxf(v, 0.f), 1.f);
    dst[(long long)y * dw + x] = __fmul_rn(v, post);  // post = 1: exact no-op
}

// ============================================================================================
// cv.resize(INTER_LANCZOS4), uint8 (C channels) and float32 (one channel): the 8 taps of
// cv::interpolateLanczos4 (double sin / cos of -(frac + 3) pi / 4 rotated by multiples of 45
// degrees, divided by y^2, normalised in float32).  A 32 x 8 block needs 32 column and 8 row tap
// sets: 40 threads compute them into shared memory, then every thread gathers its 8 x 8
// neighbourhood with replicated borders.  uint8: taps rounded to 11 bits, integer passes,
// (sum + 2^21) >> 22 (bit identical to cv2, IPP or not); float32: products and sums rounded to
// float32 in tap order.
// ============================================================================================
__device__ __forceinline__ void lanczos4_taps(float frac, float* c) {
    const double s45 = 0.70710678118654752440084436210485;
    const double rot_s[8] = {1.0, -s45, 0.0, s45, -1.0, s45, 0.0, -s45};
    const double rot_c[8] = {0.0, -s45, 1.0, -s45, 0.0, s45, -1.0, s45};
    const double pi = 3.1415926535897932384626433832795;
    const float x3 = __fadd_rn(frac, 3.f);
    const double y0 = __dmul_rn(__dmul_rn(-(double)x3, pi), 0.25);
    double s0, c0;
    sincos(y0, &s0, &c0);
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float d = __fsub_rn(x3, (float)i);
        if (fabsf(d) >= 1e-6f) {
            const double y = __dmul_rn(__dmul_rn(-(double)d, pi), 0.25);
            const double num = __dadd_rn(__dmul_rn(rot_s[i], s0), __dmul_rn(rot_c[i], c0));
            c[i] = (float)__ddiv_rn(num, __dmul_rn(y, y));
        } else {
            c[i] = 1e30f;
        }
        sum = __fadd_rn(sum, c[i]);
    }
    sum = __fdiv_rn(1.f, sum);
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i] = __fmul_rn(c[i], sum);
}

template <typename T, int C>
__global__ void __launch_bounds__(256) resize_lanczos4_kernel(const T* __restrict__ src, int sh, int sw,
                                                              T* __restrict__ dst, int dh, int dw,
                                                              double scale_x, double scale_y, int clip01,
                                                              int mask_thr, float post) {
    __shared__ float taps[40][8];
    __shared__ int first[40];
    const int tid = threadIdx.y * 32 + threadIdx.x;
    if (tid < 40) {
        const bool col = tid < 32;
        const int d = col ? blockIdx.x * 32 + tid : blockIdx.y * 8 + (tid - 32);
        float f = (float)(((double)d + 0.5) * (col ? scale_x : scale_y) - 0.5);
        const int si = (int)floorf(f);
        f = __fsub_rn(f, (float)si);
        float c[8];
        lanczos4_taps(f, c);
        first[tid] = si - 3;
#pragma unroll
        for (int k = 0; k < 8; ++k) taps[tid][k] = c[k];
    }
    __syncthreads();
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = blockIdx.y * 8 + threadIdx.y;
    if (x >= dw || y >= dh) return;
    const float* cx = taps[threadIdx.x];
    const float* cy = taps[32 + threadIdx.y];
    const int x0 = first[threadIdx.x], y0 = first[32 + threadIdx.y];
    int xs[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) xs[k] = min(max(x0 + k, 0), sw - 1) * C;
    if constexpr (sizeof(T) == 1) {
        int ax[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) ax[k] = __float2int_rn(__fmul_rn(cx[k], 2048.f));
        long long acc[C];
#pragma unroll
        for (int c = 0; c < C; ++c) acc[c] = 0;
#pragma unroll 1
        for (int j = 0; j < 8; ++j) {
            const T* row = src + (long long)min(max(y0 + j, 0), sh - 1) * sw * C;
            const int ay = __float2int_rn(__fmul_rn(cy[j], 2048.f));
#pragma unroll
            for (int c = 0; c < C; ++c) {
                int hsum = 0;
#pragma unroll
                for (int k = 0; k < 8; ++k) hsum += px_in(row[xs[k] + c], mask_thr) * ax[k];
                acc[c] += (long long)hsum * ay;
            }
        }
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const long long v = (acc[c] + (1 << 21)) >> 22;
            dst[((long long)y * dw + x) * C + c] = px_out((int)(v < 0 ? 0 : (v > 255 ? 255 : v)), mask_thr);
        }
    } else {
        float t[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const T* row = src + (long long)min(max(y0 + j, 0), sh - 1) * sw;
            float hsum = __fmul_rn(row[xs[0]], cx[0]);
#pragma unroll
            for (int k = 1; k < 8; ++k) hsum = __fadd_rn(hsum, __fmul_rn(row[xs[k]], cx[k]));
            t[j] = __fmul_rn(hsum, cy[j]);
        }
        // rows added last to first in cv2's vector code, first to last in its scalar tail
        float v;
        if (x < (dw & ~3)) {
            v = t[7];
#pragma unroll
            for (int j = 6; j >= 0; --j) v = __fadd_rn(t[j], v);
        } else {
            v = t[0];
#pragma unroll
            for (int j = 1; j < 8; ++j) v = __fadd_rn(v, t[j]);
        }
        if (clip01) v = fminf(fmaxf(v, 0.f), 1.f);
        dst[(long long)y * dw + x] = __fmul_rn(v, post);
    }
}

// ============================================================================================
// cv.resize(INTER_AREA) when shrinking on both axes (what page_resizing samples), uint8 with C
// channels and float32 with one.  Thread per destination pixel.
//   FAST (integer ratios): box sum; uint8 2x2 -> (sum + 2) >> 2, otherwise
//   round(float(sum) * float(1 / area)); float32 adds in cv2's unrolled order (groups of four in
//   row-major tap order; the vector part of 2x2 rows pairs the two rows first).
//   Otherwise: the weights of cv::computeResizeAreaTab recomputed per pixel in double (head /
//   full / tail source pixels of the cell), horizontal accumulation then vertical accumulation in
//   table order, all in float32 without contraction.
// ============================================================================================
struct AreaSpan {
    int s0, n;
    float a_head, a_mid, a_tail;
    bool has_head, has_tail;
    __device__ __forceinline__ float weight(int i) const {
        return (has_head && i == 0) ? a_head : ((has_tail && i == n - 1) ? a_tail : a_mid);
    }
};

__device__ __forceinline__ AreaSpan area_span(int d, double scale, int ssize) {
    const double f1 = __dmul_rn((double)d, scale);
    const double f2 = __dadd_rn(f1, scale);
    const double cell = fmin(scale, __dsub_rn((double)ssize, f1));
    int s1 = (int)ceil(f1), s2 = (int)floor(f2);
    s2 = min(s2, ssize - 1);
    s1 = min(s1, s2);
    AreaSpan r;
    const double head = __dsub_rn((double)s1, f1), tail = __dsub_rn(f2, (double)s2);
    r.has_head = head > 1e-3;
    r.has_tail = tail > 1e-3;
    r.a_head = (float)__ddiv_rn(head, cell);
    r.a_mid = (float)__ddiv_rn(1.0, cell);
    r.a_tail = (float)__ddiv_rn(fmin(fmin(tail, 1.0), cell), cell);
    r.s0 = s1 - (r.has_head ? 1 : 0);
    r.n = (r.has_head ? 1 : 0) + (s2 - s1) + (r.has_tail ? 1 : 0);
    return r;
}

template <typename T, int C, bool FAST>
__global__ void __launch_bounds__(256) resize_area_kernel(const T* __restrict__ src, int sh, int sw,
                                                          T* __restrict__ dst, int dh, int dw,
                                                          double scale_x, double scale_y, int isx,
                                                          int isy, int clip01, int mask_thr, float post) {
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = blockIdx.y * 8 + threadIdx.y;
    if (x >= dw || y >= dh) return;
    T* d = dst + ((long long)y * dw + x) * C;
    if constexpr (FAST) {
        const int area = isx * isy;
        const T* base = src + ((long long)y * isy * sw + (long long)x * isx) * C;
        if constexpr (sizeof(T) == 1) {
            const float scale = __fdiv_rn(1.f, (float)area);
#pragma unroll
            for (int c = 0; c < C; ++c) {
                int sum = 0;
                for (int ky = 0; ky < isy; ++ky)
                    for (int kx = 0; kx < isx; ++kx) sum += px_in(base[((long long)ky * sw + kx) * C + c], mask_thr);
                const int v = (isx == 2 && isy == 2) ? (sum + 2) >> 2
                                                     : __float2int_rn(__fmul_rn((float)sum, scale));
                d[c] = px_out(min(max(v, 0), 255), mask_thr);
            }
        } else {
            const float scale = __fdiv_rn(1.f, (float)area);
            float v;
            if (isx == 2 && isy == 2 && x < (dw & ~3)) {
                v = __fadd_rn(__fadd_rn(base[0], base[1]), __fadd_rn(base[sw], base[sw + 1]));
            } else {
                v = 0.f;
                int k = 0, ky = 0, kx = 0;
                auto tap = [&]() {
                    const float t = base[(long long)ky * sw + kx];
                    if (++kx == isx) { kx = 0; ++ky; }
                    return t;
                };
                for (; k + 4 <= area; k += 4) {
                    float g = tap();  // one tap per statement: argument evaluation order is unspecified
                    g = __fadd_rn(g, tap());
                    g = __fadd_rn(g, tap());
                    g = __fadd_rn(g, tap());
                    v = k ? __fadd_rn(v, g) : g;
                }
                for (; k < area; ++k) {
                    const float t = tap();
                    v = k ? __fadd_rn(v, t) : t;
                }
            }
            v = __fmul_rn(v, scale);
            if (clip01) v = fminf(fmaxf(v, 0.f), 1.f);
            d[0] = __fmul_rn(v, post);
        }
    } else {
        const AreaSpan xs = area_span(x, scale_x, sw);
        const AreaSpan ys = area_span(y, scale_y, sh);
        float acc[C];
        for (int j = 0; j < ys.n; ++j) {
            const T* row = src + ((long long)(ys.s0 + j) * sw + xs.s0) * C;
            const float beta = ys.weight(j);
#pragma unroll
            for (int c = 0; c < C; ++c) {
                float h = __fmul_rn((float)px_in(row[c], mask_thr), xs.weight(0));
                for (int i = 1; i < xs.n; ++i)
                    h = __fadd_rn(h, __fmul_rn((float)px_in(row[i * C + c], mask_thr), xs.weight(i)));
                const float t = __fmul_rn(beta, h);
                acc[c] = j ? __fadd_rn(acc[c], t) : t;
            }
        }
#pragma unroll
        for (int c = 0; c < C; ++c) {
            if constexpr (sizeof(T) == 1) {
                d[c] = px_out(min(max(__float2int_rn(acc[c]), 0), 255), mask_thr);
            } else {
                float v = acc[c];
                if (clip01) v = fminf(fmaxf(v, 0.f), 1.f);
                d[c] = __fmul_rn(v, post);
            }
        }
    }
}

// cv::resize's choice between the integer-ratio box sum and the weight tables
static bool area_is_fast(double scale_x, double scale_y, int& isx, int& isy) {
    isx = (int)nearbyint(scale_x);
    isy = (int)nearbyint(scale_y);
    return fabs(scale_x - isx) < 2.220446049250313e-16 && fabs(scale_y - isy) < 2.220446049250313e-16;
}

template <typename T, int C>
static void launch_resize_area(const T* src, int sh, int sw, T* dst, int dh, int dw, double scale_x,
                               double scale_y, int clip01, int mask_thr, cudaStream_t st,
                               float post = 1.f) {
    int isx, isy;
    dim3 grid((dw + 31) / 32, (dh + 7) / 8);
    if (area_is_fast(scale_x, scale_y, isx, isy))
        resize_area_kernel<T, C, true><<<grid, dim3(32, 8), 0, st>>>(src, sh, sw, dst, dh, dw, scale_x, scale_y, isx, isy, clip01, mask_thr, post);
    else
        resize_area_kernel<T, C, false><<<grid, dim3(32, 8), 0, st>>>(src, sh, sw, dst, dh, dw, scale_x, scale_y, isx, isy, clip01, mask_thr, post);
}

// ============================================================================================
// mat[pos_y, pos_x]: the pixel permutation of glass_blur (photometric/blur.py:216-264).
// ============================================================================================
template <int C>
__global__ void __launch_bounds__(256) gather_pixels_kernel(const uint8_t* __restrict__ src,
                                                            uint8_t* __restrict__ dst, long long n, int w,
                                                            const int32_t* __restrict__ pos_y,
                                                            const int32_t* __restrict__ pos_x) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t* p = src + ((long long)pos_y[i] * w + pos_x[i]) * C;
#pragma unroll
    for (int c = 0; c < C; ++c) dst[i * C + c] = p[c];
}

// ============================================================================================
// cv.resize(uint8, INTER_CUBIC), sources below 4 x 4 pixels (above that the wheel runs IPP, see
// resize_cubic_pixel_ideal) -- cv2's own path: float32 coefficients (A = -0.75)
// rounded to 11 bits, integer horizontal pass, float32 vertical pass (cv2's vector code: taps
// times 2^-22, summed last to first, rounded half to even), replicated borders.  The cv2 wheel routes cubic through Intel IPP by default, which differs
// from this by +-1 on ~5 % of the pixels of a random image (measured; cv.ipp.setUseIPP(False)
// gives this path).  Used by Image.to_resized_image and zoom_in_blur
// (photometric/blur.py:278-330).
// ============================================================================================
__device__ __forceinline__ void resize_cubic_coef(int d, double scale, int& s0, int* a) {
    float f = (float)(((double)d + 0.5) * scale - 0.5);
    const int si = (int)floorf(f);
    f = __fsub_rn(f, (float)si);
    s0 = si;
    const float A = -0.75f;
    const float f1 = __fadd_rn(f, 1.f), g = __fsub_rn(1.f, f);
    float c[4];
    c[0] = __fsub_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fsub_rn(__fmul_rn(A, f1), 5.f * A), f1), 8.f * A), f1), 4.f * A);
    c[1] = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(A + 2.f, f), A + 3.f), f), f), 1.f);
    c[2] = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(A + 2.f, g), A + 3.f), g), g), 1.f);
    c[3] = __fsub_rn(__fsub_rn(__fsub_rn(1.f, c[0]), c[1]), c[2]);
#pragma unroll
    for (int k = 0; k < 4; ++k) a[k] = min(max(__float2int_rn(__fmul_rn(c[k], 2048.f)), -32768), 32767);
}

// Intel IPP's cubic (what the cv2 wheel runs by default for sources of at least 4 x 4 pixels): the
// bicubic (A = -0.75) without coefficient quantisation.  Restated in float64 -- taps, horizontal
// then vertical sums in tap order, no contraction -- and rounded half to even; the wheel differs
// from this on < 1e-4 of the pixels (+-1 at near ties).
__device__ __forceinline__ void resize_cubic_coef_f64(int d, double scale, int& s0, double* c) {
    double f = __dsub_rn(__dmul_rn((double)d + 0.5, scale), 0.5);
    const double fl = floor(f);
    s0 = (int)fl - 1;
    f = __dsub_rn(f, fl);
    const double f1 = __dadd_rn(f, 1.0), g = __dsub_rn(1.0, f);
    c[0] = __dsub_rn(__dmul_rn(__dadd_rn(__dmul_rn(__dsub_rn(__dmul_rn(-0.75, f1), -3.75), f1), -6.0), f1), -3.0);
    c[1] = __dadd_rn(__dmul_rn(__dmul_rn(__dsub_rn(__dmul_rn(1.25, f), 2.25), f), f), 1.0);
    c[2] = __dadd_rn(__dmul_rn(__dmul_rn(__dsub_rn(__dmul_rn(1.25, g), 2.25), g), g), 1.0);
    c[3] = __dsub_rn(__dsub_rn(__dsub_rn(1.0, c[0]), c[1]), c[2]);
}

template <int C>
__device__ __forceinline__ void resize_cubic_pixel_ideal(const uint8_t* __restrict__ src, int sh,
                                                         int sw, int x, int y, double scale_x,
                                                         double scale_y, int* out, int mask_thr) {
    int sx, sy;
    double ax[4], ay[4];
    resize_cubic_coef_f64(x, scale_x, sx, ax);
    resize_cubic_coef_f64(y, scale_y, sy, ay);
    int xs[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) xs[k] = min(max(sx + k, 0), sw - 1) * C;
    double acc[C];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint8_t* row = src + (long long)min(max(sy + j, 0), sh - 1) * sw * C;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            double hsum = __dmul_rn((double)px_in(row[xs[0] + c], mask_thr), ax[0]);
#pragma unroll
            for (int k = 1; k < 4; ++k)
                hsum = __dadd_rn(hsum, __dmul_rn((double)px_in(row[xs[k] + c], mask_thr), ax[k]));
            const double term = __dmul_rn(hsum, ay[j]);
            acc[c] = j ? __dadd_rn(acc[c], term) : term;
        }
    }
#pragma unroll
    for (int c = 0; c < C; ++c) out[c] = min(max(__double2int_rn(acc[c]), 0), 255);
}

template <int C>
__device__ __forceinline__ void resize_cubic_pixel(const uint8_t* __restrict__ src, int sh, int sw,
                                                   int x, int y, double scale_x, double scale_y,
                                                   int* out, int mask_thr = -1) {
    if (sh >= 4 && sw >= 4) {  // the wheel's IPP path; smaller sources take cv2's own (below)
        resize_cubic_pixel_ideal<C>(src, sh, sw, x, y, scale_x, scale_y, out, mask_thr);
        return;
    }
    int sx, sy, ax[4], ay[4];
    resize_cubic_coef(x, scale_x, sx, ax);
    resize_cubic_coef(y, scale_y, sy, ay);
    // cv2's vertical pass (VResizeCubicVec_32s8u) is float32: taps scaled by 2^-22 (exact), the
    // four products added from the last row to the first without FMA, rounded half to even
    float acc[C];
#pragma unroll
    for (int j = 3; j >= 0; --j) {
        const int yy = min(max(sy - 1 + j, 0), sh - 1);
        const uint8_t* row = src + (long long)yy * sw * C;
        int hsum[C];
#pragma unroll
        for (int c = 0; c < C; ++c) hsum[c] = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int xx = min(max(sx - 1 + k, 0), sw - 1);
#pragma unroll
            for (int c = 0; c < C; ++c) hsum[c] += px_in(row[xx * C + c], mask_thr) * ax[k];
        }
        const float beta = __fmul_rn((float)ay[j], 2.384185791015625e-07f);  // 2^-22
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const float term = __fmul_rn((float)hsum[c], beta);
            acc[c] = j == 3 ? term : __fadd_rn(term, acc[c]);
        }
    }
#pragma unroll
    for (int c = 0; c < C; ++c) out[c] = min(max(__float2int_rn(acc[c]), 0), 255);
}

template <int C>
__global__ void __launch_bounds__(256) resize_cubic_u8_kernel(const uint8_t* __restrict__ src, int sh, int sw,
                                                              uint8_t* __restrict__ dst, int dh, int dw,
                                                              double scale_x, double scale_y, int mask_thr) {
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = blockIdx.y * 8 + threadIdx.y;
    if (x >= dw || y >= dh) return;
    int px[C];
    resize_cubic_pixel<C>(src, sh, sw, x, y, scale_x, scale_y, px, mask_thr);
#pragma unroll
    for (int c = 0; c < C; ++c) dst[((long long)y * dw + x) * C + c] = px_out(px[c], mask_thr);
}

// zoom_in_blur (photometric/blur.py:278-330): the page plus its cubic enlargements (centre
// crops), averaged (uint16 sum / count, rounded half to even), blended with the page in float64
// and truncated.  The enlargements are sampled on the fly, never materialised.
template <int C>
__global__ void __launch_bounds__(256) zoom_in_blur_kernel(const uint8_t* __restrict__ src,
                                                           uint8_t* __restrict__ dst, int h, int w,
                                                           const vkb_zoom_level* __restrict__ levels,
                                                           int n_levels, double alpha) {
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = blockIdx.y * 8 + threadIdx.y;
    if (x >= w || y >= h) return;
    int acc[C], base[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = base[c] = src[((long long)y * w + x) * C + c];
    for (int l = 0; l < n_levels; ++l) {
        const vkb_zoom_level lv = levels[l];
        int px[C];
        resize_cubic_pixel<C>(src, h, w, x + lv.left, y + lv.up, lv.scale_x, lv.scale_y, px);
#pragma unroll
        for (int c = 0; c < C; ++c) acc[c] += px[c];
    }
    const double count = (double)(n_levels + 1);
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const double mean = rint(__ddiv_rn((double)acc[c], count));
        const double v = __dadd_rn(__dmul_rn(1.0 - alpha, (double)base[c]), __dmul_rn(alpha, mean));
        dst[((long long)y * w + x) * C + c] = (uint8_t)(int)fmin(fmax(v, 0.0), 255.0);
    }
}

// dst = src > threshold ? high : low (Mask.to_resized_mask: element/mask.py:454-479)
__global__ void __launch_bounds__(256) threshold_u8_kernel(const uint8_t* __restrict__ src,
                                                           uint8_t* __restrict__ dst, long long n,
                                                           int threshold, int low, int high) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = (uint8_t)((int)src[i] > threshold ? high : low);
}

// ============================================================================================
// Batched photometric chain: Gaussian blur (optional) followed by a per-pixel op list, one pass
// over a ragged batch of pages (per-page shapes, taps and op lists).  The chained form of
// gaussian_blur -> color_shift / brightness_shift / mean_shift / ... as RandomDistortion applies
// them (distortion_policy/random_distortion.py:350-392), without the intermediate image.
//
// Block = 256 threads on a 32 x 32 output tile; the interleaved channels are treated as a byte
// stream (a horizontal tap is C bytes away).
//   load   tile + halo -> shared memory, rows re-aligned to the tile's first byte (interior
//          tiles: aligned 32-bit loads + funnel shift; border tiles: BORDER_REFLECT_101 bytes);
//   horiz  a thread blurs 4 consecutive bytes of two vertically adjacent rows: per tap one
//          funnel shift + two PRMT unpack the bytes into 16-bit lane pairs and one IMAD per pair
//          accumulates both lanes at once (the taps sum to 256, so a lane never exceeds
//          255 * 256 and cannot carry into its neighbour); results are re-paired vertically
//          (row 2t, row 2t+1) per byte and stored as 32-bit words;
//   vert   a thread owns one pixel of two adjacent output rows: per channel R+1 word loads and
//          IDP.2A (16-bit pair x byte weights) per row, + 2^15, >> 16;
//   ops    the op list runs on the blurred pixel in registers; bytes are stored.
// ============================================================================================
template <int C, int R>
struct ChainGeom {
    static constexpr int TH = 32 + 2 * R;                    // tile rows incl. halo (even)
    static constexpr int TWB = (32 + 2 * R) * C;             // tile row bytes incl. halo
    static constexpr int ROW_WORDS = (TWB + 3) / 4 + 1;      // + 1: the horizontal pass may read one word past
    static constexpr int NB = 32 * C;                        // output bytes per tile row
    static constexpr int TILE_BYTES = (TH * ROW_WORDS * 4 + 15) & ~15;  // V stays 16-byte aligned
    static constexpr int V_BYTES = (TH / 2) * NB * 4;
    static constexpr int SMEM = TILE_BYTES + V_BYTES;
};

template <int C, int R>
__device__ __forceinline__ void hblur4(const uint32_t* __restrict__ w, const int* __restrict__ taps,
                                       uint32_t& lo, uint32_t& hi) {
    lo = 0;
    hi = 0;
#pragma unroll
    for (int k = 0; k <= 2 * R; ++k) {
        const int o = C * k, wi = o >> 2, sh = o & 3;
        const uint32_t win = sh ? __funnelshift_r(w[wi], w[wi + 1], 8 * sh) : w[wi];
        lo += (uint32_t)taps[k] * __byte_perm(win, 0u, 0x4140);
        hi += (uint32_t)taps[k] * __byte_perm(win, 0u, 0x4342);
    }
}

template <int C, int R, bool POS>
__device__ __forceinline__ void chain_tile(const vkb_photo_page& pg, unsigned char* smem,
                                           const int* __restrict__ taps,
                                           const uint32_t* __restrict__ wev,
                                           const uint32_t* __restrict__ wod,
                                           const HsvTables& tables) {
    using G = ChainGeom<C, R>;
    const int h = pg.h, w = pg.w;
    const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;
    const uint8_t* __restrict__ src = pg.src;
    uint8_t* __restrict__ dst = pg.dst;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    const int lane = threadIdx.x, warp = threadIdx.y;
    uint32_t* tile = reinterpret_cast<uint32_t*>(smem);                        // TH x ROW_WORDS
    uint32_t* V = reinterpret_cast<uint32_t*>(smem + G::TILE_BYTES);           // TH/2 x NB

    // ---- load ---------------------------------------------------------------------------
    const bool interior = x0 - R >= 0 && y0 - R >= 0 && x0 + 32 + R <= w && y0 + 32 + R <= h;
    if (interior) {
        // Every word that is read holds at least one byte of the span, so the reads stay inside
        // the plane's allocation (device allocations are sized in multiples of >= 16 bytes).
        const size_t pitch = (size_t)w * C;
        uintptr_t a = reinterpret_cast<uintptr_t>(src) + ((size_t)(y0 - R) * w + (x0 - R)) * C
                      + warp * pitch;
        for (int ty = warp; ty < G::TH; ty += 8, a += 8 * pitch) {
            const uint32_t* __restrict__ a0 = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
            const int mis = (int)(a & 3);
            const int last = (mis + G::TWB - 1) >> 2;  // last word with span bytes
            for (int k = lane; k < G::ROW_WORDS - 1; k += 32) {
                uint32_t v = __ldg(a0 + k);
                if (mis) {
                    const uint32_t nxt = k < last ? __ldg(a0 + k + 1) : 0u;
                    v = __funnelshift_r(v, nxt, 8 * mis);
                }
                tile[ty * G::ROW_WORDS + k] = v;
            }
        }
    } else {
        uint8_t* tb = reinterpret_cast<uint8_t*>(tile);
        constexpr int TW = 32 + 2 * R;
        for (int i = tid; i < G::TH * TW; i += 256) {
            const int ty = i / TW, tx = i - ty * TW;
            const int sy = reflect101(y0 + ty - R, h), sx = reflect101(x0 + tx - R, w);
            const uint8_t* p = src + ((long long)sy * w + sx) * C;
#pragma unroll
            for (int c = 0; c < C; ++c) tb[ty * (G::ROW_WORDS * 4) + tx * C + c] = p[c];
        }
    }
    __syncthreads();

    // ---- horizontal: (row pair t, 4-byte group q) ------------------------------------------
    constexpr int NQ = G::NB / 4;
    constexpr int NW = 1 + (2 * R * C + 3) / 4;  // words a 4-byte group reads per row
    for (int i = tid; i < (G::TH / 2) * NQ; i += 256) {
        const int t = i / NQ, q = i - t * NQ;
        uint32_t w0[NW + 1], w1[NW + 1];
        const uint32_t* r0 = tile + (2 * t) * G::ROW_WORDS + q;
        const uint32_t* r1 = r0 + G::ROW_WORDS;
#pragma unroll
        for (int j = 0; j < NW; ++j) {
            w0[j] = r0[j];
            w1[j] = r1[j];
        }
        w0[NW] = 0;
        w1[NW] = 0;
        uint32_t lo0, hi0, lo1, hi1;
        hblur4<C, R>(w0, taps, lo0, hi0);
        hblur4<C, R>(w1, taps, lo1, hi1);
        uint4 out;
        out.x = __byte_perm(lo0, lo1, 0x5410);  // byte 0: (row 2t, row 2t+1)
        out.y = __byte_perm(lo0, lo1, 0x7632);
        out.z = __byte_perm(hi0, hi1, 0x5410);
        out.w = __byte_perm(hi0, hi1, 0x7632);
        *reinterpret_cast<uint4*>(V + t * G::NB + 4 * q) = out;
    }
    __syncthreads();

    // ---- vertical + ops + store: (pixel x, output row pair) --------------------------------
    const int n_ops = pg.n_ops;
    const int x = x0 + lane;
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        const int op_ = warp + 8 * it;  // output row pair 0..15
        const int ya = y0 + 2 * op_;
        int pa[4] = {0, 0, 0, 0}, pb[4] = {0, 0, 0, 0};
#pragma unroll
        for (int c = 0; c < C; ++c) {
            uint32_t acc_a = 1u << 15, acc_b = 1u << 15;
#pragma unroll
            for (int m = 0; m <= R; ++m) {
                const uint32_t v = V[(op_ + m) * G::NB + lane * C + c];
                acc_a = __dp2a_lo(v, wev[m], acc_a);
                acc_b = __dp2a_lo(v, wod[m], acc_b);
            }
            pa[c] = (int)(acc_a >> 16);
            pb[c] = (int)(acc_b >> 16);
        }
        if (x < w && ya < h) {
            for (int k = 0; k < n_ops; ++k)
                apply_color_op<POS>(pg.ops[k], pa, C, tables, x, ya, (long long)ya * w + x);
            uint8_t* d = dst + ((long long)ya * w + x) * C;
#pragma unroll
            for (int c = 0; c < C; ++c) d[c] = (uint8_t)pa[c];
            if (ya + 1 < h) {
                for (int k = 0; k < n_ops; ++k)
                    apply_color_op<POS>(pg.ops[k], pb, C, tables, x, ya + 1, (long long)(ya + 1) * w + x);
                d += (long long)w * C;
#pragma unroll
                for (int c = 0; c < C; ++c) d[c] = (uint8_t)pb[c];
            }
        }
    }
}

template <int C, bool POS>
__global__ void __launch_bounds__(256) photo_chain_kernel(const vkb_photo_page* __restrict__ pages) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ int taps[17];
    __shared__ uint32_t wev[9], wod[9];
    __shared__ HsvTables tables;
    const vkb_photo_page& pg = pages[blockIdx.z];
    const int h = pg.h, w = pg.w;
    const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;
    if (x0 >= w || y0 >= h) return;
    const int r = pg.blur_radius;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    if (pg.n_ops > 0) stage_hsv_tables(tables, tid, 256);
    if (r == 0) __syncthreads();

    if (r == 0) {
        const uint8_t* __restrict__ src = pg.src;
        uint8_t* __restrict__ dst = pg.dst;
        const int n_ops = pg.n_ops;
        const int x = x0 + threadIdx.x;
        if (x >= w) return;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int y = y0 + threadIdx.y + 8 * j;
            if (y >= h) break;
            const long long i = (long long)y * w + x;
            int px[4] = {0, 0, 0, 0};
#pragma unroll
            for (int c = 0; c < C; ++c) px[c] = src[i * C + c];
            for (int k = 0; k < n_ops; ++k) apply_color_op<POS>(pg.ops[k], px, C, tables, x, y, i);
#pragma unroll
            for (int c = 0; c < C; ++c) dst[i * C + c] = (uint8_t)px[c];
        }
        return;
    }
    // taps, and the byte-weight pairs of the vertical pass: output row 2o reads row pairs
    // m = 0..R with weights (w[2m], w[2m+1]); output row 2o+1 with (w[2m-1], w[2m]).
    if (tid <= 2 * r) taps[tid] = pg.blur_taps[tid];
    if (tid >= 32 && tid < 32 + 9) {
        const int m = tid - 32;
        auto tap = [&](int k) { return (k >= 0 && k <= 2 * r) ? (uint32_t)pg.blur_taps[k] : 0u; };
        wev[m] = tap(2 * m) | (tap(2 * m + 1) << 8);
        wod[m] = tap(2 * m - 1) | (tap(2 * m) << 8);
    }
    __syncthreads();
    switch (r) {
        case 1: chain_tile<C, 1, POS>(pg, smem, taps, wev, wod, tables); break;
        case 2: chain_tile<C, 2, POS>(pg, smem, taps, wev, wod, tables); break;
        case 3: chain_tile<C, 3, POS>(pg, smem, taps, wev, wod, tables); break;
        case 4: chain_tile<C, 4, POS>(pg, smem, taps, wev, wod, tables); break;
        case 5: chain_tile<C, 5, POS>(pg, smem, taps, wev, wod, tables); break;
        case 6: chain_tile<C, 6, POS>(pg, smem, taps, wev, wod, tables); break;
        case 7: chain_tile<C, 7, POS>(pg, smem, taps, wev, wod, tables); break;
        default: chain_tile<C, 8, POS>(pg, smem, taps, wev, wod, tables); break;
    }
}

template <int C>
static size_t chain_smem_bytes(int r) {
    switch (r) {
        case 0: return 0;
        case 1: return ChainGeom<C, 1>::SMEM;
        case 2: return ChainGeom<C, 2>::SMEM;
        case 3: return ChainGeom<C, 3>::SMEM;
        case 4: return ChainGeom<C, 4>::SMEM;
        case 5: return ChainGeom<C, 5>::SMEM;
        case 6: return ChainGeom<C, 6>::SMEM;
        case 7: return ChainGeom<C, 7>::SMEM;
        default: return ChainGeom<C, 8>::SMEM;
    }
}

// noise over a ragged batch: ops[0] of every page carries VKB_OP_NOISE (or the page is skipped)
__global__ void __launch_bounds__(256) noise_philox_batched_kernel(const vkb_photo_page* __restrict__ pages,
                                                                   int channels) {
    const vkb_photo_page& pg = pages[blockIdx.y];
    if (pg.n_ops < 1 || pg.ops[0].kind != VKB_OP_NOISE) return;
    const long long n = (long long)pg.h * pg.w;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        int px[4] = {0, 0, 0, 0};
        for (int c = 0; c < channels; ++c) px[c] = pg.src[i * channels + c];
        noise_op(pg.ops[0], px, channels, i);
        for (int c = 0; c < channels; ++c) pg.dst[i * channels + c] = (uint8_t)px[c];
    }
}

// per-page channel sums / mins / maxs of a ragged batch (std_shift, boundary_equalization)
__global__ void __launch_bounds__(256) channel_stats_batched_kernel(
    const vkb_photo_page* __restrict__ pages, int channels, unsigned long long* __restrict__ out) {
    const vkb_photo_page& pg = pages[blockIdx.y];
    const long long n = (long long)pg.h * pg.w;
    unsigned long long s[3] = {0, 0, 0};
    unsigned int mn[3] = {255, 255, 255}, mx[3] = {0, 0, 0};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        for (int c = 0; c < channels; ++c) {
            const unsigned int v = pg.src[i * channels + c];
            s[c] += v;
            mn[c] = min(mn[c], v);
            mx[c] = max(mx[c], v);
        }
    }
    // out per page: 3 sums (u64), then 3 mins and 3 maxs packed as u32 pairs in 3 more u64 slots
    unsigned long long* o = out + (size_t)blockIdx.y * 6;
    unsigned int* o32 = reinterpret_cast<unsigned int*>(o + 3);
    for (int c = 0; c < channels; ++c) {
        unsigned long long v = s[c];
        unsigned int a = mn[c], b = mx[c];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            v += __shfl_xor_sync(0xffffffffu, v, d);
            a = min(a, __shfl_xor_sync(0xffffffffu, a, d));
            b = max(b, __shfl_xor_sync(0xffffffffu, b, d));
        }
        if ((threadIdx.x & 31) == 0) {
            atomicAdd(o + c, v);
            atomicMin(o32 + c, a);
            atomicMax(o32 + 3 + c, b);
        }
    }
}

__global__ void channel_stats_batched_init_kernel(unsigned long long* out, int n_pages) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pages) return;
    unsigned long long* o = out + (size_t)i * 6;
    unsigned int* o32 = reinterpret_cast<unsigned int*>(o + 3);
    o[0] = o[1] = o[2] = 0;
    o32[0] = o32[1] = o32[2] = 0xffffffffu;
    o32[3] = o32[4] = o32[5] = 0;
}

}  // namespace vkb

// ============================================================================================
// C ABI
// ============================================================================================
using namespace vkb;

extern "C" int vkb_blend_fill(const vkb_blend_item* item_host, void* stream) {
    VKB_REQUIRE(item_host && item_host->dst, "no destination");
    const vkb_blend_item& it = *item_host;
    VKB_REQUIRE(it.channels >= 1 && it.channels <= 4, "channels must be 1..4");
    if (it.box_h <= 0 || it.box_w <= 0) return VKB_OK;
    dim3 grid((it.box_w + 31) / 32, (it.box_h + 31) / 32);
    blend_fill_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(it);
    return check_launch("blend_fill_kernel");
}

extern "C" int vkb_blend_draw_list(const vkb_blend_item* items, int32_t n_items, int32_t dst_h,
                                   int32_t dst_w, void* stream) {
    VKB_NVTX("vkb_blend_draw_list");
    VKB_REQUIRE(items && n_items >= 0 && dst_h > 0 && dst_w > 0, "bad arguments");
    if (n_items == 0) return VKB_OK;
    dim3 grid((dst_w + 31) / 32, (dst_h + 31) / 32);
    blend_draw_list_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(items, n_items, dst_h, dst_w);
    return check_launch("blend_draw_list_kernel");
}

extern "C" int vkb_cvt_color(const uint8_t* src, uint8_t* dst, int64_t n_pixels, int32_t code,
                             void* stream) {
    VKB_REQUIRE(src && dst && code >= 0 && code <= VKB_CVT_RGBA2GRAY, "bad arguments");
    if (n_pixels <= 0) return VKB_OK;
    int rc = ensure_tables((cudaStream_t)stream);
    if (rc) return rc;
    const long long blocks = (n_pixels + 255) / 256;
    cvt_color_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, dst, n_pixels, code);
    return check_launch("cvt_color_kernel");
}

extern "C" int vkb_color_ops(const uint8_t* src, uint8_t* dst, int64_t n_pixels, int32_t channels,
                             const vkb_color_op* ops_host, int32_t n_ops, void* stream) {
    VKB_NVTX("vkb_color_ops");
    VKB_REQUIRE(src && dst && ops_host, "bad arguments");
    VKB_REQUIRE(n_ops >= 0 && n_ops <= VKB_MAX_COLOR_OPS, "too many ops");
    VKB_REQUIRE(channels == 1 || channels == 3 || channels == 4, "channels must be 1, 3 or 4");
    if (n_pixels <= 0) return VKB_OK;
    int rc = ensure_tables((cudaStream_t)stream);
    if (rc) return rc;
    ColorOpList list;
    list.n = n_ops;
    for (int i = 0; i < n_ops; ++i) list.ops[i] = ops_host[i];
    const unsigned blocks = (unsigned)((n_pixels + 255) / 256);
    cudaStream_t st = (cudaStream_t)stream;
    if (channels == 1) color_ops_kernel<1><<<blocks, 256, 0, st>>>(src, dst, n_pixels, list);
    else if (channels == 3) color_ops_kernel<3><<<blocks, 256, 0, st>>>(src, dst, n_pixels, list);
    else color_ops_kernel<4><<<blocks, 256, 0, st>>>(src, dst, n_pixels, list);
    return check_launch("color_ops_kernel");
}

extern "C" int vkb_channel_stats(const uint8_t* src, int64_t n_pixels, int32_t channels, void* out,
                                 void* stream) {
    VKB_REQUIRE(src && out && channels >= 1 && channels <= 3, "bad arguments");
    unsigned long long* sums = reinterpret_cast<unsigned long long*>(out);
    unsigned int* mins = reinterpret_cast<unsigned int*>(sums + 3);
    unsigned int* maxs = mins + 3;
    cudaStream_t st = (cudaStream_t)stream;
    channel_stats_init_kernel<<<1, 32, 0, st>>>(sums, mins, maxs);
    if (n_pixels > 0) {
        const long long want = (n_pixels + 255) / 256;
        const unsigned blocks = (unsigned)(want < 148 * 8 ? want : 148 * 8);
        channel_stats_kernel<<<blocks, 256, 0, st>>>(src, n_pixels, channels, sums, mins, maxs);
    }
    return check_launch("channel_stats_kernel");
}

extern "C" int vkb_histogram_u8(const uint8_t* src, int64_t n_pixels, int32_t channels,
                                uint32_t* out, void* stream) {
    VKB_REQUIRE(src && out && channels >= 1 && channels <= 3, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    VKB_CUDA(cudaMemsetAsync(out, 0, sizeof(uint32_t) * 3 * 256, st));
    if (n_pixels > 0) {
        const long long want = (n_pixels + 255) / 256;
        const unsigned blocks = (unsigned)(want < 148 * 8 ? want : 148 * 8);
        histogram_kernel<<<blocks, 256, 0, st>>>(src, n_pixels, channels, out);
    }
    return check_launch("histogram_kernel");
}

extern "C" int vkb_apply_lut(const uint8_t* src, uint8_t* dst, int64_t n_pixels, int32_t channels,
                             const uint8_t* lut, int32_t channel_bits, void* stream) {
    VKB_REQUIRE(src && dst && lut, "bad arguments");
    VKB_REQUIRE(channels == 1 || channels == 3 || channels == 4, "channels must be 1, 3 or 4");
    if (n_pixels <= 0) return VKB_OK;
    const unsigned blocks = (unsigned)((n_pixels + 255) / 256);
    cudaStream_t st = (cudaStream_t)stream;
    if (channels == 1) apply_lut_kernel<1><<<blocks, 256, 0, st>>>(src, dst, n_pixels, lut, channel_bits);
    else if (channels == 3) apply_lut_kernel<3><<<blocks, 256, 0, st>>>(src, dst, n_pixels, lut, channel_bits);
    else apply_lut_kernel<4><<<blocks, 256, 0, st>>>(src, dst, n_pixels, lut, channel_bits);
    return check_launch("apply_lut_kernel");
}

extern "C" int vkb_gaussian_blur_u8(const uint8_t* src, uint8_t* dst, int32_t h, int32_t w,
                                    int32_t channels, const int32_t* kernel_host, int32_t ksize,
                                    void* stream) {
    VKB_NVTX("vkb_gaussian_blur_u8");
    VKB_REQUIRE(src && dst && kernel_host && h > 0 && w > 0, "bad arguments");
    VKB_REQUIRE(ksize >= 1 && (ksize & 1) && ksize <= 2 * kGaussMaxR + 1, "ksize must be odd and <= 17");
    VKB_REQUIRE(channels == 1 || channels == 3 || channels == 4, "channels must be 1, 3 or 4");
    GaussKernel gk;
    gk.r = ksize / 2;
    for (int i = 0; i < ksize; ++i) gk.k[i] = kernel_host[i];
    cudaStream_t st = (cudaStream_t)stream;
    const int tma = gaussian_blur_tma(src, dst, h, w, channels, gk, st);
    if (tma != 0) return tma > 0 ? VKB_OK : tma;
    const int TW = 32 + 2 * gk.r;
    const size_t smem = ((size_t)(TW * TW * channels + 3) & ~(size_t)3) + (size_t)TW * 32 * channels * 2;
    dim3 grid((w + 31) / 32, (h + 31) / 32);
    if (channels == 1) gaussian_blur_kernel<1><<<grid, dim3(32, 8), smem, st>>>(src, dst, h, w, gk);
    else if (channels == 3) gaussian_blur_kernel<3><<<grid, dim3(32, 8), smem, st>>>(src, dst, h, w, gk);
    else gaussian_blur_kernel<4><<<grid, dim3(32, 8), smem, st>>>(src, dst, h, w, gk);
    return check_launch("gaussian_blur_kernel");
}

extern "C" int vkb_noise_philox(const uint8_t* src, uint8_t* dst, int64_t n_pixels, int32_t channels,
                                int32_t kind, double p0, double p1, uint64_t seed, void* stream) {
    VKB_REQUIRE(src && dst && kind >= 0 && kind <= 3 && channels >= 1 && channels <= 4, "bad arguments");
    if (n_pixels <= 0) return VKB_OK;
    // parameters in float32 precision, like the op records of the batched form: a page gets the
    // same noise alone or in a batch
    p0 = (double)(float)p0;
    p1 = (double)(float)p1;
    noise_philox_kernel<<<(unsigned)((n_pixels + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        src, dst, n_pixels, channels, kind, p0, p1, seed);
    return check_launch("noise_philox_kernel");
}

extern "C" int vkb_noise_field(const uint8_t* src, uint8_t* dst, int64_t n_pixels, int32_t channels,
                               int32_t kind, const void* field, void* stream) {
    VKB_REQUIRE(src && dst && field && kind >= 0 && kind <= 3 && channels >= 1 && channels <= 4,
                "bad arguments");
    if (n_pixels <= 0) return VKB_OK;
    noise_field_kernel<<<(unsigned)((n_pixels + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        src, dst, n_pixels, channels, kind, field);
    return check_launch("noise_field_kernel");
}

extern "C" int vkb_streak_line(uint8_t* image, int32_t h, int32_t w, int32_t channels,
                               int32_t thickness, int32_t gap, int32_t dash_thickness,
                               int32_t dash_gap, int32_t enable_vert, int32_t enable_hori,
                               const float* color_host, float alpha, void* stream) {
    VKB_REQUIRE(image && color_host && h > 0 && w > 0 && thickness + gap > 0, "bad arguments");
    StreakColor col;
    for (int c = 0; c < 4; ++c) col.c[c] = c < channels ? color_host[c] : 0.f;
    dim3 grid((w + 31) / 32, (h + 7) / 8);
    streak_line_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(
        image, h, w, channels, thickness, gap, dash_thickness, dash_gap, enable_vert, enable_hori, col, alpha);
    return check_launch("streak_line_kernel");
}

extern "C" int vkb_fill_rects(uint8_t* mask, int32_t h, int32_t w, const vkb_rect* rects,
                              int32_t n_rects, void* stream) {
    VKB_REQUIRE(mask && h > 0 && w > 0, "bad arguments");
    if (n_rects <= 0) return VKB_OK;
    fill_rects_kernel<<<n_rects, 256, 0, (cudaStream_t)stream>>>(mask, h, w, rects);
    return check_launch("fill_rects_kernel");
}

extern "C" int vkb_streak_masks(uint8_t* image, int32_t h, int32_t w, int32_t channels,
                                const uint8_t* mask_vert, const uint8_t* mask_hori,
                                int32_t dash_thickness, int32_t dash_gap, const float* color_host,
                                float alpha, void* stream) {
    VKB_REQUIRE(image && color_host && h > 0 && w > 0, "bad arguments");
    StreakColor col;
    for (int c = 0; c < 4; ++c) col.c[c] = c < channels ? color_host[c] : 0.f;
    dim3 grid((w + 31) / 32, (h + 7) / 8);
    streak_masks_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(
        image, h, w, channels, mask_vert, mask_hori, dash_thickness, dash_gap, col, alpha);
    return check_launch("streak_masks_kernel");
}

extern "C" int vkb_photo_chain_batched(const vkb_photo_page* pages, const vkb_photo_page* pages_host,
                                       int32_t n_pages, int32_t channels, void* stream) {
    VKB_NVTX("vkb_photo_chain_batched");
    VKB_REQUIRE(pages && pages_host && n_pages > 0 && n_pages <= 65535, "bad arguments");
    VKB_REQUIRE(channels == 1 || channels == 3 || channels == 4, "channels must be 1, 3 or 4");
    int max_h = 0, max_w = 0, max_r = 0;
    bool pos = false;  // some page carries a position dependent op
    for (int i = 0; i < n_pages; ++i) {
        const vkb_photo_page& p = pages_host[i];
        VKB_REQUIRE(p.src && p.dst && p.h > 0 && p.w > 0, "page without planes");
        VKB_REQUIRE(p.blur_radius >= 0 && p.blur_radius <= 8, "blur radius must be 0..8");
        VKB_REQUIRE(p.n_ops >= 0 && p.n_ops <= VKB_MAX_COLOR_OPS, "too many ops");
        VKB_REQUIRE(p.blur_radius == 0 || p.src != p.dst, "blur cannot run in place");
        for (int k = 0; k <= 2 * p.blur_radius && p.blur_radius > 0; ++k)
            VKB_REQUIRE(p.blur_taps[k] >= 0 && p.blur_taps[k] <= 255,
                        "blur taps must be 0..255 (a 256 centre tap is the identity: use radius 0)");
        max_h = p.h > max_h ? p.h : max_h;
        max_w = p.w > max_w ? p.w : max_w;
        max_r = p.blur_radius > max_r ? p.blur_radius : max_r;
        for (int k = 0; k < p.n_ops; ++k) {
            VKB_REQUIRE(p.ops[k].kind != VKB_OP_NOISE, "noise runs through vkb_noise_philox_batched");
            pos = pos || p.ops[k].kind == VKB_OP_LINE_STREAK;
        }
    }
    cudaStream_t st = (cudaStream_t)stream;
    int rc = ensure_tables(st);
    if (rc) return rc;
    dim3 grid((max_w + 31) / 32, (max_h + 31) / 32, n_pages);
    VKB_REQUIRE(grid.y <= 65535, "page too tall");
#define VKB_LAUNCH_CHAIN(CH)                                                                          \
    do {                                                                                              \
        if (pos) photo_chain_kernel<CH, true><<<grid, dim3(32, 8), chain_smem_bytes<CH>(max_r), st>>>(pages); \
        else photo_chain_kernel<CH, false><<<grid, dim3(32, 8), chain_smem_bytes<CH>(max_r), st>>>(pages);    \
    } while (0)
    if (channels == 1) VKB_LAUNCH_CHAIN(1);
    else if (channels == 3) VKB_LAUNCH_CHAIN(3);
    else VKB_LAUNCH_CHAIN(4);
#undef VKB_LAUNCH_CHAIN
    return check_launch("photo_chain_kernel");
}

extern "C" int vkb_channel_stats_batched(const vkb_photo_page* pages, int32_t n_pages,
                                         int32_t channels, void* out, void* stream) {
    VKB_REQUIRE(pages && out && n_pages > 0 && n_pages <= 65535, "bad arguments");
    VKB_REQUIRE(channels >= 1 && channels <= 3, "channels must be 1..3");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long* o = reinterpret_cast<unsigned long long*>(out);
    channel_stats_batched_init_kernel<<<(n_pages + 255) / 256, 256, 0, st>>>(o, n_pages);
    channel_stats_batched_kernel<<<dim3(64, n_pages), 256, 0, st>>>(pages, channels, o);
    return check_launch("channel_stats_batched_kernel");
}

extern "C" int vkb_filter2d_u8(const uint8_t* src, uint8_t* dst, int32_t h, int32_t w,
                               int32_t channels, const float* taps_dev, int32_t kh, int32_t kw,
                               void* stream) {
    VKB_REQUIRE(src && dst && taps_dev && h > 0 && w > 0 && src != dst, "bad arguments");
    VKB_REQUIRE(kh >= 1 && kw >= 1 && (kh & 1) && (kw & 1) && kh <= 63 && kw <= 63,
                "kernel sides must be odd and <= 63");
    VKB_REQUIRE(channels == 1 || channels == 3 || channels == 4, "channels must be 1, 3 or 4");
    const size_t smem = (((size_t)kh * kw * 4 + 15) & ~(size_t)15)
                        + (size_t)(32 + kw - 1) * (32 + kh - 1) * channels;
    dim3 grid((w + 31) / 32, (h + 31) / 32);
    cudaStream_t st = (cudaStream_t)stream;
#define VKB_LAUNCH_F2D(CH)                                                                       \
    do {                                                                                         \
        if (smem > 48 * 1024)                                                                    \
            VKB_CUDA(cudaFuncSetAttribute(filter2d_kernel<CH>,                                   \
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        filter2d_kernel<CH><<<grid, dim3(32, 8), smem, st>>>(src, dst, h, w, taps_dev, kh, kw);  \
    } while (0)
    if (channels == 1) VKB_LAUNCH_F2D(1);
    else if (channels == 3) VKB_LAUNCH_F2D(3);
    else VKB_LAUNCH_F2D(4);
#undef VKB_LAUNCH_F2D
    return check_launch("filter2d_kernel");
}

static int resize_u8_impl(const uint8_t* src, int32_t src_h, int32_t src_w, uint8_t* dst,
                          int32_t dst_h, int32_t dst_w, int32_t channels, int32_t interpolation,
                          int mask_thr, void* stream) {
    VKB_REQUIRE(src && dst && src_h > 0 && src_w > 0 && dst_h > 0 && dst_w > 0, "bad arguments");
    VKB_REQUIRE(channels == 1 || channels == 3 || channels == 4, "channels must be 1, 3 or 4");
    VKB_REQUIRE(interpolation == VKB_INTER_NEAREST || interpolation == VKB_INTER_LINEAR
                    || interpolation == VKB_INTER_CUBIC || interpolation == VKB_INTER_AREA
                    || interpolation == VKB_INTER_LANCZOS4 || interpolation == VKB_INTER_LINEAR_EXACT
                    || interpolation == VKB_INTER_NEAREST_EXACT,
                "interpolation must be VKB_INTER_NEAREST / LINEAR / CUBIC / AREA / LANCZOS4 / LINEAR_EXACT / NEAREST_EXACT");
    VKB_REQUIRE(src_h < 32768 && src_w < 32768 && dst_h < 32768 && dst_w < 32768,
                "planes of at most 32767 pixels per side");
    // cv::resize: inv_scale = dsize / ssize, scale = 1 / inv_scale (both double)
    const double scale_x = 1.0 / ((double)dst_w / (double)src_w);
    const double scale_y = 1.0 / ((double)dst_h / (double)src_h);
    dim3 grid((dst_w + 31) / 32, (dst_h + 7) / 8);
    const dim3 block(32, 8);
    cudaStream_t st = (cudaStream_t)stream;
#define VKB_BY_CHANNELS(LAUNCH)             \
    do {                                    \
        if (channels == 1) { LAUNCH(1); }   \
        else if (channels == 3) { LAUNCH(3); } \
        else { LAUNCH(4); }                 \
    } while (0)
    if (interpolation == VKB_INTER_CUBIC) {
#define VKB_L(CH) resize_cubic_u8_kernel<CH><<<grid, block, 0, st>>>(src, src_h, src_w, dst, dst_h, dst_w, scale_x, scale_y, mask_thr)
        VKB_BY_CHANNELS(VKB_L);
#undef VKB_L
        return check_launch("resize_cubic_u8_kernel");
    }
    if (interpolation == VKB_INTER_AREA && dst_h <= src_h && dst_w <= src_w) {
#define VKB_L(CH) launch_resize_area<uint8_t, CH>(src, src_h, src_w, dst, dst_h, dst_w, scale_x, scale_y, 0, mask_thr, st)
        VKB_BY_CHANNELS(VKB_L);
#undef VKB_L
        return check_launch("resize_area_kernel");
    }
    if (interpolation == VKB_INTER_LANCZOS4) {
#define VKB_L(CH) resize_lanczos4_kernel<uint8_t, CH><<<grid, block, 0, st>>>(src, src_h, src_w, dst, dst_h, dst_w, scale_x, scale_y, 0, mask_thr, 1.f)
        VKB_BY_CHANNELS(VKB_L);
#undef VKB_L
        return check_launch("resize_lanczos4_kernel");
    }
    const int nearest = interpolation == VKB_INTER_NEAREST ? 1
                        : interpolation == VKB_INTER_NEAREST_EXACT ? 2
                        : interpolation == VKB_INTER_LINEAR_EXACT ? 3
                        : interpolation == VKB_INTER_AREA ? 4 : 0;  // 4: an enlarging axis
#define VKB_L(CH) resize_u8_kernel<CH><<<grid, block, 0, st>>>(src, src_h, src_w, dst, dst_h, dst_w, scale_x, scale_y, nearest, mask_thr)
    VKB_BY_CHANNELS(VKB_L);
#undef VKB_L
#undef VKB_BY_CHANNELS
    return check_launch("resize_u8_kernel");
}

extern "C" int vkb_resize_u8(const uint8_t* src, int32_t src_h, int32_t src_w, uint8_t* dst,
                             int32_t dst_h, int32_t dst_w, int32_t channels, int32_t interpolation,
                             void* stream) {
    return resize_u8_impl(src, src_h, src_w, dst, dst_h, dst_w, channels, interpolation, -1, stream);
}

extern "C" int vkb_resize_mask_u8(const uint8_t* src, int32_t src_h, int32_t src_w, uint8_t* dst,
                                  int32_t dst_h, int32_t dst_w, int32_t interpolation,
                                  int32_t binarization_threshold, void* stream) {
    VKB_REQUIRE(binarization_threshold >= 0 && binarization_threshold <= 255,
                "binarization_threshold must be in 0..255");
    return resize_u8_impl(src, src_h, src_w, dst, dst_h, dst_w, 1, interpolation,
                          binarization_threshold, stream);
}

extern "C" int vkb_resize_f32_scaled(const float* src, int32_t src_h, int32_t src_w, float* dst,
                                     int32_t dst_h, int32_t dst_w, int32_t interpolation,
                                     int32_t clip01, float post_scale, void* stream) {
    VKB_REQUIRE(src && dst && src_h > 0 && src_w > 0 && dst_h > 0 && dst_w > 0, "bad arguments");
    VKB_REQUIRE(interpolation == VKB_INTER_NEAREST || interpolation == VKB_INTER_LINEAR
                    || interpolation == VKB_INTER_CUBIC || interpolation == VKB_INTER_AREA
                    || interpolation == VKB_INTER_LANCZOS4 || interpolation == VKB_INTER_LINEAR_EXACT
                    || interpolation == VKB_INTER_NEAREST_EXACT,
                "interpolation must be VKB_INTER_NEAREST / LINEAR / CUBIC / AREA / LANCZOS4 / LINEAR_EXACT / NEAREST_EXACT");
    VKB_REQUIRE(src_h < 32768 && src_w < 32768 && dst_h < 32768 && dst_w < 32768,
                "planes of at most 32767 pixels per side");
    const double scale_x = 1.0 / ((double)dst_w / (double)src_w);
    const double scale_y = 1.0 / ((double)dst_h / (double)src_h);
    dim3 grid((dst_w + 31) / 32, (dst_h + 7) / 8);
    cudaStream_t st = (cudaStream_t)stream;
    const int clip = clip01 ? 1 : 0;
    if (interpolation == VKB_INTER_NEAREST || interpolation == VKB_INTER_NEAREST_EXACT)
        resize_f32_kernel<1><<<grid, dim3(32, 8), 0, st>>>(
            src, src_h, src_w, dst, dst_h, dst_w, scale_x, scale_y,
            clip | (interpolation == VKB_INTER_NEAREST_EXACT ? 2 : 0), post_scale);
    else if (interpolation == VKB_INTER_LINEAR || interpolation == VKB_INTER_LINEAR_EXACT)
        // cv::resize runs INTER_LINEAR for float data when INTER_LINEAR_EXACT is asked for
        resize_f32_kernel<2><<<grid, dim3(32, 8), 0, st>>>(src, src_h, src_w, dst, dst_h, dst_w, scale_x, scale_y, clip, post_scale);
    else if (interpolation == VKB_INTER_CUBIC)
        resize_f32_kernel<4><<<grid, dim3(32, 8), 0, st>>>(src, src_h, src_w, dst, dst_h, dst_w, scale_x, scale_y, clip, post_scale);
    else if (interpolation == VKB_INTER_AREA && !(dst_h <= src_h && dst_w <= src_w))
        // an enlarging axis: the bilinear passes with "area mode" fractions
        resize_f32_kernel<2><<<grid, dim3(32, 8), 0, st>>>(src, src_h, src_w, dst, dst_h, dst_w, scale_x, scale_y, clip | 4, post_scale);
    else if (interpolation == VKB_INTER_AREA)
        launch_resize_area<float, 1>(src, src_h, src_w, dst, dst_h, dst_w, scale_x, scale_y, clip, -1, st, post_scale);
    else
        resize_lanczos4_kernel<float, 1><<<grid, dim3(32, 8), 0, st>>>(src, src_h, src_w, dst, dst_h, dst_w, scale_x, scale_y, clip, -1, post_scale);
    return check_launch("resize_f32_kernel");
}

extern "C" int vkb_resize_f32(const float* src, int32_t src_h, int32_t src_w, float* dst,
                              int32_t dst_h, int32_t dst_w, int32_t interpolation, int32_t clip01,
                              void* stream) {
    return vkb_resize_f32_scaled(src, src_h, src_w, dst, dst_h, dst_w, interpolation, clip01, 1.f,
                                 stream);
}

extern "C" int vkb_gather_pixels_u8(const uint8_t* src, uint8_t* dst, int32_t h, int32_t w,
                                    int32_t channels, const int32_t* pos_y, const int32_t* pos_x,
                                    void* stream) {
    VKB_REQUIRE(src && dst && pos_y && pos_x && h > 0 && w > 0 && src != dst, "bad arguments");
    VKB_REQUIRE(channels == 1 || channels == 3 || channels == 4, "channels must be 1, 3 or 4");
    const long long n = (long long)h * w;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    cudaStream_t st = (cudaStream_t)stream;
    if (channels == 1) gather_pixels_kernel<1><<<blocks, 256, 0, st>>>(src, dst, n, w, pos_y, pos_x);
    else if (channels == 3) gather_pixels_kernel<3><<<blocks, 256, 0, st>>>(src, dst, n, w, pos_y, pos_x);
    else gather_pixels_kernel<4><<<blocks, 256, 0, st>>>(src, dst, n, w, pos_y, pos_x);
    return check_launch("gather_pixels_kernel");
}

extern "C" int vkb_noise_philox_batched(const vkb_photo_page* pages, int32_t n_pages,
                                        int32_t channels, int32_t blocks_per_page, void* stream) {
    VKB_REQUIRE(pages && n_pages > 0 && n_pages <= 65535, "bad arguments");
    VKB_REQUIRE(channels >= 1 && channels <= 4 && blocks_per_page > 0, "bad arguments");
    noise_philox_batched_kernel<<<dim3(blocks_per_page, n_pages), 256, 0, (cudaStream_t)stream>>>(
        pages, channels);
    return check_launch("noise_philox_batched_kernel");
}

extern "C" int vkb_zoom_in_blur_u8(const uint8_t* src, uint8_t* dst, int32_t h, int32_t w,
                                   int32_t channels, const vkb_zoom_level* levels_dev,
                                   int32_t n_levels, double alpha, void* stream) {
    VKB_REQUIRE(src && dst && src != dst && h > 0 && w > 0 && n_levels >= 0, "bad arguments");
    VKB_REQUIRE(n_levels == 0 || levels_dev, "levels missing");
    VKB_REQUIRE(n_levels < 255, "at most 254 enlargements (the reference sums in uint16)");
    VKB_REQUIRE(channels == 1 || channels == 3 || channels == 4, "channels must be 1, 3 or 4");
    dim3 grid((w + 31) / 32, (h + 7) / 8);
    cudaStream_t st = (cudaStream_t)stream;
    if (channels == 1) zoom_in_blur_kernel<1><<<grid, dim3(32, 8), 0, st>>>(src, dst, h, w, levels_dev, n_levels, alpha);
    else if (channels == 3) zoom_in_blur_kernel<3><<<grid, dim3(32, 8), 0, st>>>(src, dst, h, w, levels_dev, n_levels, alpha);
    else zoom_in_blur_kernel<4><<<grid, dim3(32, 8), 0, st>>>(src, dst, h, w, levels_dev, n_levels, alpha);
    return check_launch("zoom_in_blur_kernel");
}

extern "C" int vkb_threshold_u8(const uint8_t* src, uint8_t* dst, int64_t n, int32_t threshold,
                                int32_t low, int32_t high, void* stream) {
    VKB_REQUIRE(src && dst && n >= 0, "bad arguments");
    if (n == 0) return VKB_OK;
    threshold_u8_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        src, dst, n, threshold, low, high);
    return check_launch("threshold_u8_kernel");
}
