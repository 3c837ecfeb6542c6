// compose.cu -- the step before text-layer compositing:
//
//   * background synthesis: ImageCombinerEngine.synthesize_image
//     (vkit/engine/image/combiner.py:178-333) pastes texture segments into a canvas one
//     NumPy slice at a time, marks a band around every segment border in an edge mask, blurs the
//     WHOLE canvas with cv.GaussianBlur and copies the blurred pixels back under the mask.
//     Here the page is produced in ONE pass: every 32 x 32 output tile resolves which segment
//     owns each of its (tile + halo) pixels straight from the segment table, reads those pixels
//     from the device-resident textures, evaluates the edge bands analytically and runs the
//     8.8 fixed-point Gaussian only in tiles a band crosses.  No zero canvas, no mask plane, no
//     blurred copy: 3 B/px written, 3 B/px (+ halo) read.
//   * glyph atlas: FreeType coverage bitmaps uploaded once become the three planes the text-line
//     renderer blends from (engine/font/freetype.py:136-221, 314-380): the glyph mask
//     (`bitmap > 0`, any channel for LCD glyphs), the gamma-corrected float32 alpha
//     (`np.power(bitmap / 255, gamma)`, a 256-entry table evaluated by the caller with NumPy so
//     the values are the reference's own) and, for LCD glyphs, the inverted uint8 image
//     (`((1 - np.power(bitmap / 255.0, gamma)) * 255).astype(uint8)`, again a table).
#include "common.cuh"

namespace vkb {

constexpr int kComposeMaxR = 8;
constexpr int kComposeMaxItems = 4096;

struct ComposeTaps {
    int k[2 * kComposeMaxR + 1];
    int r;
};

__device__ __forceinline__ int compose_reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
    return i;
}

// Is (y, x) inside one of the four bands fill_np_edge_mask draws for this segment
// (combiner.py:147-176)?  Rows of the up / down bands span [left, right] only, columns of the
// left / right bands span [up, down] only; both are clipped to the canvas by the caller's range.
__device__ __forceinline__ bool in_edge_band(const vkb_paste_item& it, int y, int x, int g) {
    const bool cols = x >= it.left && x <= it.right;
    const bool rows = y >= it.up && y <= it.down;
    const bool near_up = y >= it.up - g && y <= it.up + g;
    const bool near_down = y >= it.down - g && y <= it.down + g;
    const bool near_left = x >= it.left - g && x <= it.left + g;
    const bool near_right = x >= it.right - g && x <= it.right + g;
    return (cols && (near_up || near_down)) || (rows && (near_left || near_right));
}

template <int C>
__global__ void __launch_bounds__(256) background_compose_kernel(
    uint8_t* __restrict__ dst, int h, int w, const vkb_paste_item* __restrict__ items, int n_items,
    int band, const ComposeTaps taps) {
    extern __shared__ unsigned char smem[];
    __shared__ uint32_t cand[kComposeMaxItems / 32];
    __shared__ int any_band;
    const int r = taps.r;
    const int TW = 32 + 2 * r, TH = 32 + 2 * r;
    uint8_t* tile = smem;  // TH x TW x C
    unsigned short* rows = reinterpret_cast<unsigned short*>(smem + ((TH * TW * C + 3) & ~3));  // TH x 32 x C
    const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    const int n_words = (n_items + 31) >> 5;

    for (int i = tid; i < n_words; i += 256) cand[i] = 0u;
    if (tid == 0) any_band = 0;
    __syncthreads();
    // segments whose rectangle grown by the band (>= the halo the blur needs of them is covered by
    // the plain rectangle test below) meets the tile grown by the halo
    const int reach = max(band, r);
    for (int i = tid; i < n_items; i += 256) {
        const vkb_paste_item it = items[i];
        if (it.left - reach <= x0 + 31 + r && it.right + reach >= x0 - r &&
            it.up - reach <= y0 + 31 + r && it.down + reach >= y0 - r)
            atomicOr(&cand[i >> 5], 1u << (i & 31));
    }
    __syncthreads();

    // stage tile + halo: the LAST segment containing the pixel owns it (later slices overwrite)
    for (int i = tid; i < TH * TW; i += 256) {
        const int ty = i / TW, tx = i - ty * TW;
        const int sy = compose_reflect101(y0 + ty - r, h), sx = compose_reflect101(x0 + tx - r, w);
        const uint8_t* p = nullptr;
        for (int wd = n_words - 1; wd >= 0 && !p; --wd) {
            uint32_t bits = cand[wd];
            while (bits) {
                const int b = 31 - __clz(bits);
                bits &= ~(1u << b);
                const vkb_paste_item& it = items[wd * 32 + b];
                if (sy >= it.up && sy <= it.down && sx >= it.left && sx <= it.right) {
                    p = it.src + ((long long)(sy - it.up) * it.src_pitch + (sx - it.left)) * C;
                    break;
                }
            }
        }
#pragma unroll
        for (int c = 0; c < C; ++c) tile[i * C + c] = p ? __ldg(p + c) : (uint8_t)0;
    }
    // which of the tile's own pixels sit in an edge band
    bool banded[4] = {false, false, false, false};
    {
        const int x = x0 + threadIdx.x;
        for (int wd = 0; wd < n_words; ++wd) {
            uint32_t bits = cand[wd];
            while (bits) {
                const int b = __ffs(bits) - 1;
                bits &= bits - 1;
                const vkb_paste_item it = items[wd * 32 + b];
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    banded[j] = banded[j] || in_edge_band(it, y0 + threadIdx.y + 8 * j, x, band);
            }
        }
        if (banded[0] || banded[1] || banded[2] || banded[3]) any_band = 1;  // benign race: all store 1
    }
    __syncthreads();
    const bool blur_tile = any_band != 0;
    if (blur_tile) {
        for (int i = tid; i < TH * 32; i += 256) {
            const int ty = i >> 5, tx = i & 31;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                int acc = 0;
                for (int k = 0; k <= 2 * r; ++k) acc += (int)tile[(ty * TW + tx + k) * C + c] * taps.k[k];
                rows[i * C + c] = (unsigned short)min(acc, 65535);
            }
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
            int out = tile[((ly + r) * TW + threadIdx.x + r) * C + c];
            if (banded[j]) {
                int acc = 0;
                for (int k = 0; k <= 2 * r; ++k)
                    acc += (int)rows[((ly + k) * 32 + threadIdx.x) * C + c] * taps.k[k];
                out = min((acc + (1 << 15)) >> 16, 255);
            }
            dst[((long long)y * w + x) * C + c] = (uint8_t)out;
        }
    }
}

// One block per glyph.
__global__ void __launch_bounds__(256) glyph_prepare_kernel(const vkb_glyph_item* __restrict__ items) {
    const vkb_glyph_item it = items[blockIdx.x];
    const int C = it.channels;
    for (int i = threadIdx.x; i < it.n_pixels; i += blockDim.x) {
        uint8_t v[3];
        bool on = false;
        for (int c = 0; c < C; ++c) {
            v[c] = it.bitmap[(long long)i * C + c];
            on = on || v[c] > 0;
        }
        it.mask[i] = on ? 1 : 0;
        if (it.alpha) it.alpha[i] = it.alpha_lut[v[0]];
        if (it.lcd_image)
            for (int c = 0; c < C; ++c) it.lcd_image[(long long)i * C + c] = it.lcd_lut[v[c]];
    }
}

}  // namespace vkb

using namespace vkb;

extern "C" int vkb_background_compose(uint8_t* dst, int32_t h, int32_t w, int32_t channels,
                                      const vkb_paste_item* items, int32_t n_items, int32_t band,
                                      const int32_t* kernel_host, int32_t ksize, void* stream) {
    VKB_NVTX("vkb_background_compose");
    VKB_REQUIRE(dst && items && kernel_host && h > 0 && w > 0, "bad arguments");
    VKB_REQUIRE(n_items >= 0 && n_items <= kComposeMaxItems, "at most 4096 segments per canvas");
    VKB_REQUIRE(ksize >= 1 && (ksize & 1) && ksize <= 2 * kComposeMaxR + 1, "ksize must be odd and <= 17");
    VKB_REQUIRE(band >= 0, "band must be >= 0");
    VKB_REQUIRE(channels == 1 || channels == 3 || channels == 4, "channels must be 1, 3 or 4");
    ComposeTaps taps;
    taps.r = ksize / 2;
    for (int i = 0; i < ksize; ++i) taps.k[i] = kernel_host[i];
    const int TW = 32 + 2 * taps.r;
    const size_t smem = ((size_t)(TW * TW * channels + 3) & ~(size_t)3) + (size_t)TW * 32 * channels * 2;
    dim3 grid((w + 31) / 32, (h + 31) / 32);
    cudaStream_t st = (cudaStream_t)stream;
    if (channels == 1)
        background_compose_kernel<1><<<grid, dim3(32, 8), smem, st>>>(dst, h, w, items, n_items, band, taps);
    else if (channels == 3)
        background_compose_kernel<3><<<grid, dim3(32, 8), smem, st>>>(dst, h, w, items, n_items, band, taps);
    else
        background_compose_kernel<4><<<grid, dim3(32, 8), smem, st>>>(dst, h, w, items, n_items, band, taps);
    return check_launch("background_compose_kernel");
}

extern "C" int vkb_glyph_prepare(const vkb_glyph_item* items, const vkb_glyph_item* items_host,
                                 int32_t n_items, void* stream) {
    VKB_NVTX("vkb_glyph_prepare");
    VKB_REQUIRE(items && items_host && n_items >= 0, "bad arguments");
    for (int i = 0; i < n_items; ++i) {
        const vkb_glyph_item& it = items_host[i];
        VKB_REQUIRE(it.bitmap && it.mask && it.n_pixels >= 0, "glyph item without bitmap / mask");
        VKB_REQUIRE(it.channels == 1 || it.channels == 3, "glyph bitmaps are H x W or H x W x 3");
        VKB_REQUIRE(!it.alpha || it.alpha_lut, "alpha plane without its table");
        VKB_REQUIRE(!it.lcd_image || it.lcd_lut, "LCD image without its table");
    }
    if (n_items == 0) return VKB_OK;
    glyph_prepare_kernel<<<n_items, 256, 0, (cudaStream_t)stream>>>(items);
    return check_launch("glyph_prepare_kernel");
}
