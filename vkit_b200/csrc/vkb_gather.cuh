// vkb_gather.cuh -- cv::remap's fixed-point bilinear gather (INTER_LINEAR, BORDER_CONSTANT 0) as the
// warp / remap kernels issue it: tap weights shared by all containers of a pixel, aligned 32-bit row
// loads for packed RGB, out-of-image footprints fixed behind one branch.
#pragma once
#include <stdint.h>
#include "vkb_math.cuh"

namespace vkb {

// ---- bilinear taps ---------------------------------------------------------------------------
// (sum p*w + 2^14) >> 15 with w = (32-fy)(32-fx)*32 ...: all weights share the factor 32, so it
// equals (gy*a + fy*b + 512) >> 10 with a, b the horizontally blended rows (gx*p0 + fx*p1).

// One row of an RGB pixel pair: the 6 bytes at `p` (any alignment) as two words
// lo = R0 G0 B0 R1, hi = G1 B1 . .  (2 aligned 32-bit loads, a third when the bytes straddle).
struct RowRgb {
    uint32_t w0, w1, w2, off;
};

__device__ __forceinline__ RowRgb row_rgb_load(const uint8_t* __restrict__ p) {
    RowRgb r;
    const uintptr_t addr = reinterpret_cast<uintptr_t>(p);
    r.off = (uint32_t)addr & 3u;
    const uint32_t* __restrict__ q = reinterpret_cast<const uint32_t*>(addr & ~(uintptr_t)3);
    r.w0 = __ldg(q);
    r.w1 = __ldg(q + 1);
    r.w2 = 0;
    if (r.off == 3u) r.w2 = __ldg(q + 2);  // only then do the 6 bytes reach into a third word
    return r;
}

// horizontal blend with IDP.4A: wr / wg0,wg1 / wb0,wb1 are byte-weight words on (lo, hi)
__device__ __forceinline__ void row_rgb_blend(const RowRgb& row, uint32_t wr, uint32_t wg0,
                                              uint32_t wg1, uint32_t wb0, uint32_t wb1, int& r,
                                              int& g, int& b) {
    const uint32_t sel = 0x3210u + row.off * 0x1111u;
    const uint32_t lo = __byte_perm(row.w0, row.w1, sel);
    const uint32_t hi = __byte_perm(row.w1, row.w2, sel);
    r = (int)__dp4a(lo, wr, 0u);
    g = (int)__dp4a(hi, wg1, __dp4a(lo, wg0, 0u));
    b = (int)__dp4a(hi, wb1, __dp4a(lo, wb0, 0u));
}

// Branch-free taps.  The 2 x 2 footprint at (x0, y0) is read from the in-image window that
// starts at xs = clamp(x0, 0, w-2), ys = clamp(y0, 0, h-2); taps that fall outside the image
// (BORDER_CONSTANT 0) get weight 0 and the surviving tap keeps its own weight:
//   x0 == xs: (32-fx, fx)   x0 == xs-1: (fx, 0)   x0 == xs+1: (0, 32-fx)   otherwise (0, 0).
// No divergent border branch, so the loads of all pixels of a thread can be in flight together.
// Needs w >= 2 and h >= 2 (the kernel routes smaller planes to sample_u8_small).
struct TapWeights {
    int xs, ys;
    int wx0, wx1, wy0, wy1;
};

__device__ __forceinline__ TapWeights tap_weights_plain(int X, int Y) {
    TapWeights t;
    t.xs = X >> kInterBits;
    t.ys = Y >> kInterBits;
    t.wx1 = X & (kInterTab - 1);
    t.wy1 = Y & (kInterTab - 1);
    t.wx0 = kInterTab - t.wx1;
    t.wy0 = kInterTab - t.wy1;
    return t;
}

// true when the 2 x 2 footprint leaves the image (sizes are below 32768, checked by the caller,
// so cv's saturate_cast<short> of the integer coordinates cannot turn an outside tap into an
// inside one)
__device__ __forceinline__ bool tap_outside(const TapWeights& t, int h, int w) {
    return (unsigned)t.xs > (unsigned)(w - 2) || (unsigned)t.ys > (unsigned)(h - 2);
}

__device__ __forceinline__ void tap_border_fix(TapWeights& t, int h, int w) {
    const int x0 = t.xs, y0 = t.ys, fx = t.wx1, fy = t.wy1;
    t.xs = min(max(x0, 0), w - 2);
    t.ys = min(max(y0, 0), h - 2);
    const int dx = x0 - t.xs, dy = y0 - t.ys;
    t.wx0 = dx == 0 ? kInterTab - fx : (dx == -1 ? fx : 0);
    t.wx1 = dx == 0 ? fx : (dx == 1 ? kInterTab - fx : 0);
    t.wy0 = dy == 0 ? kInterTab - fy : (dy == -1 ? fy : 0);
    t.wy1 = dy == 0 ? fy : (dy == 1 ? kInterTab - fy : 0);
}


// Taps of one pixel, requested now and blended later (so a thread keeps the loads of all its
// pixels in flight).
template <int C>
struct Taps {
    TapWeights t;
    RowRgb rgb[2];                // C == 3
    int p[C == 3 ? 1 : 4 * C];    // C != 3: p00, p01, p10, p11 per channel
};

template <int C>
__device__ __forceinline__ void taps_load(const uint8_t* __restrict__ src, int w, Taps<C>& k) {
    const int pitch = w * C;  // a page plane is < 2 GiB (checked by the caller)
    const uint8_t* r0 = src + (k.t.ys * pitch + k.t.xs * C);
    const uint8_t* r1 = r0 + pitch;
    if (C == 3) {
        k.rgb[0] = row_rgb_load(r0);
        k.rgb[1] = row_rgb_load(r1);
    } else {
#pragma unroll
        for (int c = 0; c < C; ++c) {
            k.p[(4 * c + 0) % (C == 3 ? 1 : 4 * C)] = __ldg(r0 + c);
            k.p[(4 * c + 1) % (C == 3 ? 1 : 4 * C)] = __ldg(r0 + C + c);
            k.p[(4 * c + 2) % (C == 3 ? 1 : 4 * C)] = __ldg(r1 + c);
            k.p[(4 * c + 3) % (C == 3 ? 1 : 4 * C)] = __ldg(r1 + C + c);
        }
    }
}

template <int C>
__device__ __forceinline__ void taps_blend(const Taps<C>& k, uint8_t* __restrict__ out) {
    const TapWeights& t = k.t;
    if (C == 3) {
        const uint32_t wr = (uint32_t)t.wx0 | ((uint32_t)t.wx1 << 24);
        const uint32_t wg0 = (uint32_t)t.wx0 << 8, wg1 = (uint32_t)t.wx1;
        const uint32_t wb0 = (uint32_t)t.wx0 << 16, wb1 = (uint32_t)t.wx1 << 8;
        int a[3], b[3];
        row_rgb_blend(k.rgb[0], wr, wg0, wg1, wb0, wb1, a[0], a[1], a[2]);
        row_rgb_blend(k.rgb[1], wr, wg0, wg1, wb0, wb1, b[0], b[1], b[2]);
#pragma unroll
        for (int c = 0; c < 3; ++c) out[c % C] = (uint8_t)((t.wy0 * a[c] + t.wy1 * b[c] + 512) >> 10);
    } else {
#pragma unroll
        for (int c = 0; c < C; ++c) {
            constexpr int M = C == 3 ? 1 : 4 * C;
            const int a = t.wx0 * k.p[(4 * c + 0) % M] + t.wx1 * k.p[(4 * c + 1) % M];
            const int b = t.wx0 * k.p[(4 * c + 2) % M] + t.wx1 * k.p[(4 * c + 3) % M];
            out[c] = (uint8_t)((t.wy0 * a + t.wy1 * b + 512) >> 10);
        }
    }
}

// ---- second form: both blend directions folded into IDP.2A ------------------------------------
// out = (sum_ij wy_i * wx_j * p_ij + 512) >> 10.  The four weight products (<= 1024 each) travel
// as two 16-bit pairs, w_row0 = wy0*wx0 | wy0*wx1 << 16 and w_row1 likewise (one IMAD each from
// wx0 | wx1 << 16).  A channel's two horizontal neighbours are brought side by side with PRMT and
// one IDP.2A adds both products of a row to the accumulator: two IDP.2A per channel and pixel, no
// separate vertical pass.  Same integer result as the two-pass form above.
struct Tap2 {
    int xs, ys;               // top-left tap, inside the image
    uint32_t w_row0, w_row1;  // packed weight products of the two rows
};

// in-image footprint: plain weights
__device__ __forceinline__ Tap2 tap2_plain(int X, int Y) {
    Tap2 t;
    t.xs = X >> kInterBits;
    t.ys = Y >> kInterBits;
    const int fx = X & (kInterTab - 1), fy = Y & (kInterTab - 1);
    const uint32_t wx = (uint32_t)kInterTab + (uint32_t)fx * 0xFFFFu;  // (32 - fx) | fx << 16
    t.w_row0 = (uint32_t)(kInterTab - fy) * wx;
    t.w_row1 = (uint32_t)fy * wx;
    return t;
}

// true when the 2 x 2 footprint of a plain tap leaves the image (h, w >= 2)
__device__ __forceinline__ bool tap2_outside(const Tap2& t, int h, int w) {
    return (unsigned)t.xs > (unsigned)(w - 2) || (unsigned)t.ys > (unsigned)(h - 2);
}

// footprint that leaves the image: read the in-image window at (clamp(x0), clamp(y0)); taps
// outside (BORDER_CONSTANT 0) get weight 0, the surviving tap keeps its own weight
__device__ __forceinline__ Tap2 tap2_border(int X, int Y, int h, int w) {
    Tap2 t;
    const int x0 = X >> kInterBits, y0 = Y >> kInterBits;
    const int fx = X & (kInterTab - 1), fy = Y & (kInterTab - 1);
    t.xs = min(max(x0, 0), w - 2);
    t.ys = min(max(y0, 0), h - 2);
    const int dx = x0 - t.xs, dy = y0 - t.ys;
    const int wx0 = dx == 0 ? kInterTab - fx : (dx == -1 ? fx : 0);
    const int wx1 = dx == 0 ? fx : (dx == 1 ? kInterTab - fx : 0);
    const uint32_t wx = (uint32_t)wx0 | ((uint32_t)wx1 << 16);
    const int wy0 = dy == 0 ? kInterTab - fy : (dy == -1 ? fy : 0);
    const int wy1 = dy == 0 ? fy : (dy == 1 ? kInterTab - fy : 0);
    t.w_row0 = (uint32_t)wy0 * wx;
    t.w_row1 = (uint32_t)wy1 * wx;
    return t;
}

// h, w >= 2 (smaller planes go to sample_u8_small)
__device__ __forceinline__ Tap2 tap2_make(int X, int Y, int h, int w) {
    Tap2 t = tap2_plain(X, Y);
    if (tap2_outside(t, h, w)) t = tap2_border(X, Y, h, w);  // rare
    return t;
}

// Requested taps of one pixel (loads in flight), blended later.
//   C == 3: per row the aligned words that hold the 6 bytes R0 G0 B0 R1 G1 B1 (+ their offset);
//   C == 4: per row the two pixels as words;   C == 1: per row the two bytes.
template <int C>
struct Fetch2 {
    uint32_t a[C == 3 ? 4 : 2];  // row 0: w0, w1, (w2, off)
    uint32_t b[C == 3 ? 4 : 2];  // row 1
};

// `base` = the plane's base address rounded DOWN to 4 bytes, `mis` = the bytes dropped
// (base & 3): byte offsets below are relative to the rounded base.  pitch in bytes.
template <int C>
__device__ __forceinline__ void fetch2_request(const uint32_t* __restrict__ base, int mis, int pitch,
                                               const Tap2& t, Fetch2<C>& f) {
    const int o0 = t.ys * pitch + (t.xs * C + mis);
    const int o1 = o0 + pitch;
    if (C == 3) {
        const uint32_t* __restrict__ q0 = base + (o0 >> 2);
        const uint32_t* __restrict__ q1 = base + (o1 >> 2);
        f.a[3 % (C == 3 ? 4 : 2)] = ((uint32_t)o0 & 3u) * 8u;
        f.b[3 % (C == 3 ? 4 : 2)] = ((uint32_t)o1 & 3u) * 8u;
        f.a[0] = __ldg(q0);
        f.a[1] = __ldg(q0 + 1);
        f.b[0] = __ldg(q1);
        f.b[1] = __ldg(q1 + 1);
        f.a[2 % (C == 3 ? 4 : 2)] = 0u;
        f.b[2 % (C == 3 ? 4 : 2)] = 0u;
        // only an offset of 3 pushes the 6 bytes into a third word
        if (((uint32_t)o0 & 3u) == 3u) f.a[2 % (C == 3 ? 4 : 2)] = __ldg(q0 + 2);
        if (((uint32_t)o1 & 3u) == 3u) f.b[2 % (C == 3 ? 4 : 2)] = __ldg(q1 + 2);
    } else if (C == 4) {
        const uint32_t* __restrict__ q0 = base + (o0 >> 2);  // RGBA planes are 4-byte aligned
        const uint32_t* __restrict__ q1 = base + (o1 >> 2);
        f.a[0] = __ldg(q0);
        f.a[1] = __ldg(q0 + 1);
        f.b[0] = __ldg(q1);
        f.b[1] = __ldg(q1 + 1);
    } else {
        const uint8_t* __restrict__ p0 = reinterpret_cast<const uint8_t*>(base) + o0;
        const uint8_t* __restrict__ p1 = reinterpret_cast<const uint8_t*>(base) + o1;
        f.a[0] = __ldg(p0);
        f.a[1] = __ldg(p0 + 1);
        f.b[0] = __ldg(p1);
        f.b[1] = __ldg(p1 + 1);
    }
}

// RGB taps addressed in BITS: `pitch8` = 8 * pitch, `mis8` = 8 * (base & 3); the word index is
// the bit offset >> 5 and the funnel shift takes its low five bits by itself, so a row costs
// no separate shift amount.  For planes below 2^28 bytes (the caller checks).
__device__ __forceinline__ void fetch2_request_rgb_bits(const uint32_t* __restrict__ base, int mis8,
                                                        int pitch8, const Tap2& t, Fetch2<3>& f) {
    const int o0 = t.ys * pitch8 + (t.xs * 24 + mis8);
    const int o1 = o0 + pitch8;
    const uint32_t* __restrict__ q0 = base + (o0 >> 5);
    const uint32_t* __restrict__ q1 = base + (o1 >> 5);
    f.a[3] = (uint32_t)o0;
    f.b[3] = (uint32_t)o1;
    f.a[0] = __ldg(q0);
    f.a[1] = __ldg(q0 + 1);
    f.b[0] = __ldg(q1);
    f.b[1] = __ldg(q1 + 1);
    f.a[2] = 0u;
    f.b[2] = 0u;
    // only a byte offset of 3 pushes the 6 bytes into a third word
    if ((~(uint32_t)o0 & 24u) == 0u) f.a[2] = __ldg(q0 + 2);
    if ((~(uint32_t)o1 & 24u) == 0u) f.b[2] = __ldg(q1 + 2);
}

// RGB row: align the 6 bytes (funnel shifts by w[3] mod 32 bits), pair the channels (PRMT):
//   rg = R0 R1 G0 G1,  bb = B0 B1 . .
__device__ __forceinline__ void rgb_row_pairs(const uint32_t* w, uint32_t& rg, uint32_t& bb) {
    const uint32_t sh = w[3];
    const uint32_t lo = __funnelshift_r(w[0], w[1], sh);  // R0 G0 B0 R1
    const uint32_t hi = __funnelshift_r(w[1], w[2], sh);  // G1 B1 .  .
    rg = __byte_perm(lo, hi, 0x4130);
    bb = __byte_perm(lo, hi, 0x5252);
}

// blended channels (each already shifted down: 0 .. 255)
template <int C>
__device__ __forceinline__ void fetch2_blend(const Fetch2<C>& f, const Tap2& t, uint32_t* out) {
    if (C == 3) {
        uint32_t rg0, bb0, rg1, bb1;
        rgb_row_pairs(f.a, rg0, bb0);
        rgb_row_pairs(f.b, rg1, bb1);
        // sums are < 2^18, so (v >> 10) is the byte
        out[0] = __dp2a_lo(t.w_row0, rg0, __dp2a_lo(t.w_row1, rg1, 512u)) >> 10;
        out[1 % C] = __dp2a_hi(t.w_row0, rg0, __dp2a_hi(t.w_row1, rg1, 512u)) >> 10;
        out[2 % C] = __dp2a_lo(t.w_row0, bb0, __dp2a_lo(t.w_row1, bb1, 512u)) >> 10;
    } else if (C == 4) {
        const uint32_t rg0 = __byte_perm(f.a[0], f.a[1], 0x5140), ba0 = __byte_perm(f.a[0], f.a[1], 0x7362);
        const uint32_t rg1 = __byte_perm(f.b[0], f.b[1], 0x5140), ba1 = __byte_perm(f.b[0], f.b[1], 0x7362);
        out[0] = __dp2a_lo(t.w_row0, rg0, __dp2a_lo(t.w_row1, rg1, 512u)) >> 10;
        out[1 % C] = __dp2a_hi(t.w_row0, rg0, __dp2a_hi(t.w_row1, rg1, 512u)) >> 10;
        out[2 % C] = __dp2a_lo(t.w_row0, ba0, __dp2a_lo(t.w_row1, ba1, 512u)) >> 10;
        out[3 % C] = __dp2a_hi(t.w_row0, ba0, __dp2a_hi(t.w_row1, ba1, 512u)) >> 10;
    } else {
        const uint32_t p0 = f.a[0] | (f.a[1] << 8), p1 = f.b[0] | (f.b[1] << 8);
        out[0] = __dp2a_lo(t.w_row0, p0, __dp2a_lo(t.w_row1, p1, 512u)) >> 10;
    }
}

// planes narrower or shorter than 2 px: plain per-tap bounds checks
template <int C>
__device__ __noinline__ uint32_t sample_u8_small(const uint8_t* __restrict__ src, int h, int w,
                                                 int X, int Y) {
    uint8_t out[4] = {0, 0, 0, 0};
    bilinear_u8<C>(src, h, w, (long long)w * C, X, Y, out);
    return out[0] | (out[1] << 8) | (out[2] << 16) | ((uint32_t)out[3] << 24);
}

}  // namespace vkb
