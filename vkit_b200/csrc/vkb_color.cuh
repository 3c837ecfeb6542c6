// vkb_color.cuh -- per-pixel colour arithmetic (cv.cvtColor codes used by the reference,
// vkit/element/image.py:183-207, and the photometric ops of photometric/color.py).
// `__host__ __device__` so tests/hostsim can run the same code on the CPU.
#pragma once
#include "vkb_math.cuh"

namespace vkb {

// Division tables of cv::RGB2HSV_b: sdiv[i] = round((255 << 12) / i), hdiv[i] = round((256 << 12) / (6 i)).
// Filled once on the host (fill_hsv_tables) and copied to constant memory by photometric.cu.
struct HsvTables {
    int sdiv[256];
    int hdiv[256];
};

inline void fill_hsv_tables(HsvTables& t) {
    t.sdiv[0] = t.hdiv[0] = 0;
    for (int i = 1; i < 256; ++i) {
        t.sdiv[i] = (int)rint((255 << 12) / (1.0 * i));
        t.hdiv[i] = (int)rint((256 << 12) / (6.0 * i));
    }
}

// COLOR_RGB2HSV_FULL, uint8: integer arithmetic, hue range 256 (bit exact vs cv2 on 2^24 colours).
VKB_HD void rgb2hsv_full(const HsvTables& t, int r, int g, int b, int& h, int& s, int& v) {
    v = r > g ? (r > b ? r : b) : (g > b ? g : b);
    const int vmin = r < g ? (r < b ? r : b) : (g < b ? g : b);
    const int diff = v - vmin;
    const int vr = (v == r) ? -1 : 0;
    const int vg = (v == g) ? -1 : 0;
    s = (diff * t.sdiv[v] + (1 << 11)) >> 12;
    int hh = (vr & (g - b)) + (~vr & ((vg & (b - r + 2 * diff)) + ((~vg) & (r - g + 4 * diff))));
    hh = (hh * t.hdiv[diff] + (1 << 11)) >> 12;
    hh += hh < 0 ? 256 : 0;
    h = hh > 255 ? 255 : hh;
}

VKB_HD int round_u8(float x) {
#if defined(__CUDA_ARCH__)
    return min(max(__float2int_rn(x), 0), 255);  // F2I rounds half to even and saturates
#else
    const float r = rintf(x);
    return r < 0.f ? 0 : (r > 255.f ? 255 : (int)r);
#endif
}

// sector tables of cv::HSV2RGB_native / HLS2RGB_native: which of tab[0..3] goes to b, g, r.
VKB_HD float tab_pick(const float* tab, int idx) {
    const float lo = (idx & 1) ? tab[1] : tab[0];
    const float hi = (idx & 1) ? tab[3] : tab[2];
    return (idx & 2) ? hi : lo;
}

VKB_HD void sector_pick(const float* tab, int sector, float& r, float& g, float& b) {
    // {b, g, r} = {{1,3,0},{1,0,2},{3,0,1},{0,2,1},{0,1,3},{2,1,0}}[sector], two bits per entry;
    // selects instead of a switch: the sectors of neighbouring pixels differ (no divergence).
    const int sh = 2 * sector;
    r = tab_pick(tab, (0x358 >> sh) & 3);
    g = tab_pick(tab, (0x583 >> sh) & 3);
    b = tab_pick(tab, (0x835 >> sh) & 3);
}

// COLOR_HSV2RGB_FULL, uint8 through the float32 path, hue scale 6/255.
VKB_HD void hsv2rgb_full(int H, int S, int V, int& r, int& g, int& b) {
    const float h = VKB_FMUL((float)H, (float)(6.0 / 255.0));
    const float s = VKB_FMUL((float)S, (float)(1.0 / 255.0));
    const float v = VKB_FMUL((float)V, (float)(1.0 / 255.0));
    int sector = (int)floorf(h);  // 0..6 for H in 0..255
    const float frac = VKB_FSUB(h, (float)sector);
    sector = sector >= 6 ? sector - 6 : sector;
    float tab[4];
    tab[0] = v;
    tab[1] = VKB_FMUL(v, VKB_FSUB(1.f, s));
    tab[2] = VKB_FMUL(v, VKB_FSUB(1.f, VKB_FMUL(s, frac)));
    tab[3] = VKB_FMUL(v, VKB_FSUB(1.f, VKB_FMUL(s, VKB_FSUB(1.f, frac))));
    float fr, fg, fb;
    sector_pick(tab, sector, fr, fg, fb);
    r = round_u8(VKB_FMUL(fr, 255.f));
    g = round_u8(VKB_FMUL(fg, 255.f));
    b = round_u8(VKB_FMUL(fb, 255.f));
}

// COLOR_RGB2HLS_FULL, uint8, hue scale 255/360 (the wheel's default IPP backend).
// L = round_half_even((max + min) / 2): exact.  H, S through float32: within +-1 of cv2, whose
// own result depends on the IPP / SIMD / scalar backend (appendix A.6).
VKB_HD void rgb2hls_full(int R, int G, int B, int& h, int& l, int& s) {
    const float k = (float)(1.0 / 255.0);
    const float r = VKB_FMUL((float)R, k), g = VKB_FMUL((float)G, k), b = VKB_FMUL((float)B, k);
    const float vmax = fmaxf(fmaxf(r, g), b);
    const float vmin = fminf(fminf(r, g), b);
    const float diff = VKB_FSUB(vmax, vmin);
    const float sum = VKB_FADD(vmax, vmin);
    const float lf = VKB_FMUL(sum, 0.5f);
    float hf = 0.f, sf = 0.f;
    if (diff > 1.1920929e-07f) {
        sf = lf < 0.5f ? diff / sum : diff / VKB_FSUB(VKB_FSUB(2.f, vmax), vmin);
        const float d = 60.f / diff;
        if (vmax == r) hf = VKB_FMUL(VKB_FSUB(g, b), d);
        else if (vmax == g) hf = VKB_FADD(VKB_FMUL(VKB_FSUB(b, r), d), 120.f);
        else hf = VKB_FADD(VKB_FMUL(VKB_FSUB(r, g), d), 240.f);
        if (hf < 0.f) hf = VKB_FADD(hf, 360.f);
    }
    h = round_u8(VKB_FMUL(hf, (float)(255.0 / 360.0)));
    const int imax = R > G ? (R > B ? R : B) : (G > B ? G : B);
    const int imin = R < G ? (R < B ? R : B) : (G < B ? G : B);
    const int isum = imax + imin, half = isum >> 1;
    l = (isum & 1) ? half + (half & 1) : half;
    s = round_u8(VKB_FMUL(sf, 255.f));
}

// COLOR_HLS2RGB_FULL, uint8 through float32, hue scale 6/255.
VKB_HD void hls2rgb_full(int H, int L, int S, int& r, int& g, int& b) {
    const float h = VKB_FMUL((float)H, (float)(6.0 / 255.0));
    const float l = VKB_FMUL((float)L, (float)(1.0 / 255.0));
    const float s = VKB_FMUL((float)S, (float)(1.0 / 255.0));
    float fr = l, fg = l, fb = l;
    if (S != 0) {
        const float p2 = l <= 0.5f ? VKB_FMUL(l, VKB_FADD(1.f, s))
                                   : VKB_FSUB(VKB_FADD(l, s), VKB_FMUL(l, s));
        const float p1 = VKB_FSUB(VKB_FMUL(2.f, l), p2);
        int sector = (int)floorf(h);  // 0..6 for H in 0..255
        const float frac = VKB_FSUB(h, (float)sector);
        sector = sector >= 6 ? sector - 6 : sector;
        float tab[4];
        tab[0] = p2;
        tab[1] = p1;
        tab[2] = VKB_FADD(p1, VKB_FMUL(VKB_FSUB(p2, p1), VKB_FSUB(1.f, frac)));
        tab[3] = VKB_FADD(p1, VKB_FMUL(VKB_FSUB(p2, p1), frac));
        sector_pick(tab, sector, fr, fg, fb);
    }
    r = round_u8(VKB_FMUL(fr, 255.f));
    g = round_u8(VKB_FMUL(fg, 255.f));
    b = round_u8(VKB_FMUL(fb, 255.f));
}

// COLOR_RGB2GRAY, uint8: 15-bit coefficients of cv2 4.13 (pinned on all 2^24 colours).
VKB_HD int rgb2gray(int r, int g, int b) {
    return (r * 9798 + g * 19235 + b * 3735 + (1 << 14)) >> 15;
}

// fill_np_array blend (element/opt.py:171-209): float32, separately rounded, truncated.
VKB_HD float blend_f32(float dst, float value, float a) {
    const float wm = VKB_FSUB(1.f, a);
    return VKB_FADD(VKB_FMUL(wm, dst), VKB_FMUL(a, value));
}

VKB_HD int floor_mod_256(int v) { return v & 255; }  // two's complement: floor modulo
VKB_HD int clip_u8(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

}  // namespace vkb
