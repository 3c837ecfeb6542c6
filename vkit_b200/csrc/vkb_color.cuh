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

// RCPPS(d) for d = 1..255: the 12-bit hardware reciprocal approximation of x86 (Intel's table, the
// values `_mm_rcp_ss` returns; generator: oracle/ipp_rcp_table.c).  The IPP routine behind
// cv.cvtColor(RGB2HLS_FULL) in the cv2 wheel divides with it, so its results are these values'.
#define VKB_IPP_RCP_VALUES \
    0.999755859f, 0.49987793f, 0.333251953f, 0.249938965f, 0.199951172f, 0.166625977f, \
    0.142822266f, 0.124969482f, 0.111083984f, 0.0999755859f, 0.0908966064f, 0.0833129883f, \
    0.0769042969f, 0.0714111328f, 0.0666503906f, 0.0624847412f, 0.058807373f, 0.0555419922f, \
    0.0526199341f, 0.049987793f, 0.0476074219f, 0.0454483032f, 0.04347229f, 0.0416564941f, \
    0.0399932861f, 0.0384521484f, 0.0370330811f, 0.0357055664f, 0.0344772339f, 0.0333251953f, \
    0.0322570801f, 0.0312423706f, 0.0302963257f, 0.0294036865f, 0.0285644531f, 0.0277709961f, \
    0.0270195007f, 0.026309967f, 0.0256347656f, 0.0249938965f, 0.0243873596f, 0.0238037109f, \
    0.0232505798f, 0.0227241516f, 0.0222167969f, 0.021736145f, 0.0212745667f, 0.0208282471f, \
    0.0204048157f, 0.0199966431f, 0.0196037292f, 0.0192260742f, 0.018863678f, 0.0185165405f, \
    0.0181808472f, 0.0178527832f, 0.017539978f, 0.0172386169f, 0.0169487f, 0.0166625977f, \
    0.0163917542f, 0.01612854f, 0.0158691406f, 0.0156211853f, 0.0153808594f, 0.0151481628f, \
    0.0149211884f, 0.0147018433f, 0.0144901276f, 0.0142822266f, 0.014081955f, 0.013885498f, \
    0.0136947632f, 0.0135097504f, 0.0133304596f, 0.0131549835f, 0.0129852295f, 0.0128173828f, \
    0.0126552582f, 0.0124969482f, 0.012342453f, 0.0121936798f, 0.012046814f, 0.0119018555f, \
    0.011762619f, 0.0116252899f, 0.0114917755f, 0.0113620758f, 0.0112342834f, 0.0111083984f, \
    0.0109863281f, 0.0108680725f, 0.0107517242f, 0.0106372833f, 0.0105247498f, 0.0104141235f, \
    0.010307312f, 0.0102024078f, 0.010099411f, 0.00999832153f, 0.0098991394f, 0.00980186462f, \
    0.00970649719f, 0.00961303711f, 0.00952148438f, 0.00943183899f, 0.00934410095f, 0.00925827026f, \
    0.00917243958f, 0.00909042358f, 0.00900840759f, 0.0089263916f, 0.00884819031f, 0.00876998901f, \
    0.00869369507f, 0.00861930847f, 0.00854492188f, 0.00847434998f, 0.00840187073f, 0.00833129883f, \
    0.00826263428f, 0.00819587708f, 0.00812911987f, 0.00806427002f, 0.00799942017f, 0.00793457031f, \
    0.00787353516f, 0.00781059265f, 0.00775051117f, 0.00769042969f, 0.00763130188f, 0.00757408142f, \
    0.00751686096f, 0.00746059418f, 0.00740528107f, 0.00735092163f, 0.00729751587f, 0.00724506378f, \
    0.00719261169f, 0.00714111328f, 0.00709056854f, 0.00704097748f, 0.00699138641f, 0.00694274902f, \
    0.00689506531f, 0.00684738159f, 0.00680160522f, 0.00675487518f, 0.00671005249f, 0.0066652298f, \
    0.00662136078f, 0.00657749176f, 0.00653457642f, 0.00649261475f, 0.00645065308f, 0.00640869141f, \
    0.00636768341f, 0.00632762909f, 0.00628852844f, 0.00624847412f, 0.00621032715f, 0.0061712265f, \
    0.0061340332f, 0.0060968399f, 0.00605964661f, 0.00602340698f, 0.00598716736f, 0.00595092773f, \
    0.00591564178f, 0.00588130951f, 0.00584697723f, 0.00581264496f, 0.00577926636f, 0.00574588776f, \
    0.00571346283f, 0.0056810379f, 0.00564861298f, 0.00561714172f, 0.00558567047f, 0.00555419922f, \
    0.00552368164f, 0.00549316406f, 0.00546360016f, 0.00543403625f, 0.00540447235f, 0.00537586212f, \
    0.00534629822f, 0.00531864166f, 0.00529003143f, 0.00526237488f, 0.00523471832f, 0.00520706177f, \
    0.00518035889f, 0.00515365601f, 0.00512695312f, 0.00510120392f, 0.00507545471f, 0.00504970551f, \
    0.0050239563f, 0.00499916077f, 0.00497436523f, 0.0049495697f, 0.00492572784f, 0.00490093231f, \
    0.00487709045f, 0.0048532486f, 0.00483036041f, 0.00480651855f, 0.00478363037f, 0.00476074219f, \
    0.00473880768f, 0.00471591949f, 0.00469398499f, 0.00467205048f, 0.00465011597f, 0.00462913513f, \
    0.00460720062f, 0.00458621979f, 0.00456523895f, 0.00454521179f, 0.00452423096f, 0.0045042038f, \
    0.00448322296f, 0.0044631958f, 0.00444412231f, 0.00442409515f, 0.00440502167f, 0.00438499451f, \
    0.00436592102f, 0.00434684753f, 0.00432872772f, 0.00430965424f, 0.00429153442f, 0.00427246094f, \
    0.00425434113f, 0.00423717499f, 0.00421905518f, 0.00420093536f, 0.00418376923f, 0.00416564941f, \
    0.00414848328f, 0.00413131714f, 0.00411510468f, 0.00409793854f, 0.0040807724f, 0.00406455994f, \
    0.00404834747f, 0.00403213501f, 0.00401592255f, 0.00399971008f, 0.00398349762f, 0.00396728516f, \
    0.00395202637f, 0.00393676758f, 0.00392150879f, \

static const float kIppRcpHost[256] = {0.f, VKB_IPP_RCP_VALUES};
#if defined(__CUDACC__)
static __device__ const float kIppRcpDev[256] = {0.f, VKB_IPP_RCP_VALUES};
#endif

VKB_HD float ipp_rcp(int d) {
#if defined(__CUDA_ARCH__)
    return __ldg(&kIppRcpDev[d]);
#else
    return kIppRcpHost[d];
#endif
}

// COLOR_RGB2HLS_FULL, uint8, as the wheel's default backend (Intel IPP) computes it -- pinned on all
// 2^24 colours (tests/test_oracle_cv2_model.py):
//   L = round_half_even((max + min) / 2)
//   S = rint(diff * RCPPS(den) * 255),  den = max + min when <= 255, else 510 - (max + min)
//   H = rint(h * 42.5),  h = (g - b) * RCPPS(diff) (+ 2, + 4 for a green / blue maximum, tested in
//       the order R, G, B), + 6 when negative; 256 wraps to 0.  All products in float32.
VKB_HD void rgb2hls_full(int R, int G, int B, int& h, int& l, int& s) {
    const int imax = R > G ? (R > B ? R : B) : (G > B ? G : B);
    const int imin = R < G ? (R < B ? R : B) : (G < B ? G : B);
    const int isum = imax + imin, half = isum >> 1, diff = imax - imin;
    l = (isum & 1) ? half + (half & 1) : half;
    h = 0;
    s = 0;
    if (diff == 0) return;
    const int den = isum <= 255 ? isum : 510 - isum;
    s = round_u8(VKB_FMUL(VKB_FMUL((float)diff, ipp_rcp(den)), 255.f));
    const float rd = ipp_rcp(diff);
    float hf;
    if (imax == R) hf = VKB_FMUL((float)(G - B), rd);
    else if (imax == G) hf = VKB_FADD(VKB_FMUL((float)(B - R), rd), 2.f);
    else hf = VKB_FADD(VKB_FMUL((float)(R - G), rd), 4.f);
    if (hf < 0.f) hf = VKB_FADD(hf, 6.f);
#if defined(__CUDA_ARCH__)
    const int hi = __float2int_rn(VKB_FMUL(hf, 42.5f));
#else
    const int hi = (int)rintf(VKB_FMUL(hf, 42.5f));
#endif
    h = hi >= 256 ? hi - 256 : hi;
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
