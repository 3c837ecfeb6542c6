// jpeg.cu -- the pixel arithmetic of a baseline JPEG round trip on the device: what
// cv.imencode('.jpeg', mat, [IMWRITE_JPEG_QUALITY, q]) + cv.imdecode do to the pixels of a page
// through libjpeg-turbo (vkit jpeg_quality, photometric/effect.py:26-55).  Entropy coding is
// lossless and skipped; every step that changes a sample follows libjpeg's integer arithmetic:
//
//   jpeg_forward_kernel   per 16 x 16 MCU: jccolor.c RGB -> YCbCr (16-bit fixed point), edge
//                         replication (jcsample.c expand_right_edge, jcprepct.c expand_bottom_edge),
//                         h2v2 box downsampling with the 1,2,1,2 bias, then for the six 8 x 8
//                         blocks jfdctint.c (islow) -> quantise (jcdctmgr.c) -> dequantise ->
//                         jidctint.c (islow) + range limit; decoded Y / Cb / Cr planes out
//   jpeg_inverse_kernel   per pixel: jdsample.c h2v2_fancy_upsample (triangle filter; box
//                         replication when the chroma plane is at most 2 samples wide) and
//                         jdcolor.c YCbCr -> RGB
//
// Integer work, HBM bound: 3 B/px in, 1.5 B/px of planes written and read back, 3 B/px out.
// Bit exact against cv2 4.13 (oracle/jpeg_model.py is pinned against the wheel; the GPU tests
// compare with fixtures of the live reference).
#include "common.cuh"

namespace vkb {

struct JpegTables {
    uint16_t luma[64];
    uint16_t chroma[64];
};

// jfdctint.c / jidctint.c: CONST_BITS = 13, PASS1_BITS = 2
constexpr int kF0298 = 2446, kF0390 = 3196, kF0541 = 4433, kF0765 = 6270, kF0899 = 7373,
              kF1175 = 9633, kF1501 = 12299, kF1847 = 15137, kF1961 = 16069, kF2053 = 16819,
              kF2562 = 20995, kF3072 = 25172;

__device__ __forceinline__ int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

// one 8-point forward pass (rows: FIRST = true, columns: false), in place
template <bool FIRST>
__device__ __forceinline__ void fdct8(int* d) {
    const int t0 = d[0] + d[7], t7 = d[0] - d[7];
    const int t1 = d[1] + d[6], t6 = d[1] - d[6];
    const int t2 = d[2] + d[5], t5 = d[2] - d[5];
    const int t3 = d[3] + d[4], t4 = d[3] - d[4];
    const int t10 = t0 + t3, t13 = t0 - t3, t11 = t1 + t2, t12 = t1 - t2;
    constexpr int n = FIRST ? 13 - 2 : 13 + 2;
    d[0] = FIRST ? (t10 + t11) << 2 : descale(t10 + t11, 2);
    d[4] = FIRST ? (t10 - t11) << 2 : descale(t10 - t11, 2);
    int z1 = (t12 + t13) * kF0541;
    d[2] = descale(z1 + t13 * kF0765, n);
    d[6] = descale(z1 - t12 * kF1847, n);
    z1 = t4 + t7;
    int z2 = t5 + t6, z3 = t4 + t6, z4 = t5 + t7;
    const int z5 = (z3 + z4) * kF1175;
    const int a4 = t4 * kF0298, a5 = t5 * kF2053, a6 = t6 * kF3072, a7 = t7 * kF1501;
    z1 = -z1 * kF0899;
    z2 = -z2 * kF2562;
    z3 = -z3 * kF1961 + z5;
    z4 = -z4 * kF0390 + z5;
    d[7] = descale(a4 + z1 + z3, n);
    d[5] = descale(a5 + z2 + z4, n);
    d[3] = descale(a6 + z2 + z3, n);
    d[1] = descale(a7 + z1 + z4, n);
}

// one 8-point inverse pass (columns: FIRST = true, rows: false), in place
template <bool FIRST>
__device__ __forceinline__ void idct8(int* d) {
    int z2 = d[2], z3 = d[6];
    int z1 = (z2 + z3) * kF0541;
    const int t2 = z1 - z3 * kF1847, t3 = z1 + z2 * kF0765;
    z2 = d[0];
    z3 = d[4];
    const int t0 = (z2 + z3) << 13, t1 = (z2 - z3) << 13;
    const int t10 = t0 + t3, t13 = t0 - t3, t11 = t1 + t2, t12 = t1 - t2;
    int a0 = d[7], a1 = d[5], a2 = d[3], a3 = d[1];
    z1 = a0 + a3;
    z2 = a1 + a2;
    z3 = a0 + a2;
    int z4 = a1 + a3;
    const int z5 = (z3 + z4) * kF1175;
    a0 *= kF0298;
    a1 *= kF2053;
    a2 *= kF3072;
    a3 *= kF1501;
    z1 = -z1 * kF0899;
    z2 = -z2 * kF2562;
    z3 = -z3 * kF1961 + z5;
    z4 = -z4 * kF0390 + z5;
    a0 += z1 + z3;
    a1 += z2 + z4;
    a2 += z2 + z3;
    a3 += z1 + z4;
    constexpr int n = FIRST ? 13 - 2 : 13 + 2 + 3;
    d[0] = descale(t10 + a3, n);
    d[7] = descale(t10 - a3, n);
    d[1] = descale(t11 + a2, n);
    d[6] = descale(t11 - a2, n);
    d[2] = descale(t12 + a1, n);
    d[5] = descale(t12 - a1, n);
    d[3] = descale(t13 + a0, n);
    d[4] = descale(t13 - a0, n);
}

constexpr int kFix = 65536;
__device__ __forceinline__ int fix(double x) { return (int)(x * kFix + 0.5); }

// Block = one 16 x 16 MCU, 256 threads.  C == 3: channel 0 plays BLUE (cv2 hands libjpeg B, G, R
// and vkit passes RGB arrays); C == 1: luma only.
template <int C>
__global__ void __launch_bounds__(256) jpeg_forward_kernel(const uint8_t* __restrict__ src, int h,
                                                           int w, JpegTables tables,
                                                           uint8_t* __restrict__ plane_y,
                                                           uint8_t* __restrict__ plane_cb,
                                                           uint8_t* __restrict__ plane_cr, int wp) {
    __shared__ int blk[6][64];       // 4 luma blocks (raster order inside the MCU), Cb, Cr
    __shared__ int cfull[2][16][16];  // full-resolution Cb / Cr of the MCU
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int x0 = blockIdx.x * 16, y0 = blockIdx.y * 16;
    const int xc = min(x0 + tx, w - 1);
    {
        const int yc = min(y0 + ty, h - 1);  // luma: the last row / column repeats
        const uint8_t* p = src + ((long long)yc * w + xc) * C;
        int yv;
        if (C == 3) {
            const int b = p[0], g = p[1 % C], r = p[2 % C];
            yv = (19595 * r + 38470 * g + 7471 * b + 32768) >> 16;
            // chroma: the colour rows repeat to a whole row GROUP only; below that the
            // DOWNSAMPLED rows repeat (jcprepct.c), so a padded row pair reads the last real pair
            const int ch_real = (h + 1) >> 1;
            const int cyg = min((y0 + ty) >> 1, ch_real - 1);
            const int yrow = min(2 * cyg + (ty & 1), h - 1);
            int b2 = b, g2 = g, r2 = r;
            if (yrow != yc) {
                const uint8_t* q = src + ((long long)yrow * w + xc) * C;
                b2 = q[0];
                g2 = q[1 % C];
                r2 = q[2 % C];
            }
            cfull[0][ty][tx] = (-11059 * r2 - 21709 * g2 + 32768 * b2 + (128 << 16) + 32767) >> 16;
            cfull[1][ty][tx] = (32768 * r2 - 27439 * g2 - 5329 * b2 + (128 << 16) + 32767) >> 16;
        } else {
            yv = p[0];
        }
        blk[(ty >> 3) * 2 + (tx >> 3)][(ty & 7) * 8 + (tx & 7)] = yv - 128;
    }
    __syncthreads();
    if (C == 3 && threadIdx.x < 128) {
        const int comp = threadIdx.x >> 6, i = threadIdx.x & 63, cy = i >> 3, cx = i & 7;
        const int s = cfull[comp][2 * cy][2 * cx] + cfull[comp][2 * cy][2 * cx + 1]
                      + cfull[comp][2 * cy + 1][2 * cx] + cfull[comp][2 * cy + 1][2 * cx + 1];
        blk[4 + comp][i] = ((s + 1 + (cx & 1)) >> 2) - 128;  // bias 1, 2, 1, 2 (MCU origin is even)
    }
    __syncthreads();
    constexpr int kBlocks = C == 3 ? 6 : 4;
    const int task = threadIdx.x;  // (block, line)
    const int b = task >> 3, line = task & 7;
    int d[8];
    if (task < kBlocks * 8) {  // forward pass 1: rows
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] = blk[b][line * 8 + i];
        fdct8<true>(d);
#pragma unroll
        for (int i = 0; i < 8; ++i) blk[b][line * 8 + i] = d[i];
    }
    __syncthreads();
    if (task < kBlocks * 8) {
        // forward pass 2: columns; quantise (divisor q << 3, half away from zero), dequantise;
        // inverse pass 1: columns -- the column stays in registers
        const uint16_t* q = b < 4 ? tables.luma : tables.chroma;
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] = blk[b][i * 8 + line];
        fdct8<false>(d);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int qv = q[i * 8 + line], qval = qv << 3;
            const int mag = (abs(d[i]) + (qval >> 1)) / qval;
            d[i] = (d[i] < 0 ? -mag : mag) * qv;
        }
        idct8<true>(d);
#pragma unroll
        for (int i = 0; i < 8; ++i) blk[b][i * 8 + line] = d[i];
    }
    __syncthreads();
    if (task < kBlocks * 8) {  // inverse pass 2: rows, range limit, store
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] = blk[b][line * 8 + i];
        idct8<false>(d);
        uint32_t lo = 0, hi = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            lo |= (uint32_t)min(max(d[i] + 128, 0), 255) << (8 * i);
            hi |= (uint32_t)min(max(d[4 + i] + 128, 0), 255) << (8 * i);
        }
        uint8_t* out;
        if (b < 4) {
            out = plane_y + (long long)(y0 + (b >> 1) * 8 + line) * wp + x0 + (b & 1) * 8;
        } else {
            uint8_t* plane = b == 4 ? plane_cb : plane_cr;
            out = plane + (long long)((y0 >> 1) + line) * (wp >> 1) + (x0 >> 1);
        }
        *reinterpret_cast<uint2*>(out) = make_uint2(lo, hi);  // planes are 8-byte aligned per block
    }
}

template <int C>
__global__ void __launch_bounds__(256) jpeg_inverse_kernel(const uint8_t* __restrict__ plane_y,
                                                           const uint8_t* __restrict__ plane_cb,
                                                           const uint8_t* __restrict__ plane_cr,
                                                           int wp, uint8_t* __restrict__ dst, int h,
                                                           int w) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= w || y >= h) return;
    const int yv = plane_y[(long long)y * wp + x];
    if (C == 1) {
        dst[(long long)y * w + x] = (uint8_t)yv;
        return;
    }
    const int cw = (w + 1) >> 1, ch = (h + 1) >> 1, cwp = wp >> 1;
    const int cx = x >> 1, cy = y >> 1;
    int cb, cr;
    if (cw <= 2) {  // jinit_upsampler: narrow components are box replicated
        cb = plane_cb[(long long)cy * cwp + cx];
        cr = plane_cr[(long long)cy * cwp + cx];
    } else {
        // h2v2_fancy_upsample: 3/4 nearer + 1/4 further in each direction; the missing neighbour
        // of an edge sample is the sample itself (clamped index)
        const int oy = min(max((y & 1) ? cy + 1 : cy - 1, 0), ch - 1);
        const int ox = min(max((x & 1) ? cx + 1 : cx - 1, 0), cw - 1);
        const int bias = (x & 1) ? 7 : 8;
        const long long r0 = (long long)cy * cwp, r1 = (long long)oy * cwp;
        const int cb_this = plane_cb[r0 + cx] * 3 + plane_cb[r1 + cx];
        const int cb_other = plane_cb[r0 + ox] * 3 + plane_cb[r1 + ox];
        const int cr_this = plane_cr[r0 + cx] * 3 + plane_cr[r1 + cx];
        const int cr_other = plane_cr[r0 + ox] * 3 + plane_cr[r1 + ox];
        cb = (cb_this * 3 + cb_other + bias) >> 4;
        cr = (cr_this * 3 + cr_other + bias) >> 4;
    }
    const int xcb = cb - 128, xcr = cr - 128;
    const int r = yv + ((91881 * xcr + 32768) >> 16);
    const int g = yv + ((-22554 * xcb + 32768 - 46802 * xcr) >> 16);
    const int b = yv + ((116130 * xcb + 32768) >> 16);
    uint8_t* p = dst + ((long long)y * w + x) * 3;
    p[0] = (uint8_t)min(max(b, 0), 255);  // channel 0 played blue on the way in
    p[1] = (uint8_t)min(max(g, 0), 255);
    p[2] = (uint8_t)min(max(r, 0), 255);
}

}  // namespace vkb

using namespace vkb;

static void fill_quant(const int* base, int quality, uint16_t* out) {
    // jpeg_set_quality(quality, force_baseline = TRUE)
    quality = quality < 1 ? 1 : (quality > 100 ? 100 : quality);
    const int scale = quality < 50 ? 5000 / quality : 200 - quality * 2;
    for (int i = 0; i < 64; ++i) {
        int v = (base[i] * scale + 50) / 100;
        out[i] = (uint16_t)(v < 1 ? 1 : (v > 255 ? 255 : v));
    }
}

extern "C" int vkb_jpeg_round_trip_u8(const uint8_t* src, uint8_t* dst, int32_t h, int32_t w,
                                      int32_t channels, int32_t quality, uint8_t* planes,
                                      int64_t planes_bytes, void* stream) {
    VKB_NVTX("vkb_jpeg_round_trip_u8");
    static const int kLuma[64] = {16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55,
                                  14, 13, 16, 24, 40, 57, 69, 56, 14, 17, 22, 29, 51, 87, 80, 62,
                                  18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92,
                                  49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
    static const int kChroma[64] = {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99,
                                    24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99,
                                    99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
                                    99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};
    VKB_REQUIRE(src && dst && planes && h > 0 && w > 0, "bad arguments");
    VKB_REQUIRE(channels == 1 || channels == 3, "channels must be 1 (GRAYSCALE) or 3");
    VKB_REQUIRE(h <= 65500 && w <= 65500, "JPEG dimensions are limited to 65500");
    const int hp = (h + 15) / 16 * 16, wp = (w + 15) / 16 * 16;
    const int64_t need = (int64_t)hp * wp + 2 * (int64_t)(hp / 2) * (wp / 2);
    VKB_REQUIRE(planes_bytes >= need, "planes workspace too small (1.5 bytes per padded pixel)");
    VKB_REQUIRE((reinterpret_cast<uintptr_t>(planes) & 15) == 0, "planes must be 16-byte aligned");
    JpegTables tables;
    fill_quant(kLuma, quality, tables.luma);
    fill_quant(kChroma, quality, tables.chroma);
    uint8_t* plane_y = planes;
    uint8_t* plane_cb = planes + (int64_t)hp * wp;
    uint8_t* plane_cr = plane_cb + (int64_t)(hp / 2) * (wp / 2);
    cudaStream_t st = (cudaStream_t)stream;
    const dim3 mcus(wp / 16, hp / 16), tiles((w + 31) / 32, (h + 7) / 8);
    if (channels == 3) {
        jpeg_forward_kernel<3><<<mcus, 256, 0, st>>>(src, h, w, tables, plane_y, plane_cb, plane_cr, wp);
        jpeg_inverse_kernel<3><<<tiles, 256, 0, st>>>(plane_y, plane_cb, plane_cr, wp, dst, h, w);
    } else {
        jpeg_forward_kernel<1><<<mcus, 256, 0, st>>>(src, h, w, tables, plane_y, plane_cb, plane_cr, wp);
        jpeg_inverse_kernel<1><<<tiles, 256, 0, st>>>(plane_y, plane_cb, plane_cr, wp, dst, h, w);
    }
    return check_launch("jpeg kernels");
}
