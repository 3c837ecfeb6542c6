// draw.cu -- cv.ellipse on the device for ellipse_streak
// (vkit/mechanism/distortion/photometric/streak.py:282-337): the host turns every ellipse into its
// polyline segments (EllipseEx / ellipse2Poly: a few dozen scalar operations per ellipse), one
// thread per segment draws it into a uint8 mask with cv's own primitives (vkb_draw.cuh).  Every
// write stores 1, so segments and ellipses need no ordering.  Latency-bound, kilobytes of traffic.
#include <vector>
#include "common.cuh"
#include "vkb_draw.cuh"
#include "vkb_draw_host.h"

namespace vkb {

__global__ void __launch_bounds__(64) ellipse_segments_kernel(uint8_t* __restrict__ mask, int h, int w,
                                                              const DrawSegment* __restrict__ segs,
                                                              int n, int thickness) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const DrawSegment sg = segs[i];
    auto plot = [&](int x, int y) { mask[(size_t)y * w + x] = 1; };
    auto hline = [&](int y, int xa, int xb) {
        uint8_t* row = mask + (size_t)y * w;
        for (int x = xa; x <= xb; ++x) row[x] = 1;
    };
    draw_thick_segment(w, h, sg.p0, sg.p1, thickness, sg.flags, plot, hline);
}

}  // namespace vkb

using namespace vkb;

extern "C" int vkb_draw_ellipses(uint8_t* mask, int32_t h, int32_t w, const int32_t* ellipses_host,
                                 int32_t n_ellipses, int32_t thickness, void* seg_workspace,
                                 int64_t workspace_bytes, void* stream) {
    VKB_NVTX("vkb_draw_ellipses");
    VKB_REQUIRE(mask && ellipses_host && seg_workspace && h > 0 && w > 0, "bad arguments");
    VKB_REQUIRE(thickness >= 1 && thickness <= 255, "thickness must be 1..255 (outlines only)");
    VKB_REQUIRE(h < 16384 && w < 16384, "canvases below 16384 px");
    if (n_ellipses <= 0) return VKB_OK;
    std::vector<DrawSegment> segs;
    segs.reserve((size_t)n_ellipses * 74);
    for (int i = 0; i < n_ellipses; ++i) {
        const int32_t* e = ellipses_host + 4 * (size_t)i;
        VKB_REQUIRE(e[2] >= 0 && e[3] >= 0 && e[2] < 32768 && e[3] < 32768, "axes out of range");
        ellipse_segments(e[0], e[1], e[2], e[3], segs);
    }
    const size_t bytes = segs.size() * sizeof(DrawSegment);
    VKB_REQUIRE((int64_t)bytes <= workspace_bytes, "segment workspace too small (74 segments of 40 bytes per ellipse)");
    cudaStream_t st = (cudaStream_t)stream;
    // pageable source: the call returns once the segments are staged, the vector may go
    VKB_CUDA(cudaMemcpyAsync(seg_workspace, segs.data(), bytes, cudaMemcpyHostToDevice, st));
    const int n = (int)segs.size();
    ellipse_segments_kernel<<<(n + 63) / 64, 64, 0, st>>>(mask, h, w,
                                                          reinterpret_cast<const DrawSegment*>(seg_workspace),
                                                          n, thickness);
    return check_launch("ellipse_segments_kernel");
}
