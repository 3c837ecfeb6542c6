// remap.cu -- the fused remap of the grid-based ops (camera_*, similarity_mls):
// owner cell + inverse homography + cv::remap bilinear gather of Image + Mask + ScoreMap in one
// pass (type.py:209-261, grid_blender.py:54-81).  Two kernels share the flat tile work list that
// geometric.cu builds:
//
//   grid_remap_tiles_kernel   tiles with at most 15 candidate cells (99.8 % of them)
//   grid_remap_kernel<LARGE>  the remaining tiles (more candidates, or the exact slow path)
//
// HBM / issue bound integer work; no tensor cores (nothing here is a contraction).
#include <mutex>
#include <stdlib.h>
#include "common.cuh"
#include "vkb_math.cuh"
#include "vkb_lattice.cuh"
#include "vkb_grid.cuh"
#include "vkb_gather.cuh"

namespace vkb {

// ============================================================================================
// Fused remap.  Block = 32 x (32 / R) threads on one 32 x 32 dst tile; warp w owns rows
// R*w .. R*w + R-1, lane = column.
//
//   prologue  the records of the tile's candidate cells (<= VKB_TILE_CAP, ascending cell order,
//             read from the tile's sorted list) are staged in shared memory: bbox and the
//             float32 form of the cell's inverse homography re-centred on the cell's bbox
//             origin (CellLocal, one record per cell, written by grid_cells_kernel);
//   owner     lane-parallel: lane i fetches candidate i's coverage words for the warp's R rows
//             (shifted to the tile's columns), a ballot keeps the candidates that touch the
//             band, and each survivor is broadcast with shuffles in ascending order; a pixel
//             whose coverage bit is set takes the survivor's key, so the last (= largest) cell
//             wins, exactly like the reference's cell-by-cell map writes;
//   coords    float32 fast path with a proven acceptance test, float64 path for the pixels that
//             sit next to a rounding boundary (vkb_math.cuh);
//   gather    cv::remap's fixed-point bilinear for Image (C channels), Mask and ScoreMap.  The
//             L1 data pipe is the scarce resource here, so the RGB taps of a row (6 contiguous
//             bytes at any alignment) come in as 2-3 aligned 32-bit loads, are aligned with
//             PRMT and blended horizontally with IDP.4A (byte weights) instead of 6 byte loads.
// ============================================================================================
#ifndef VKB_REMAP_ROWS
#define VKB_REMAP_ROWS 4  // dst rows per thread of the remap kernel (4 or 8)
#endif
#ifndef VKB_REMAP_BLOCKS
// resident blocks per SM = the register cap: 3 x 256 threads leaves 85 registers (78 used, no
// spills); 4 blocks spill and measured 2 % slower, 2 blocks 19 % slower
#define VKB_REMAP_BLOCKS ((VKB_REMAP_ROWS == 4) ? 3 : 4)
#endif


// Rare paths are kept out of line so the hot loop stays small (instruction cache).
__device__ __noinline__ int2 cell_coord_exact(const double* __restrict__ H, int x, int y) {
    int X, Y;
    cell_coord(H, x, y, X, Y);
    return make_int2(X, Y);
}

// coverage of one over-budget cell on row y, restricted to the 32 columns starting at tx0
__device__ __noinline__ uint32_t cell_row_window_slow(const int32_t* __restrict__ lat, int cols,
                                                      int cell, int y, int tx0) {
    const int ccols = cols - 1;
    const int r = cell / ccols, c = cell - r * ccols;
    const int i00 = r * cols + c, i01 = i00 + 1, i11 = i00 + cols + 1, i10 = i00 + cols;
    const int px[4] = {lat[2 * i00], lat[2 * i01], lat[2 * i11], lat[2 * i10]};
    const int py[4] = {lat[2 * i00 + 1], lat[2 * i01 + 1], lat[2 * i11 + 1], lat[2 * i10 + 1]};
    uint32_t bits = 0;
    poly_row_mask<4>(px, py, y, tx0, &bits, 1);
    return bits;
}

// ---- cp.async plumbing of the persistent remap kernel -----------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) {
    return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void cp_async_16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// Persistent kernel, WARP-private work items: a block owns a contiguous share of the flat tile
// list and its warps take the tiles of that share round robin (neighbouring tiles at the same
// time, so their source rows meet in L1); a warp walks the R-row bands of its 32 x 32 tile by
// itself.  No block-level synchronisation exists: every warp stages the candidate records of its
// next tile with cp.async into its own half of a 2 x 32-record shared-memory buffer while it
// works on the current one (a tile with 33..64 records takes both halves and is loaded when it
// starts), so a warp never waits for another warp.
constexpr int kWarpSlots = 64;

// LARGE = false: the tiles with at most kPlaneCands candidates (99.8 % of them), owners resolved
// once per tile with lane = row; LARGE = true: the remaining tiles (more candidates, or the exact
// slow path), owners resolved band by band.  Two launches over the same tile list, each skipping
// the other's tiles, keep both kernels inside the register budget.
template <int C, bool MASK, bool SCORE, int R, bool LARGE>
__global__ void __launch_bounds__(32 * (VKB_TILE / R), VKB_REMAP_BLOCKS) grid_remap_kernel(
    const vkb_planes* __restrict__ planes, const vkb_grid_page* __restrict__ pages, int n_pages,
    int c_max, int p_max, const double* __restrict__ hinv, const int4* __restrict__ cell_box,
    const uint32_t* __restrict__ cell_masks, const int32_t* __restrict__ tile_base,
    const RemapTile* __restrict__ headers, const TileSlot* __restrict__ slots,
    const int32_t* __restrict__ lattice_i, const int32_t* __restrict__ large, int s_cap) {
    constexpr int kWarps = VKB_TILE / R;
    __shared__ __align__(16) TileSlot sm_all[kWarps][kWarpSlots];
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    TileSlot* __restrict__ sm = sm_all[warp];

    // LARGE = false: contiguous share of the flat tile list per block, round robin over its
    // warps.  LARGE = true: the (short, clustered) list of large tiles, strided over all warps.
    const int total = LARGE ? large[0] : tile_base[n_pages];
    const int per = (total + gridDim.x - 1) / gridDim.x;
    const int w_begin = LARGE ? 0 : blockIdx.x * per;
    const int n_block = LARGE ? total : min(total, w_begin + per) - w_begin;
    const int first = LARGE ? blockIdx.x * kWarps + warp : warp;
    const int stride = LARGE ? gridDim.x * kWarps : kWarps;
    const int n_tiles = n_block > first ? (n_block - first + stride - 1) / stride : 0;  // of this warp
    if (n_tiles <= 0) return;

    auto load_header = [&](int k) {
        RemapTile t;
        int index = w_begin + first + stride * min(k, n_tiles - 1);
        if (LARGE) index = large[1 + index];
        const int4* __restrict__ src = reinterpret_cast<const int4*>(headers + index);
        const int4 a = __ldg(src), b = __ldg(src + 1);
        t.page = a.x; t.tx0 = a.y; t.ty0 = a.z; t.count = a.w; t.tile = b.x; t.lim = b.y;
        return t;
    };
    // this warp's copies of the records of one tile's candidate cells, in the order of the
    // tile's sorted list (four 16-byte pieces per record)
    auto stage = [&](const RemapTile& t, int base) {
        const int chunks = t.count * (VKB_TILE_SLOT_BYTES / 16);  // <= 256; negative on the slow path
        const uint16_t* __restrict__ list = tile_list(slots, t.page, s_cap, c_max, t.tile);
        const char* g = reinterpret_cast<const char*>(slots + (size_t)t.page * s_cap);
        char* d = reinterpret_cast<char*>(sm + base);
        for (int i = lane; i < chunks; i += 32) {
            const int cell = (int)__ldg(list + (i >> 2));
            cp_async_16(d + i * 16, g + (size_t)cell * VKB_TILE_SLOT_BYTES + (i & 3) * 16);
        }
        cp_async_commit();
    };

    RemapTile h0 = load_header(0), h1 = load_header(1);
    int cur_base = 0;
    bool cur_staged = false;

    // per-page state, reloaded when the page changes
    int ctx_page = -1;
    int dst_h = 0, dst_w = 0, src_h = 0, src_w = 0, cols = 0;
    const uint8_t* __restrict__ src_image = nullptr;
    uint8_t* __restrict__ dst_image = nullptr;
    const uint8_t* __restrict__ src_mask = nullptr;
    uint8_t* __restrict__ dst_mask = nullptr;
    const float* __restrict__ src_score = nullptr;
    float* __restrict__ dst_score = nullptr;

    auto mine = [](const RemapTile& t) {
        return ((unsigned)t.count <= (unsigned)kPlaneCands) != LARGE;
    };
    for (int k = 0; k < n_tiles; ++k) {
        const RemapTile cur = h0;
        if (!mine(cur)) {  // the other launch's tile (never staged ahead)
            h0 = h1;
            h1 = load_header(k + 2);
            continue;
        }
        if (!cur_staged) {
            cur_base = 0;
            stage(cur, 0);
        }
        // the next tile's records go to the other half when both tiles fit a half
        const bool ahead = k + 1 < n_tiles && mine(h1) && cur.count <= kWarpSlots / 2
                           && h1.count <= kWarpSlots / 2;
        const int next_base = cur_base ? 0 : kWarpSlots / 2;
        if (ahead) {
            stage(h1, next_base);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncwarp();

        const int page = cur.page;
        if (page != ctx_page) {
            const vkb_planes* __restrict__ pl = planes + page;
            dst_h = pl->dst_h; dst_w = pl->dst_w; src_h = pl->src_h; src_w = pl->src_w;
            src_image = pl->src_image; dst_image = pl->dst_image;
            src_mask = pl->src_mask; dst_mask = pl->dst_mask;
            src_score = pl->src_score; dst_score = pl->dst_score;
            cols = pages[page].cols;
            ctx_page = page;
        }
        const TileSlot* __restrict__ S = sm + cur_base;
        const int tx0 = cur.tx0, ty0 = cur.ty0, count = cur.count, fast_lim = cur.lim;
        const bool fast = count >= 0;
        const size_t page_cell0 = (size_t)page * c_max;
        const int x = tx0 + lane;

        // ---- owner, whole tile at once: lane = dst row ------------------------------------
        // Bit plane b of a row holds bit b of (slot + 1) of every pixel's owner (0 = uncovered).
        // Candidates come in ascending cell order and later ones overwrite earlier ones, exactly
        // like the reference's cell-by-cell map writes; one coverage word per candidate and row.
        // Four planes cover tiles with up to 15 candidates (mean 8.7).
        uint32_t pl0 = 0, pl1 = 0, pl2 = 0, pl3 = 0;
        constexpr bool planes_ok = !LARGE;
        if (planes_ok) {
            const uint32_t* __restrict__ page_masks = cell_masks + page_cell0 * VKB_CELL_MASK_WORDS;
            const int row_y = ty0 + lane;
            for (int s = 0; s < count; ++s) {
                const int4 b = *reinterpret_cast<const int4*>(&S[s].x0);  // same for all lanes
                const int r = row_y - b.y;
                uint32_t win = 0u;
                if ((unsigned)r <= (unsigned)b.z) {
                    if (b.w >= 0) {
                        const uint32_t wd = __ldg(page_masks + (b.w * VKB_CELL_MASK_WORDS + r));
                        const int rel = tx0 - b.x;  // |rel| < 32: the bbox overlaps the tile
                        win = rel >= 0 ? (wd >> rel) : (wd << (-rel));
                    } else {
                        win = cell_row_window_slow(lattice_i + (size_t)page * p_max * 2, cols,
                                                   b.w & 0x7FFFFFFF, row_y, tx0);
                    }
                }
                const uint32_t id = (uint32_t)s + 1u;
                pl0 = (pl0 & ~win) | (win & (0u - (id & 1u)));
                pl1 = (pl1 & ~win) | (win & (0u - ((id >> 1) & 1u)));
                pl2 = (pl2 & ~win) | (win & (0u - ((id >> 2) & 1u)));
                pl3 = (pl3 & ~win) | (win & (0u - ((id >> 3) & 1u)));
            }
        }

#pragma unroll 1
        for (int band = 0; band < kWarps; ++band) {
        const int ry0 = ty0 + band * R;
        if (ry0 >= dst_h) break;

        int X[R], Y[R];
        if (ry0 < dst_h) {
            const uint32_t* __restrict__ page_masks = cell_masks + page_cell0 * VKB_CELL_MASK_WORDS;
            // ---- owner ---------------------------------------------------------------------
            // key: fast mode -> info of the owning slot; overflow mode -> cell index; -1 = none
            int key[R];
#pragma unroll
            for (int j = 0; j < R; ++j) key[j] = -1;
            if (planes_ok) {
#pragma unroll
                for (int j = 0; j < R; ++j) {
                    const int row = band * R + j;
                    uint32_t id = (__shfl_sync(0xffffffffu, pl0, row) >> lane) & 1u;
                    id |= ((__shfl_sync(0xffffffffu, pl1, row) >> lane) & 1u) << 1;
                    id |= ((__shfl_sync(0xffffffffu, pl2, row) >> lane) & 1u) << 2;
                    id |= ((__shfl_sync(0xffffffffu, pl3, row) >> lane) & 1u) << 3;
                    key[j] = (int)id - 1;
                }
            }
            const int n_cand = planes_ok ? 0 : (fast ? count : (pages[page].rows - 1) * (cols - 1));
            for (int base = 0; base < n_cand; base += 32) {
                const int s = base + lane;
                uint32_t win[R];
#pragma unroll
                for (int j = 0; j < R; ++j) win[j] = 0u;
                int my_key = s;
                if (s < n_cand) {
                    int4 b;
                    if (fast) {
                        b = *reinterpret_cast<const int4*>(&S[s].x0);  // records sit in rank order: key = s
                    } else {
                        const int4 g = cell_box[page_cell0 + s];
                        b = make_int4(g.x, g.y, g.w - g.y, s | ((g.z & 0x40000000) ? (int)0x80000000 : 0));
                    }
                    const int r = ry0 - b.y;    // band row 0 relative to the bbox
                    const int rel = tx0 - b.x;  // tile column 0 relative to the bbox
                    if (r + (R - 1) >= 0 && r <= b.z) {
                        if (b.w >= 0) {
                            if (fast || (rel > -32 && rel < 32)) {
                                const uint32_t* __restrict__ m = page_masks + (b.w * VKB_CELL_MASK_WORDS + r);
#pragma unroll
                                for (int j = 0; j < R; ++j) {
                                    if ((unsigned)(r + j) <= (unsigned)b.z) {
                                        const uint32_t wd = __ldg(m + j);
                                        win[j] = rel >= 0 ? (wd >> rel) : (wd << (-rel));
                                    }
                                }
                            }
                        } else {
                            const int32_t* lat = lattice_i + (size_t)page * p_max * 2;
#pragma unroll
                            for (int j = 0; j < R; ++j)
                                win[j] = cell_row_window_slow(lat, cols, b.w & 0x7FFFFFFF, ry0 + j, tx0);
                        }
                    }
                }
                uint32_t any = win[0];
#pragma unroll
                for (int j = 1; j < R; ++j) any |= win[j];
                unsigned active = __ballot_sync(0xffffffffu, any != 0u);
                while (active) {
                    const int src_lane = __ffs(active) - 1;
                    active &= active - 1;
                    const int kk = __shfl_sync(0xffffffffu, my_key, src_lane);
#pragma unroll
                    for (int j = 0; j < R; ++j) {
                        const uint32_t wd = __shfl_sync(0xffffffffu, win[j], src_lane);
                        if ((wd >> lane) & 1u) key[j] = kk;
                    }
                }
            }

            {
                // ---- coordinates -----------------------------------------------------------
                // uncovered pixels keep map value (0, 0); the fast path is evaluated for every
                // pixel (slot 0 for uncovered ones) and the result selected afterwards
                const float xr = (float)x;
                const float yr0 = (float)ry0;
#pragma unroll
                for (int j = 0; j < R; ++j) {
                    const bool covered = key[j] >= 0;
                    if (fast && covered) {
                        const int slot = key[j];
                        const int2 base = *reinterpret_cast<const int2*>(&S[slot].xm);
                        const float2 no = *reinterpret_cast<const float2*>(&S[slot].nox);
                        const bool ok = cell_coord_fast(S[slot].loc, xr + no.x, yr0 + (float)j + no.y,
                                                        base.x, base.y, fast_lim, X[j], Y[j]) >= 0;
                        if (covered && !ok) {
                            const int cell = S[slot].cellf & 0x7FFFFFFF;
                            const int2 e = cell_coord_exact(hinv + (page_cell0 + cell) * 9, x, ry0 + j);
                            X[j] = e.x;
                            Y[j] = e.y;
                        }
                    } else if (covered) {
                        const int2 e = cell_coord_exact(hinv + (page_cell0 + key[j]) * 9, x, ry0 + j);
                        X[j] = e.x;
                        Y[j] = e.y;
                    }
                    if (!covered) {
                        X[j] = 0;
                        Y[j] = 0;
                    }
                }
            }
        }
        if (ry0 < dst_h) {
            if (x < dst_w) {
                // ---- gather ----------------------------------------------------------------
                const int di0 = ry0 * dst_w + x;
                const bool tiny = src_h < 2 || src_w < 2;
                // tap weights of the R pixels, shared by Image and Mask; footprints that leave
                // the image are rare, so all of them are fixed up behind ONE branch
                TapWeights tw[R];
                bool outside = false;
#pragma unroll
                for (int j = 0; j < R; ++j) {
                    tw[j] = tap_weights_plain(X[j], Y[j]);
                    outside |= tap_outside(tw[j], src_h, src_w);
                }
                if (outside && !tiny) {
#pragma unroll
                    for (int j = 0; j < R; ++j)
                        if (tap_outside(tw[j], src_h, src_w)) tap_border_fix(tw[j], src_h, src_w);
                }
                if (C > 0) {
                    constexpr int CC = C > 0 ? C : 1;
                    uint8_t px[R][CC];
                    if (tiny) {
#pragma unroll
                        for (int j = 0; j < R; ++j) {
                            const uint32_t v = sample_u8_small<CC>(src_image, src_h, src_w, X[j], Y[j]);
#pragma unroll
                            for (int c = 0; c < CC; ++c) px[j][c] = (uint8_t)(v >> (8 * c));
                        }
                    } else {
                        Taps<CC> taps[R];
#pragma unroll
                        for (int j = 0; j < R; ++j) {
                            taps[j].t = tw[j];
                            taps_load<CC>(src_image, src_w, taps[j]);
                        }
#pragma unroll
                        for (int j = 0; j < R; ++j) taps_blend<CC>(taps[j], px[j]);
                    }
#pragma unroll
                    for (int j = 0; j < R; ++j) {
                        if (ry0 + j < dst_h) {
                            uint8_t* d = dst_image + (di0 + j * dst_w) * CC;
                            if (CC == 4) {
                                *reinterpret_cast<uchar4*>(d) =
                                    make_uchar4(px[j][0], px[j][1 % CC], px[j][2 % CC], px[j][3 % CC]);
                            } else {
#pragma unroll
                                for (int c = 0; c < CC; ++c) d[c] = px[j][c];
                            }
                        }
                    }
                }
                if (MASK) {
                    uint8_t m[R][1];
                    if (tiny) {
#pragma unroll
                        for (int j = 0; j < R; ++j)
                            m[j][0] = (uint8_t)sample_u8_small<1>(src_mask, src_h, src_w, X[j], Y[j]);
                    } else {
                        Taps<1> taps[R];
#pragma unroll
                        for (int j = 0; j < R; ++j) {
                            taps[j].t = tw[j];
                            taps_load<1>(src_mask, src_w, taps[j]);
                        }
#pragma unroll
                        for (int j = 0; j < R; ++j) taps_blend<1>(taps[j], m[j]);
                    }
#pragma unroll
                    for (int j = 0; j < R; ++j)
                        if (ry0 + j < dst_h) dst_mask[di0 + j * dst_w] = m[j][0];
                }
                if (SCORE) {
#pragma unroll
                    for (int j = 0; j < R; ++j) {
                        const float v = bilinear_f32(src_score, src_h, src_w, src_w, X[j], Y[j]);
                        if (ry0 + j < dst_h) dst_score[di0 + j * dst_w] = v;
                    }
                }
            }
        }
        }  // band
        __syncwarp();  // every lane is done with this tile's records before their half is reused
        cur_staged = ahead;
        cur_base = next_base;
        h0 = h1;
        h1 = load_header(k + 2);  // one exposed L2 latency per tile (~1 % of a tile's time)
    }
}

// ============================================================================================
// Small-tile kernel (second generation).  One WARP per 32 x 32 dst tile, persistent blocks, the
// records of the next tile's candidate cells staged with cp.async while the current one is
// processed (each lane reads its two candidate ids with the tile's header, four lanes copy one
// 64-byte cell record).  Per tile:
//
//   owner     lane = dst ROW: for every candidate (ascending cell order) the lane loads the
//             candidate's coverage word of its row and overwrites four bit planes of
//             "slot + 1" under it -- the last (largest) cell wins like the reference's
//             cell-by-cell map writes.  The four planes are then TRANSPOSED across the warp
//             (5 shuffle stages each), so that afterwards lane = dst COLUMN holds the owner bits
//             of its 32 rows; a band's four owners come out of one 16-bit word with a multiply
//             (no per-pixel shuffles).  Owner 0 = uncovered = record 0 of the warp's buffer,
//             a constant record whose map is (0, 0): uncovered pixels need no special case.
//   coords    float32 fast path per pixel with an integer acceptance test (vkb_math.cuh); a
//             pixel whose result sits next to a rounding boundary is NOT resolved on the spot:
//             it is flagged in a per-thread row mask and resolved after the band loop by the
//             float64 path (0.1 - 0.7 % of the pixels; on the spot the whole warp paid for it in
//             every other row).
//   gather    cv::remap's fixed-point bilinear with both blend directions folded into IDP.2A
//             (vkb_gather.cuh), loads of the band's four pixels in flight together.
//   Tiles without candidates and bands without owners are filled with the value of map (0, 0).
// ============================================================================================
#ifndef VKB_TILES_BLOCKS
#define VKB_TILES_BLOCKS 4
#endif
// Resident blocks per SM: 4 x 8 warps at <= 64 registers (measured 1.503 -> 1.447 ms per 256 pages
// against 3 blocks at 80 registers: the extra warps hide more of the tap-load latency); the
// image + mask + score map variants would spill at 64 registers and keep 3.
template <int C, bool MASK, bool SCORE>
constexpr int tiles_blocks_per_sm() {
    return (C >= 3 && MASK && SCORE) ? 3 : VKB_TILES_BLOCKS;
}
// Blocks per SM the launch actually uses (VKB_TILES_GRID_BLOCKS or the environment variable of the
// same name, experiments): fewer than the resident limit leaves block slots to kernels of other
// streams that run next to this persistent one.
#ifndef VKB_TILES_GRID_BLOCKS
#define VKB_TILES_GRID_BLOCKS 0
#endif
template <int C, bool MASK, bool SCORE>
static int tiles_grid_blocks_per_sm() {
    static const int from_env = [] {
        const char* e = getenv("VKB_TILES_GRID_BLOCKS");
        return e ? atoi(e) : 0;
    }();
    const int cap = tiles_blocks_per_sm<C, MASK, SCORE>();
    const int want = from_env > 0 ? from_env : (VKB_TILES_GRID_BLOCKS > 0 ? VKB_TILES_GRID_BLOCKS : cap);
    return want < cap ? want : cap;
}
constexpr int kTilesWarps = 8;
constexpr int kTilesHalf = 16;                  // records per half: [0] = the zero map, 1..15 candidates
constexpr int kTilesSlots = 2 * kTilesHalf;
#ifndef VKB_TILES_FREE_STORE
#define VKB_TILES_FREE_STORE 1
#endif
#ifndef VKB_TILES_ROWS
#define VKB_TILES_ROWS 4  // rows of a band one lane carries through coordinates -> gather together (4 or 2)
#endif
#ifndef VKB_TILES_SHARE_COLUMN
#define VKB_TILES_SHARE_COLUMN 1
#endif
#ifndef VKB_TILES_CHUNK
#define VKB_TILES_CHUNK 1
#endif
constexpr int kTilesChunk = VKB_TILES_CHUNK;  // consecutive tiles per draw from the work counter

__device__ __forceinline__ uint32_t rotr32(uint32_t v, int s) { return __funnelshift_r(v, v, s); }

// In: lane = row, bit = column of four 32 x 32 bit matrices.  Out: lane = column, bit = row.
__device__ __forceinline__ void warp_transpose4(uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d,
                                                int lane) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const uint32_t m = s == 16 ? 0x0000FFFFu : s == 8 ? 0x00FF00FFu : s == 4 ? 0x0F0F0F0Fu
                           : s == 2 ? 0x33333333u : 0x55555555u;
        // upper lanes keep their left blocks and take the partner's left blocks shifted left;
        // lower lanes keep their right blocks and take the partner's right blocks shifted right
        const bool upper = (lane & s) == 0;
        const uint32_t keep = upper ? m : ~m;
        const int rot = upper ? 32 - s : s;
        const uint32_t ya = __shfl_xor_sync(0xffffffffu, a, s);
        const uint32_t yb = __shfl_xor_sync(0xffffffffu, b, s);
        const uint32_t yc = __shfl_xor_sync(0xffffffffu, c, s);
        const uint32_t yd = __shfl_xor_sync(0xffffffffu, d, s);
        a = (a & keep) | (rotr32(ya, rot) & ~keep);
        b = (b & keep) | (rotr32(yb, rot) & ~keep);
        c = (c & keep) | (rotr32(yc, rot) & ~keep);
        d = (d & keep) | (rotr32(yd, rot) & ~keep);
    }
}

// everything a pixel needs from its page
struct RemapPage {
    int dst_h, dst_w, src_h, src_w;
    const uint8_t* __restrict__ src_image;
    uint8_t* __restrict__ dst_image;
    const uint8_t* __restrict__ src_mask;
    uint8_t* __restrict__ dst_mask;
    const float* __restrict__ src_score;
    float* __restrict__ dst_score;
};

// One pixel through the exact float64 coordinates (rare: 0.7 % of the pixels).  The gather is the
// hot path's packed one (aligned words, PRMT + IDP.2A) -- the same integers as the plain per-tap
// form, a third of its instructions; `packed` is false for planes the packed form does not
// take (smaller than 2 x 2, unaligned RGBA, RGB planes of 2^28 bytes and more).
template <int C, bool MASK, bool SCORE>
__device__ __noinline__ void remap_pixel_exact(const RemapPage pg, const double* __restrict__ H,
                                               int x, int y, bool packed) {
    constexpr int CC = C > 0 ? C : 1;
    int X, Y;
    cell_coord(H, x, y, X, Y);
    const long long di = (long long)y * pg.dst_w + x;
    if (packed) {
        const Tap2 t = tap2_make(X, Y, pg.src_h, pg.src_w);
        if (C > 0) {
            const uintptr_t a = reinterpret_cast<uintptr_t>(pg.src_image);
            const uint32_t* __restrict__ words = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
            const int mis = (int)(a & 3);
            Fetch2<CC> f;
            if constexpr (CC == 3) fetch2_request_rgb_bits(words, mis * 8, pg.src_w * 24, t, f);
            else fetch2_request<CC>(words, mis, pg.src_w * CC, t, f);
            uint32_t v[CC];
            fetch2_blend<CC>(f, t, v);
#pragma unroll
            for (int c = 0; c < CC; ++c) pg.dst_image[di * CC + c] = (uint8_t)v[c];
        }
        if (MASK) {
            const uintptr_t a = reinterpret_cast<uintptr_t>(pg.src_mask);
            Fetch2<1> f;
            fetch2_request<1>(reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3), (int)(a & 3), pg.src_w, t, f);
            uint32_t v[1];
            fetch2_blend<1>(f, t, v);
            pg.dst_mask[di] = (uint8_t)v[0];
        }
    } else {
        if (C > 0) {
            uint8_t px[CC];
            bilinear_u8<CC>(pg.src_image, pg.src_h, pg.src_w, (long long)pg.src_w * CC, X, Y, px);
#pragma unroll
            for (int c = 0; c < CC; ++c) pg.dst_image[di * CC + c] = px[c];
        }
        if (MASK) {
            uint8_t m[1];
            bilinear_u8<1>(pg.src_mask, pg.src_h, pg.src_w, (long long)pg.src_w, X, Y, m);
            pg.dst_mask[di] = m[0];
        }
    }
    if (SCORE) pg.dst_score[di] = bilinear_f32(pg.src_score, pg.src_h, pg.src_w, pg.src_w, X, Y);
}

// a tile header as the small-tile kernel carries it (uniform across the warp but `ids`)
struct TileHead {
    int page, tx0, ty0, count, lim;
    uint32_t ids;  // the lane's two candidate cells
};

template <int C, bool MASK, bool SCORE>
__global__ void __launch_bounds__(32 * kTilesWarps, (tiles_blocks_per_sm<C, MASK, SCORE>())) grid_remap_tiles_kernel(
    const vkb_planes* __restrict__ planes, const vkb_grid_page* __restrict__ pages, int n_pages,
    int c_max, int p_max, const double* __restrict__ hinv, const uint32_t* __restrict__ cell_masks,
    const int32_t* __restrict__ tile_base, const RemapTile* __restrict__ headers,
    const TileSlot* __restrict__ slots, const int32_t* __restrict__ lattice_i,
    int32_t* __restrict__ work_counter, int s_cap) {
    constexpr int CC = C > 0 ? C : 1;
    __shared__ __align__(16) TileSlot sm_all[kTilesWarps][kTilesSlots];
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    TileSlot* __restrict__ sm = sm_all[warp];

    // Dynamic work distribution: a warp draws chunks of kTilesChunk consecutive tiles from a
    // global counter (tiles without owners cost a fraction of a full tile, so static shares end
    // unevenly: 13 % of a 64-page launch was tail).  Two chunks are always in hand, so the next
    // tile and the one after it are known for the header / record prefetch.
    const int total = tile_base[n_pages];
    // The counter is drawn one chunk further ahead than it is needed: lane 0 keeps the raw result
    // of its atomicAdd (`pending`) and broadcasts it only when the chunk after next is due, a whole
    // tile later -- the warp never waits for the atomic (8 % of the stall samples before).
    auto draw = [&]() { return lane == 0 ? atomicAdd(work_counter, kTilesChunk) : 0; };
    int chunk0 = __shfl_sync(0xffffffffu, draw(), 0);
    int chunk1 = __shfl_sync(0xffffffffu, draw(), 0);
    int pending = draw(), pos = 0;
    if (chunk0 >= total) return;
    auto tile_at = [&](int ahead) {  // ahead <= kTilesChunk
        const int p = pos + ahead;
        return p < kTilesChunk ? chunk0 + p : chunk1 + (p - kTilesChunk);
    };

    // record 0 of both halves: the map of uncovered pixels, (0, 0) for every pixel, always exact
    if (lane < 2) {
        TileSlot z = {};
        z.xm = kFastBaseZero;
        z.ym = kFastBaseZero;
        z.nox = z.noy = 0.f;
        sm[lane * kTilesHalf] = z;
    }
    __syncwarp();

    auto load_header = [&](int index) {
        TileHead t;
        index = min(index, total - 1);
        const int4* __restrict__ src = reinterpret_cast<const int4*>(headers + index);
        const int4 a = __ldg(src), b = __ldg(src + 1);
        t.page = a.x; t.tx0 = a.y; t.ty0 = a.z; t.count = a.w; t.lim = b.y;
        // the lane's two candidates (lane / 4 and lane / 4 + 8 of the sorted list), requested with
        // the header: staging needs no second trip to memory
        t.ids = __ldg(reinterpret_cast<const uint32_t*>(src + 2) + (lane >> 2));
        return t;
    };
    auto mine = [](const TileHead& t) { return (unsigned)t.count <= (unsigned)kPlaneCands; };
    // The records of the tile's candidate cells go behind the half's record 0, in list order:
    // four lanes copy one 64-byte record, two rounds cover the 15 candidates.
    auto stage = [&](const TileHead& t, int base) {
        const char* g = reinterpret_cast<const char*>(slots + (size_t)t.page * s_cap) + (lane & 3) * 16;
        char* d = reinterpret_cast<char*>(sm + base + 1) + lane * 16;
        if ((lane >> 2) < t.count)
            cp_async_16(d, g + (size_t)(t.ids & 0xFFFFu) * VKB_TILE_SLOT_BYTES);
        if ((lane >> 2) + 8 < t.count)
            cp_async_16(d + 8 * VKB_TILE_SLOT_BYTES, g + (size_t)(t.ids >> 16) * VKB_TILE_SLOT_BYTES);
        cp_async_commit();
    };

    TileHead h0 = load_header(tile_at(0)), h1 = load_header(tile_at(1));
    int cur_base = 0;
    bool cur_staged = false;
    auto advance = [&]() {  // next tile of this warp's sequence
        if (++pos == kTilesChunk) {
            pos = 0;
            chunk0 = chunk1;
            chunk1 = __shfl_sync(0xffffffffu, pending, 0);
            pending = draw();
        }
        h0 = h1;
        h1 = load_header(tile_at(1));
    };

    // per-page state, reloaded when the page changes
    int ctx_page = -1;
    RemapPage pg = {};
    int cols = 0;
    const uint32_t* __restrict__ img_words = nullptr;   // image base rounded down to 4 bytes
    const uint32_t* __restrict__ mask_words = nullptr;  // mask base rounded down to 4 bytes
    int img_mis = 0, mask_mis = 0, img_pitch = 0;
    bool tiny = false;
    uint32_t fill_px[CC] = {};  // value of map (0, 0): what uncovered pixels receive
    uint32_t fill_mask = 0;
    float fill_score = 0.f;

    while (tile_at(0) < total) {
        const TileHead cur = h0;
        if (!mine(cur)) {  // the large-tile launch's tile (never staged ahead)
            advance();
            continue;
        }
        if (!cur_staged) {
            cur_base = 0;
            stage(cur, 0);
        }
        const bool ahead = tile_at(1) < total && mine(h1);
        const int next_base = cur_base ? 0 : kTilesHalf;
        if (ahead) {
            stage(h1, next_base);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncwarp();

        const int page = cur.page;
        if (page != ctx_page) {
            const vkb_planes* __restrict__ pl = planes + page;
            pg.dst_h = pl->dst_h; pg.dst_w = pl->dst_w; pg.src_h = pl->src_h; pg.src_w = pl->src_w;
            pg.src_image = pl->src_image; pg.dst_image = pl->dst_image;
            pg.src_mask = pl->src_mask; pg.dst_mask = pl->dst_mask;
            pg.src_score = pl->src_score; pg.dst_score = pl->dst_score;
            cols = pages[page].cols;
            tiny = pg.src_h < 2 || pg.src_w < 2;
            if (C > 0) {
                const uintptr_t a = reinterpret_cast<uintptr_t>(pg.src_image);
                img_mis = (int)(a & 3);
                img_words = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
                img_pitch = pg.src_w * CC;
                if (CC == 4 && img_mis) tiny = true;  // unaligned RGBA: the plain per-tap path
                if (CC == 3) {  // RGB taps are addressed in bits (fetch2_request_rgb_bits)
                    if ((long long)img_pitch * pg.src_h >= (1ll << 28)) tiny = true;  // plain per-tap path
                    img_mis *= 8;
                    img_pitch *= 8;
                }
                uint8_t px[CC];
                bilinear_u8<CC>(pg.src_image, pg.src_h, pg.src_w, (long long)pg.src_w * CC, 0, 0, px);
#pragma unroll
                for (int c = 0; c < CC; ++c) fill_px[c] = px[c];
            }
            if (MASK) {
                const uintptr_t a = reinterpret_cast<uintptr_t>(pg.src_mask);
                mask_mis = (int)(a & 3);
                mask_words = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
                uint8_t m[1];
                bilinear_u8<1>(pg.src_mask, pg.src_h, pg.src_w, (long long)pg.src_w, 0, 0, m);
                fill_mask = m[0];
            }
            if (SCORE) fill_score = bilinear_f32(pg.src_score, pg.src_h, pg.src_w, pg.src_w, 0, 0);
            ctx_page = page;
        }
        const TileSlot* __restrict__ S = sm + cur_base;  // S[0]: zero map, S[1 ..]: candidates
        const int tx0 = cur.tx0, ty0 = cur.ty0, count = cur.count, fast_lim = cur.lim;
        const size_t page_cell0 = (size_t)page * c_max;
        const int x = tx0 + lane;
        const bool x_in = x < pg.dst_w;
        const bool tile_in_x = tx0 + VKB_TILE <= pg.dst_w;

        // stores of one finished pixel
        auto store_px = [&](int di, const uint32_t* v) {
            uint8_t* d = pg.dst_image + (long long)di * CC;
            if (CC == 4) {
                *reinterpret_cast<uint32_t*>(d) = v[0] | (v[1 % CC] << 8) | (v[2 % CC] << 16) | (v[3 % CC] << 24);
            } else {
#pragma unroll
                for (int c = 0; c < CC; ++c) d[c] = (uint8_t)v[c];
            }
        };
        auto fill_rows = [&](int ry0, int n_rows) {  // rows ry0 .. ry0 + n_rows - 1 take map (0, 0)
            if (!x_in) return;
            int di = ry0 * pg.dst_w + x;
            for (int j = 0; j < n_rows; ++j, di += pg.dst_w) {
                if (ry0 + j >= pg.dst_h) break;
                if (C > 0) store_px(di, fill_px);
                if (MASK) pg.dst_mask[di] = (uint8_t)fill_mask;
                if (SCORE) pg.dst_score[di] = fill_score;
            }
        };

        if (count == 0) {
            fill_rows(ty0, VKB_TILE);
        } else {
        // ---- owner planes, lane = dst row --------------------------------------------------
        uint32_t q0 = 0, q1 = 0, q2 = 0, q3 = 0;
        {
            const uint32_t* __restrict__ page_masks = cell_masks + page_cell0 * VKB_CELL_MASK_WORDS;
            const int row_y = ty0 + lane;
            // four candidates per step: their coverage words are requested together, so a tile
            // pays ~count / 4 exposed memory latencies instead of count (1.99 -> 1.89 ms)
            for (int s0 = 1; s0 <= count; s0 += 4) {
                uint32_t wd[4];
                int rel[4], cellf[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    wd[u] = 0u;
                    rel[u] = 0;
                    cellf[u] = 0;
                    if (s0 + u <= count) {
                        const int4 b = *reinterpret_cast<const int4*>(&S[s0 + u].x0);  // same for all lanes
                        const int r = row_y - b.y;
                        rel[u] = tx0 - b.x;  // |rel| < 32: the bbox overlaps the tile
                        cellf[u] = b.w;
                        if ((unsigned)r <= (unsigned)b.z && b.w >= 0)
                            wd[u] = __ldg(page_masks + (b.w * VKB_CELL_MASK_WORDS + r));
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint32_t s = (uint32_t)(s0 + u);  // past `count`: wd = 0, nothing changes
                    uint32_t win = rel[u] >= 0 ? (wd[u] >> rel[u]) : (wd[u] << (-rel[u]));
                    if (cellf[u] < 0)  // over-budget cell (rare, uniform): rasterised here
                        win = cell_row_window_slow(lattice_i + (size_t)page * p_max * 2, cols,
                                                   cellf[u] & 0x7FFFFFFF, row_y, tx0);
                    q0 = (q0 & ~win) | (win & (0u - (s & 1u)));
                    q1 = (q1 & ~win) | (win & (0u - ((s >> 1) & 1u)));
                    q2 = (q2 & ~win) | (win & (0u - ((s >> 2) & 1u)));
                    q3 = (q3 & ~win) | (win & (0u - ((s >> 3) & 1u)));
                }
            }
        }
        // ---- lane = dst column from here on -------------------------------------------------
        warp_transpose4(q0, q1, q2, q3, lane);
        // plane p rotated left by 4p: the owner bits of a band's four rows sit in nibble p
        q1 = rotr32(q1, 28);
        q2 = rotr32(q2, 24);
        q3 = rotr32(q3, 20);
        uint32_t failbits = 0;  // rows of this column left to the float64 path
        const float xr = (float)x;  // absolute column: a cell's form takes pixel - bbox origin

#pragma unroll 1
        for (int band = 0; band < VKB_TILE / 4; ++band) {
            const int ry0 = ty0 + band * 4;
            // nibble p of `own` = plane p's bits of rows ry0 .. ry0 + 3
            const uint32_t own = (q0 & 0xFu) | (q1 & 0xF0u) | (q2 & 0xF00u) | (q3 & 0xF000u);
            q0 = rotr32(q0, 4);
            q1 = rotr32(q1, 4);
            q2 = rotr32(q2, 4);
            q3 = rotr32(q3, 4);
            if (ry0 >= pg.dst_h) continue;  // (keeps the planes' rotation count at 8)
            if (!__any_sync(0xffffffffu, own != 0u)) {
                fill_rows(ry0, 4);
                continue;
            }
            // ---- coordinates ------------------------------------------------------------
            constexpr int R = VKB_TILES_ROWS;
            static_assert(R == 4 || R == 2, "VKB_TILES_ROWS must be 4 or 2");
#pragma unroll 1
            for (int sub = 0; sub < 4 / R; ++sub) {
            const int jb = sub * R;  // first row of the sub-band inside the band
            if (R < 4 && ry0 + jb >= pg.dst_h) break;
            int X[R], Y[R];
            uint32_t fail4 = 0;
            const float yr0 = (float)(ry0 + jb);
#if VKB_TILES_SHARE_COLUMN
            // Nearly every band lies inside one row of cells: every lane then has ONE owner for its
            // four rows (each plane's nibble is 0 or F), and the column part of the cell's three
            // linear forms, its record pointer and its bases are computed once per lane.
            if (R == 4 && __all_sync(0xffffffffu, ((own & 0x1111u) * 0xFu) == own)) {
                const uint32_t id = ((own & 0x1111u) * 0x12480000u) >> 28;
                const TileSlot* __restrict__ sp = S + id;
                const int4 base = *reinterpret_cast<const int4*>(&sp->xm);  // xm, ym, nox, noy
                CellColumn col;
                cell_column(sp->loc, xr + __int_as_float(base.z), col);
                const float a1 = sp->loc.a1, b1 = sp->loc.b1, hh = sp->loc.h;
                const float yc = yr0 + __int_as_float(base.w);
                // verdicts collected from the last row down: bit j of fail4 = row j rejected
#pragma unroll
                for (int j = R - 1; j >= 0; --j)
                    fail4 = fast_collect(fail4, cell_coord_fast_row(col, a1, b1, hh, yc + (float)j,
                                                                    base.x, base.y, fast_lim, X[j], Y[j]));
            } else
#endif
            {
#pragma unroll
            for (int j = R - 1; j >= 0; --j) {
                // bits j, j+4, j+8, j+12 of `own` -> a 4-bit number (the partial products of
                // the multiplication land on distinct bits: no carries)
                const uint32_t id = (((own >> (jb + j)) & 0x1111u) * 0x12480000u) >> 28;
                const TileSlot* __restrict__ sp = S + id;
                const int4 base = *reinterpret_cast<const int4*>(&sp->xm);  // xm, ym, nox, noy
                fail4 = fast_collect(fail4, cell_coord_fast(sp->loc, xr + __int_as_float(base.z),
                                                            yr0 + (float)j + __int_as_float(base.w),
                                                            base.x, base.y, fast_lim, X[j], Y[j]));
            }
            }
            failbits |= fail4 << (band * 4 + jb);
            // rows of the band this thread stores: the ones inside the page.  (A pixel that waits
            // for the exact path is stored here all the same and overwritten after the band loop
            // by this very thread, in program order: no predicate per pixel for it.)  Bands that
            // lie inside the page completely -- warp uniform, nearly all of them -- store
            // without any predicate.
            const int rows_in = pg.dst_h - (ry0 + jb);  // >= 1
            constexpr uint32_t kAllRows = (1u << R) - 1u;
#if VKB_TILES_FREE_STORE
            const bool interior = tile_in_x && rows_in >= R;
            const uint32_t live = x_in ? (rows_in >= R ? kAllRows : ((1u << rows_in) - 1u)) : 0u;
#else
            const bool interior = false;
            const uint32_t live = x_in ? (~fail4 & (rows_in >= R ? kAllRows : ((1u << rows_in) - 1u))) : 0u;
#endif
            // ---- gather -------------------------------------------------------------------
            const int di0 = (ry0 + jb) * pg.dst_w + x;
            if (tiny) {
#pragma unroll
                for (int j = 0; j < R; ++j) {
                    if (!((live >> j) & 1u)) continue;
                    const int di = di0 + j * pg.dst_w;
                    if (C > 0) {
                        const uint32_t v = sample_u8_small<CC>(pg.src_image, pg.src_h, pg.src_w, X[j], Y[j]);
                        uint32_t ch[CC];
#pragma unroll
                        for (int c = 0; c < CC; ++c) ch[c] = (v >> (8 * c)) & 0xFFu;
                        store_px(di, ch);
                    }
                    if (MASK) pg.dst_mask[di] = (uint8_t)sample_u8_small<1>(pg.src_mask, pg.src_h, pg.src_w, X[j], Y[j]);
                    if (SCORE) pg.dst_score[di] = bilinear_f32(pg.src_score, pg.src_h, pg.src_w, pg.src_w, X[j], Y[j]);
                }
                continue;
            }
            // taps -> loads -> blend -> stores of the band's R rows, for a given set of taps
            auto emit = [&](const Tap2 (&tap)[R]) {
                if (C > 0) {
                    Fetch2<CC> f[R];
#pragma unroll
                    for (int j = 0; j < R; ++j) {
                        if constexpr (CC == 3) fetch2_request_rgb_bits(img_words, img_mis, img_pitch, tap[j], f[j]);
                        else fetch2_request<CC>(img_words, img_mis, img_pitch, tap[j], f[j]);
                    }
                    uint32_t v[R][CC];
#pragma unroll
                    for (int j = 0; j < R; ++j) fetch2_blend<CC>(f[j], tap[j], v[j]);
                    if (interior) {
#pragma unroll
                        for (int j = 0; j < R; ++j) store_px(di0 + j * pg.dst_w, v[j]);
                    } else {
#pragma unroll
                        for (int j = 0; j < R; ++j)
                            if ((live >> j) & 1u) store_px(di0 + j * pg.dst_w, v[j]);
                    }
                }
                if (MASK) {
                    Fetch2<1> f[R];
#pragma unroll
                    for (int j = 0; j < R; ++j) fetch2_request<1>(mask_words, mask_mis, pg.src_w, tap[j], f[j]);
#pragma unroll
                    for (int j = 0; j < R; ++j) {
                        uint32_t v[1];
                        fetch2_blend<1>(f[j], tap[j], v);
                        if ((live >> j) & 1u) pg.dst_mask[di0 + j * pg.dst_w] = (uint8_t)v[0];
                    }
                }
                if (SCORE) {
#pragma unroll
                    for (int j = 0; j < R; ++j) {
                        const float v = bilinear_f32(pg.src_score, pg.src_h, pg.src_w, pg.src_w, X[j], Y[j]);
                        if ((live >> j) & 1u) pg.dst_score[di0 + j * pg.dst_w] = v;
                    }
                }
            };
            Tap2 tap[R];
            bool outside = false;
#pragma unroll
            for (int j = 0; j < R; ++j) {
                tap[j] = tap2_plain(X[j], Y[j]);
                outside |= tap2_outside(tap[j], pg.src_h, pg.src_w);
            }
            if (outside) {  // rare: footprints that leave the image; its own copy of the band's
                            // code, so that the common path carries no merged values
                Tap2 tb[R];
#pragma unroll
                for (int j = 0; j < R; ++j)
                    tb[j] = tap2_outside(tap[j], pg.src_h, pg.src_w) ? tap2_border(X[j], Y[j], pg.src_h, pg.src_w)
                                                                     : tap[j];
                emit(tb);
            } else {
                emit(tap);
            }
            }  // sub-band
        }  // band

        // ---- the flagged pixels, float64 coordinates (q0 .. q3 are back in place: 8 x 4 bits) ----
        while (failbits) {
            const int row = __ffs(failbits) - 1;
            failbits &= failbits - 1;
            const uint32_t id = ((q0 >> row) & 1u) | (((rotr32(q1, 4) >> row) & 1u) << 1)
                                | (((rotr32(q2, 8) >> row) & 1u) << 2)
                                | (((rotr32(q3, 12) >> row) & 1u) << 3);
            const int cell = S[id].cellf & 0x7FFFFFFF;  // id >= 1: the zero map never fails
            if (x_in && ty0 + row < pg.dst_h)
                remap_pixel_exact<C, MASK, SCORE>(pg, hinv + (page_cell0 + cell) * 9, x, ty0 + row, !tiny);
        }
        }  // count != 0
        __syncwarp();  // every lane is done with this tile's records before their half is reused
        cur_staged = ahead;
        cur_base = next_base;
        advance();
    }
}

}  // namespace vkb

using namespace vkb;

static int remap_grid_blocks(int blocks_per_sm) {
    static int sm_count = 0;
    if (sm_count == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess
            || cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess
            || sm_count <= 0)
            sm_count = 148;
    }
    return sm_count * blocks_per_sm;
}

// One side stream + fork / join events per device, created on first use (the only state the
// library keeps besides the colour tables; work submitted through it is ordered with the
// caller's stream by the two events, so the call stays stream-ordered for the caller).
struct RemapSide {
    cudaStream_t stream;
    cudaEvent_t fork, join;
};
static RemapSide* remap_side() {
    static RemapSide sides[64];
    static bool ready[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    if (!ready[dev]) {
        RemapSide s;
        if (cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        sides[dev] = s;
        ready[dev] = true;
    }
    return &sides[dev];
}

extern "C" int vkb_grid_remap(const vkb_grid_page* pages, const vkb_planes* planes, int32_t n_pages,
                              int32_t p_max, int32_t c_max, int32_t t_max, int32_t s_cap,
                              const int32_t* lattice_i, const double* hinv, const int32_t* cell_box,
                              const uint32_t* cell_masks, const int32_t* tile_count,
                              const int32_t* tile_off, const int32_t* tile_base,
                              const void* tile_slots, void* tile_headers,
                              int32_t image_channels, int32_t has_mask, int32_t has_score,
                              void* stream) {
    VKB_NVTX("vkb_grid_remap");
    VKB_REQUIRE(pages && planes && lattice_i && hinv && cell_box && cell_masks && tile_count
                    && tile_base && tile_slots && tile_headers, "bad arguments");
    (void)tile_off;
    const int32_t* large = reinterpret_cast<const int32_t*>(
        reinterpret_cast<const char*>(tile_headers) + (size_t)n_pages * t_max * VKB_TILE_HEADER_BYTES);
    VKB_REQUIRE(n_pages > 0 && n_pages <= 65535, "1..65535 pages per launch");
    // the work counter of the small-tile kernel lives behind the list of large tiles
    int32_t* work_counter = const_cast<int32_t*>(large) + (size_t)n_pages * t_max + 1;
    VKB_CUDA(cudaMemsetAsync(work_counter, 0, sizeof(int32_t), (cudaStream_t)stream));
    VKB_REQUIRE(image_channels == 0 || image_channels == 1 || image_channels == 3
                    || image_channels == 4, "image_channels must be 0, 1, 3 or 4");
    VKB_REQUIRE(image_channels || has_mask || has_score, "nothing to remap");
    cudaStream_t st = (cudaStream_t)stream;
    constexpr int R = VKB_REMAP_ROWS;
    constexpr int kBlocksPerSm = VKB_REMAP_BLOCKS;
    const int grid = remap_grid_blocks(kBlocksPerSm);
    // VKB_REMAP_V1=1 keeps the first-generation small-tile kernel (A/B measurements only)
    static const bool use_v1 = getenv("VKB_REMAP_V1") != nullptr && getenv("VKB_REMAP_V1")[0] == '1';
#define VKB_LAUNCH_REMAP_1(CH, M, S, LARGE)                                                     \
    grid_remap_kernel<CH, M, S, R, LARGE><<<grid, 32 * (VKB_TILE / R), 0, st>>>(               \
        planes, pages, n_pages, c_max, p_max, hinv, reinterpret_cast<const int4*>(cell_box),   \
        cell_masks, tile_base, reinterpret_cast<const RemapTile*>(tile_headers),               \
        reinterpret_cast<const TileSlot*>(tile_slots), lattice_i, large, s_cap)
    // The few large tiles run on a side stream next to the main launch (disjoint dst tiles):
    // alone they are a latency-bound tail of ~30 us.
    RemapSide* side = remap_side();
    cudaStream_t main_st = st;
    // fork / launch / join are issued under a lock: the events are shared by all callers
    static std::mutex side_mutex;
    std::lock_guard<std::mutex> side_lock(side_mutex);
    if (side) {
        VKB_CUDA(cudaEventRecord(side->fork, main_st));
        VKB_CUDA(cudaStreamWaitEvent(side->stream, side->fork, 0));
    }
#define VKB_LAUNCH_REMAP(CH, M, S)                                                              \
    do {                                                                                        \
        st = side ? side->stream : main_st;                                                     \
        VKB_LAUNCH_REMAP_1(CH, M, S, true);                                                     \
        st = main_st;                                                                           \
        if (use_v1) {                                                                           \
            VKB_LAUNCH_REMAP_1(CH, M, S, false);                                                \
        } else {                                                                                \
            grid_remap_tiles_kernel<CH, M, S><<<remap_grid_blocks(tiles_grid_blocks_per_sm<CH, M, S>()), \
                                              32 * kTilesWarps, 0, st>>>(                       \
                planes, pages, n_pages, c_max, p_max, hinv, cell_masks, tile_base,              \
                reinterpret_cast<const RemapTile*>(tile_headers),                               \
                reinterpret_cast<const TileSlot*>(tile_slots), lattice_i, work_counter, s_cap); \
        }                                                                                       \
    } while (0)
    const int key = image_channels * 4 + (has_mask ? 2 : 0) + (has_score ? 1 : 0);
    switch (key) {
#ifdef VKB_REMAP_ONLY_RGB  // kernel experiments (tools/build_variants.py): one instantiation
        case 3 * 4 + 0: VKB_LAUNCH_REMAP(3, false, false); break;
        default: VKB_REQUIRE(false, "experiment build: RGB image only");
    }
#else
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
#endif
#undef VKB_LAUNCH_REMAP
#undef VKB_LAUNCH_REMAP_1
    if (side) {
        VKB_CUDA(cudaEventRecord(side->join, side->stream));
        VKB_CUDA(cudaStreamWaitEvent(main_st, side->join, 0));
    }
    return check_launch("grid_remap_kernel");
}
