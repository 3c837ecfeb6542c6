// vkb_grid.cuh -- records shared by the grid-op kernels (geometric.cu builds them, remap.cu reads
// them): per-tile candidate records and the flat tile work list of the fused remap.
#pragma once
#include <stdint.h>
#include "../../include/vkit_b200.h"
#include "vkb_math.cuh"

namespace vkb {

struct __align__(16) TileSlot {
    CellLocal loc;
    int x0, y0, nr, cellf;  // bbox origin, rows - 1, cell | (over mask budget ? 1 << 31 : 0)
    int xm, ym;             // fast_base(src corner of the cell, margin of the page): base of the fast path
    int info;               // slot | cell column << 6 | cell row << 16
    int pad;
};
static_assert(sizeof(TileSlot) == VKB_TILE_SLOT_BYTES, "TileSlot layout is part of the ABI");

// Tiles with at most this many candidates resolve their owners once per tile (lane = row, four
// bit planes); the others go on a separate work list.
constexpr int kPlaneCands = 15;

// One work item of the persistent remap kernel: a 32 x 32 dst tile (uniform across a warp).
struct __align__(16) RemapTile {
    int page, tx0, ty0;
    int count;  // candidate records; -1: the tile takes the slow exact path
    int rec;    // index of the first record
    int lim;    // fast_limit of the tile's acceptance margin (the records' bases hold the margin)
    int pad[2];
};
static_assert(sizeof(RemapTile) == VKB_TILE_HEADER_BYTES, "RemapTile layout is part of the ABI");

__device__ __forceinline__ int page_tiles(const vkb_grid_meta& m) {
    return ((m.dst_w + VKB_TILE - 1) / VKB_TILE) * ((m.dst_h + VKB_TILE - 1) / VKB_TILE);
}

}  // namespace vkb
