// vkb_grid.cuh -- records shared by the grid-op kernels (geometric.cu builds them, remap.cu reads
// them): per-tile candidate records and the flat tile work list of the fused remap.
#pragma once
#include <stdint.h>
#include "../../include/vkit_b200.h"
#include "vkb_math.cuh"

namespace vkb {

// One record per lattice cell (grid_cells_kernel writes it, the remap stages the records of a
// tile's candidate cells into shared memory): bbox, coverage-mask flag and the float32 form of
// the cell's inverse homography re-centred on the bbox origin.
struct __align__(16) TileSlot {
    CellLocal loc;
    int x0, y0, nr, cellf;  // bbox origin, rows - 1, cell | (over mask budget ? 1 << 31 : 0)
    int xm, ym;             // fast_base(src corner of the cell, margin of the cell): base of the fast path
    float nox, noy;         // -x0, -y0: pixel + (nox, noy) = the argument of `loc`
};
static_assert(sizeof(TileSlot) == VKB_TILE_SLOT_BYTES, "TileSlot layout is part of the ABI");

// Tiles with at most this many candidates resolve their owners once per tile (lane = row, four
// bit planes); the others go on a separate work list.
constexpr int kPlaneCands = 15;

// One work item of the persistent remap kernel: a 32 x 32 dst tile (uniform across a warp).
struct __align__(16) RemapTile {
    int page, tx0, ty0;
    int count;  // candidate cells; -1: the tile takes the slow exact path
    int tile;   // index of the tile within its page (its sorted candidate list: tile_list())
    int lim;    // fast_limit of the largest acceptance margin among the tile's candidates
    int pad[2];
    // the first 16 candidates in ascending cell order, packed for the small-tile kernel:
    // ids[w] = candidate w | candidate (w + 8) << 16
    uint32_t ids[8];
};
static_assert(sizeof(RemapTile) == VKB_TILE_HEADER_BYTES, "RemapTile layout is part of the ABI");

// Layout of the `tile_slots` workspace of one page (s_cap records of 64 bytes):
//   records 0 .. c_max-1         the cells' TileSlot records
//   records c_max .. c_max+2*t   per tile VKB_TILE_CAP uint16: its candidates in ascending order
__host__ __device__ __forceinline__ const uint16_t* tile_list(const TileSlot* slots, int page, int s_cap,
                                                             int c_max, int tile) {
    return reinterpret_cast<const uint16_t*>(slots + ((size_t)page * s_cap + c_max)) + (size_t)tile * VKB_TILE_CAP;
}

__device__ __forceinline__ int page_tiles(const vkb_grid_meta& m) {
    return ((m.dst_w + VKB_TILE - 1) / VKB_TILE) * ((m.dst_h + VKB_TILE - 1) / VKB_TILE);
}

}  // namespace vkb
