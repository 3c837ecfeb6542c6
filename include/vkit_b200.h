/* vkit_b200.h -- C ABI of libvkit_b200.so (sm_100a).
 *
 * The reference (vkit-x/vkit @ 98ada2d) has no FFI: its plugin boundary is the Python
 * `Distortion(config_cls, state_cls, func_image, func_mask, func_score_map, ...)` object
 * (vkit/mechanism/distortion/interface.py:141-248).  The entry points below are what the
 * `func_*` callables and `state_cls` constructors of that object bind to once the NumPy/cv2
 * arithmetic underneath them is replaced; each one cites the reference code it stands in for.
 * INTEGRATION.md shows the ctypes stubs.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in `_host`;
 *   - the caller owns all memory (outputs and workspaces are allocated by the caller);
 *   - calls are asynchronous on `stream` (a cudaStream_t passed as void*);
 *   - return value 0 = ok, negative = error; `vkb_last_error()` gives the message
 *     (thread local);
 *   - images are dense HWC uint8 (C = 1, 3 or 4), masks dense HW uint8, score maps dense HW
 *     float32 -- the layouts of `Image.mat`, `Mask.mat`, `ScoreMap.mat`
 *     (vkit/element/image.py:217-258, mask.py:71-88, score_map.py:86-108).
 */
#ifndef VKIT_B200_H_
#define VKIT_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VKB_OK 0
#define VKB_ERR_INVALID (-1)
#define VKB_ERR_CUDA (-2)

int vkb_version(void);
const char* vkb_last_error(void);

/* ---------------------------------------------------------------------------------------
 * Containers that travel together through one geometric op (Image + Mask + ScoreMap of one
 * page).  Absent containers have NULL pointers / image_channels == 0.
 * ------------------------------------------------------------------------------------- */
typedef struct vkb_planes {
    const uint8_t* src_image;
    uint8_t* dst_image;
    const uint8_t* src_mask;
    uint8_t* dst_mask;
    const float* src_score;
    float* dst_score;
    int32_t image_channels;
    int32_t src_h, src_w;
    int32_t dst_h, dst_w;
    int32_t _pad;
} vkb_planes;

/* ---------------------------------------------------------------------------------------
 * Affine family: rotate / shear_hori / shear_vert (2x3, cv.warpAffine) and skew_hori /
 * skew_vert (3x3, cv.warpPerspective).  Replaces affine_mat + affine_trait_func_image /
 * _mask / _score_map (vkit/mechanism/distortion/geometric/affine.py:38-43, 416-456).
 * `inv` is the INVERSE map (dst -> src) in double, row major; for kind 0 only inv[0..5]
 * are used.  One launch warps all present containers of all pages.
 * ------------------------------------------------------------------------------------- */
#define VKB_WARP_AFFINE 0
#define VKB_WARP_PERSPECTIVE 1

typedef struct vkb_warp_page {
    vkb_planes planes;
    double inv[9];
    int32_t kind;
    int32_t _pad;
} vkb_warp_page;

int vkb_warp_fused(const vkb_warp_page* pages, int32_t n_pages, int32_t max_dst_h,
                   int32_t max_dst_w, void* stream);

/* Points through a 2x3 / 3x3 forward matrix (affine_np_points, affine.py:46-64).
 * xy_in / xy_out: n x 2 float64 (x, y). mat: 6 or 9 doubles (host pointer, copied by value).
 * `f32_math` != 0 evaluates in float32 like the reference does for the 2x3 ops. */
int vkb_affine_points(const double* mat_host, int32_t rows, const double* xy_in, double* xy_out,
                      int32_t n, int32_t f32_math, void* stream);

/* ---------------------------------------------------------------------------------------
 * Grid-based ops: camera_* and similarity_mls.
 * Replaces create_dst_image_grid_and_shift_amounts_and_resize_ratios (grid_creator.py:44-115),
 * the point projectors (camera.py:188-212, 276-282, 375-398, 464-480; mls.py:52-135),
 * ImageGrid.get_inv_trans_mat / generate_remap_params (type.py:182-261) and
 * blend_src_to_dst_* (grid_blender.py:54-81).
 * ------------------------------------------------------------------------------------- */
#define VKB_PROJ_CAMERA 0
#define VKB_PROJ_MLS 1
#define VKB_PROJ_GIVEN 2 /* lattice_f already holds the projected lattice (tests, plugins) */

#define VKB_CAM_PLANE 0
#define VKB_CAM_CUBIC 1
#define VKB_CAM_LINE_FOLD 2
#define VKB_CAM_LINE_CURVE 3

typedef struct vkb_grid_page {
    int32_t src_h, src_w;
    int32_t grid_size;
    int32_t rows, cols; /* lattice points per column / row (grid_creator.py:22-41) */
    int32_t projector;
    int32_t strategy;
    int32_t resize_as_src;
    /* camera model (camera.py:58-196): Rodrigues matrix, translation, focal length */
    double R[9];
    double t[3];
    double focal;
    /* cubic curve (camera.py:313-398) */
    double poly[4];
    double curve_scale;
    float rot2[4];
    float proj_min, proj_range;
    /* plane line fold / curve (camera.py:432-480) */
    double line_c;
    double dist_max;
    double line_alpha;
    float line_ab[2];
    float perturb[3];
    int32_t n_handles;
    /* similarity MLS (mls.py:38-135): n_handles x 2 float32 (x, y) each */
    const float* handles_src;
    const float* handles_dst;
} vkb_grid_page;

typedef struct vkb_grid_meta {
    int32_t dst_h, dst_w;     /* result shape (type.py:41-50) */
    int32_t shift_y, shift_x; /* grid_creator.py:61-81 */
    double resize_ratio_y, resize_ratio_x;
    int32_t status; /* bit 0: some tile candidate list overflowed (slow path used) */
    int32_t n_flagged_cells;
} vkb_grid_meta;

/* Fixed per-cell budget of coverage-mask words; cells that need more are rasterised on the
 * fly by the remap kernel. */
#define VKB_CELL_MASK_WORDS 32
#define VKB_TILE 32     /* dst tile edge of the remap kernel */
#define VKB_TILE_CAP 64 /* candidate cells kept per tile before the slow path is used */

/* Phase 1: project the source lattice (p_max >= rows*cols of every page).
 * lattice_f: n_pages x p_max x 2 doubles (x, y), un-shifted projected coordinates.
 * projectors: bit (1 << VKB_PROJ_*) set for every projector some page uses (the pages live in
 * device memory; a projector nobody uses costs no launch). */
int vkb_grid_project(const vkb_grid_page* pages, int32_t n_pages, int32_t p_max,
                     double* lattice_f, int32_t projectors, void* stream);

/* Phase 1b: round (half to even), shift to the origin, optional resize_as_src, result shape.
 * lattice_i: n_pages x p_max x 2 int32 (x, y).
 * meta_mirror: NULL, or a second copy of `meta` in device-accessible (mapped, pinned) HOST
 * memory: the host needs the result shapes to allocate the outputs, and a mirror written by the
 * kernel is readable after a stream synchronise without a D2H copy (which would queue behind
 * bulk copies on the copy engine). */
int vkb_grid_finalize(const vkb_grid_page* pages, int32_t n_pages, int32_t p_max,
                      const double* lattice_f, int32_t* lattice_i, vkb_grid_meta* meta,
                      vkb_grid_meta* meta_mirror, void* stream);

/* Parameter blocks (vkb_grid_page / vkb_planes / ... arrays) from mapped pinned HOST memory to
 * device memory through a kernel instead of the copy engine, where a small copy would queue
 * behind the caller's bulk page copies.  nbytes: a multiple of 16; src_host: pinned host memory
 * (device accessible under unified addressing); stream ordered like every other call. */
int vkb_stage_params(void* dst, const void* src_host, int64_t nbytes, void* stream);

/* Phase 1c (batches, optional): the output layout on the device, so that no host round trip
 * sits between the projection and the remap.  `planes` (device, n_pages records) arrives with
 * the source fields filled and dst_image / dst_mask / dst_score = the BASE of the respective
 * output arena (NULL when absent); the call adds every page's offset (exclusive scan of
 * dst_h * dst_w over the pages, in pixels) and writes dst_h / dst_w from `meta`.
 *   layout:        n_pages + 2 int64: pixel offset per page, total pixels, status
 *   layout_mirror: NULL or the same in mapped pinned host memory (read after a synchronise)
 *   status:        0 ok; bit 0: the batch needs more than cap_pixels; bit 1: some page has more
 *                  than t_max 32 x 32 tiles.  On a non-zero status every result shape in `meta`
 *                  and `planes` is zeroed (vkb_grid_build / vkb_grid_remap then do nothing; the
 *                  true shapes stay in vkb_grid_finalize's meta_mirror) and the caller runs the
 *                  batch again with exact capacities. */
int vkb_grid_layout(vkb_grid_meta* meta, int32_t n_pages, vkb_planes* planes, int64_t cap_pixels,
                    int32_t t_max, int64_t* layout, int64_t* layout_mirror, void* stream);

/* Phase 2a: per cell inverse homography (dst -> src), bounding box, coverage mask and the
 * float32 record the remap evaluates; per dst tile the sorted list of candidate cells.
 * c_max >= (rows-1)*(cols-1); t_max >= 32x32 tiles per page; s_cap >= c_max + 2 * t_max records
 * per page (the cells' records, then the tiles' candidate lists).
 *   hinv:       n_pages x c_max x 9 doubles
 *   hfwd:       n_pages x c_max x 9 doubles or NULL (forward maps, needed by vkb_grid_points)
 *   cell_box:   n_pages x c_max x 4 int32 (x0, y0, x1, y1); bit 30 of x1 set = coverage of
 *               this cell exceeds the mask budget and is rasterised on the fly by the remap
 *   cell_masks: n_pages x c_max x VKB_CELL_MASK_WORDS uint32
 *   tile_count: n_pages x t_max int32 (zeroed by this call)
 *   tile_cells: n_pages x t_max x VKB_TILE_CAP uint16 (unordered bins)
 *   tile_off:   n_pages x t_max int32; unused since the records became per cell (kept so that
 *               the signature of the call did not change)
 *   tile_base:  n_pages + 1 int32, prefix sum of tiles per page (the remap's flat work list)
 *   tile_slots: n_pages x s_cap x VKB_TILE_SLOT_BYTES bytes, opaque: per page one record per
 *               cell (bbox, mask flag, float32 inverse map re-centred on the bbox origin), then
 *               per tile VKB_TILE_CAP uint16 with its candidate cells in ascending order
 *   tile_headers: n_pages x t_max x VKB_TILE_HEADER_BYTES bytes (the flat work list of the
 *               remap: page, tile origin, candidate count, acceptance limit of the float32
 *               coordinates, the first 16 sorted candidates) followed by
 *               (n_pages x t_max + 2) int32: the list of tiles that take the remap's second
 *               launch (count + entries: more than 15 candidates or the exact slow path) and, last,
 *               the work counter vkb_grid_remap resets and its small-tile kernel draws from; opaque */
#define VKB_TILE_SLOT_BYTES 64
#define VKB_TILE_HEADER_BYTES 64
int vkb_grid_build(const vkb_grid_page* pages, int32_t n_pages, int32_t p_max, int32_t c_max,
                   int32_t t_max, int32_t s_cap, const int32_t* lattice_i, vkb_grid_meta* meta,
                   double* hinv, double* hfwd, int32_t* cell_box, uint32_t* cell_masks,
                   int32_t* tile_count, uint16_t* tile_cells, int32_t* tile_off,
                   int32_t* tile_base, void* tile_slots, void* tile_headers, void* stream);

/* Phase 2b: the fused remap -- owner cell per dst pixel (last cell in row-major order whose
 * cv.fillPoly coverage contains it), per-pixel inverse homography in double, float32 map
 * value, 1/32 px quantisation, bilinear gather of Image + Mask + ScoreMap in one pass.
 * Persistent kernels: every warp draws 32 x 32 dst tiles from the flat tile list (a work counter
 * in the tile_headers workspace) and prefetches the next tile's records while it works on the
 * current one.  Concurrent calls need separate workspaces.
 * The kernel is specialised at compile time on the containers present, so every page of one
 * call carries the same set: image_channels (0 = no image), has_mask, has_score.
 * planes[i].dst_h / dst_w must be the result shape in meta[i]; src planes below 32768 px. */
int vkb_grid_remap(const vkb_grid_page* pages, const vkb_planes* planes, int32_t n_pages,
                   int32_t p_max, int32_t c_max, int32_t t_max, int32_t s_cap,
                   const int32_t* lattice_i, const double* hinv, const int32_t* cell_box,
                   const uint32_t* cell_masks, const int32_t* tile_count,
                   const int32_t* tile_off, const int32_t* tile_base, const void* tile_slots,
                   void* tile_headers, int32_t image_channels, int32_t has_mask,
                   int32_t has_score, void* stream);

/* Points through the forward homography of the source cell that contains the ROUNDED point
 * (FuncImageGridBased.func_point, grid_rendering/interface.py:194-216).
 * xy_in: n x 2 doubles (smooth x, y); cell_rc: n x 2 int32 (polygon_row, polygon_col). */
int vkb_grid_points(const double* hfwd_page, int32_t cols_minus_1, const double* xy_in,
                    const int32_t* cell_rc, double* xy_out, int32_t n, void* stream);

/* The same for the points of MANY pages in one launch (RandomDistortionBatch): page_cell: n x 3
 * int32 (page, polygon_row, polygon_col); hfwd: the n_pages x c_max x 9 array of vkb_grid_build. */
int vkb_grid_points_batched(const vkb_grid_page* pages, const double* hfwd, int32_t c_max,
                            const double* xy_in, const int32_t* page_cell, double* xy_out,
                            int32_t n, void* stream);

/* affine_np_points (affine.py:46-64) for the points of many pages: mats: n_pages x 9 doubles
 * (forward matrices, row major); rows_f32[page]: 2 or 3 rows, | 4 = float32 arithmetic (the 2 x 3
 * ops); page_of: the page of every point.  All device pointers. */
int vkb_affine_points_batched(const double* mats, const int32_t* rows_f32, const int32_t* page_of,
                              const double* xy_in, double* xy_out, int32_t n, void* stream);

/* Active mask: cv.fillPoly of one polygon (the dst lattice border, interface.py:177-192)
 * on a zeroed uint8 canvas.  poly_xy: n_pts x 2 int32 (x, y). */
int vkb_fill_polygon(uint8_t* mask, int32_t h, int32_t w, const int32_t* poly_xy, int32_t n_pts,
                     uint8_t value, void* stream);

/* Ordered polygon fills -- label rasterisation after distortion (text-line / char masks and
 * height score maps: pipeline/text_detection/page_distortion.py:163-314, Polygon.fill_mask /
 * fill_score_map element/polygon.py:458-487, engine/char_mask/default.py:44-56).
 * Every polygon is rasterised with cv.fillPoly semantics and applied in list order:
 *   mode 0: assign (later polygons overwrite earlier ones)   mode 1: keep max   mode 2: keep min
 * dst: h x w uint8 (dst_f32 = 0) or float32 plane, updated in place; pts_xy: all vertices
 * (x, y) int32, device; items: device array (items_host: the same records on the host);
 * keys: h x w int32 device workspace. */
typedef struct vkb_poly_item {
    int32_t first_pt, n_pts; /* slice of pts_xy */
    int32_t y_min, y_max;    /* vertical extent of the polygon */
    float value;
    int32_t pad_;
} vkb_poly_item;
int vkb_fill_polygons(void* dst, int32_t dst_f32, int32_t h, int32_t w, const int32_t* pts_xy,
                      const vkb_poly_item* items, const vkb_poly_item* items_host,
                      int32_t n_items, int32_t mode, int32_t* keys, void* stream);

/* ---------------------------------------------------------------------------------------
 * Blend: the device form of fill_np_array (vkit/element/opt.py:118-209) as reached through
 * Box / Mask / ScoreMap / Polygon .fill_image / .fill_mask / .fill_score_map
 * (box.py:311-416, score_map.py:678-687) -- glyph -> text line -> page compositing
 * (engine/font/freetype.py:314-380, pipeline/text_detection/page_assembler.py:152-236,
 * page_distortion.py:146-161).
 *   active pixel : mask != 0, else alpha_arr > 0 when alpha is an array, else every pixel
 *   alpha array  : out = trunc((1 - a) * f32(dst) + a * f32(value))   (float32, no FMA)
 *   alpha scalar : 1 -> assign (or keep max / keep min), (0,1) -> same blend with a = f32(alpha)
 * ------------------------------------------------------------------------------------- */
#define VKB_BLEND_FLOAT_CONST 4
typedef struct vkb_blend_item {
    void* dst;             /* full destination plane, HWC uint8 or HW float32 */
    const void* value_arr; /* NULL: value_const; else same dtype as dst, origin at region */
    const uint8_t* mask;   /* NULL or region shaped */
    const float* alpha_arr; /* NULL or region shaped */
    int32_t dst_f32;
    int32_t channels;
    int32_t dst_w; /* pixels per destination row */
    int32_t box_y, box_x, box_h, box_w;
    int32_t value_pitch; /* pixels per row of value_arr */
    int32_t mask_pitch;
    int32_t alpha_pitch;
    int32_t keep_mode; /* 0 none, 1 keep max, 2 keep min; | VKB_BLEND_FLOAT_CONST: an alpha blend into a
                          uint8 plane uses value_const as float32 instead of casting it to uint8
                          first (fog on a GRAYSCALE page, effect.py:194-197) */
    float alpha;
    float value_const[4];
} vkb_blend_item;

int vkb_blend_fill(const vkb_blend_item* item_host, void* stream);
/* Ordered draw list: items (device array) are applied in order; overlapping items see each
 * other's result exactly as successive fill_image calls would. All items share one dst. */
int vkb_blend_draw_list(const vkb_blend_item* items, int32_t n_items, int32_t dst_h,
                        int32_t dst_w, void* stream);

/* ---------------------------------------------------------------------------------------
 * Colour-mode conversion: Image.to_target_mode_image (vkit/element/image.py:771-814).
 * uint8 in/out; HSL is stored H,S,L (the reference slices cv2's HLS with [0,2,1]).
 * ------------------------------------------------------------------------------------- */
#define VKB_CVT_RGB2HSV 0
#define VKB_CVT_HSV2RGB 1
#define VKB_CVT_RGB2HSL 2
#define VKB_CVT_HSL2RGB 3
#define VKB_CVT_RGB2GRAY 4
#define VKB_CVT_GRAY2RGB 5
#define VKB_CVT_RGBA2RGB 6
#define VKB_CVT_RGB2RGBA 7
#define VKB_CVT_GRAY2RGBA 8
#define VKB_CVT_RGBA2GRAY 9
int vkb_cvt_color(const uint8_t* src, uint8_t* dst, int64_t n_pixels, int32_t code, void* stream);

/* ---------------------------------------------------------------------------------------
 * Per-pixel photometric ops, fused: an op list interpreted per pixel in one pass
 * (vkit/mechanism/distortion/photometric/color.py:32-439, opt.py:24-85).
 * ------------------------------------------------------------------------------------- */
#define VKB_OP_MEAN_SHIFT 0       /* i0 delta, i1 threshold (-1 none), i2 channel bits, i3 cycle */
#define VKB_OP_HUE_SHIFT_RGB 1    /* i0 delta: RGB -> HSV_FULL, H += delta (mod 256), -> RGB */
#define VKB_OP_LIGHT_SHIFT_RGB 2  /* i0 delta, i1 0: via HSL (L), 1: via HSV (V); clipped */
#define VKB_OP_STD_SHIFT 3        /* f0 scale, f1..f3 mean*(scale-1) per channel, i2 channel bits */
#define VKB_OP_COMPLEMENT 4       /* i1 threshold (-1 none), i3 lte flag, i2 channel bits */
#define VKB_OP_POSTERIZE 5        /* i0 bit mask, i2 channel bits */
#define VKB_OP_COLOR_BALANCE 6    /* f0 ratio */
#define VKB_OP_PERMUTE 7          /* i0 packed permutation (4 bits per channel) */
#define VKB_OP_BOUNDARY_EQ 8      /* f0..f2 min per channel, g0..g2 scale per channel, i2 bits */
/* position dependent ops: LINE_STREAK only through vkb_photo_chain_batched, NOISE only through
 * vkb_noise_philox_batched */
#define VKB_OP_NOISE 9            /* i0 kind (vkb_noise_philox), i1 / i2 seed low / high word, f0 p0, f1 p1 */
#define VKB_OP_LINE_STREAK 10     /* i0 thickness, i1 gap, i2 dash_thickness | dash_gap << 16, i3 bit 0 vert / bit 1 hori, f0 alpha, f1 f2 f3 g0 colour */
typedef struct vkb_color_op {
    int32_t kind;
    int32_t i0, i1, i2, i3;
    float f0, f1, f2, f3;
    float g0, g1, g2;
} vkb_color_op;
#define VKB_MAX_COLOR_OPS 8
int vkb_color_ops(const uint8_t* src, uint8_t* dst, int64_t n_pixels, int32_t channels,
                  const vkb_color_op* ops_host, int32_t n_ops, void* stream);

/* Per-channel sum / min / max of a uint8 image (std_shift mean, boundary_equalization).
 * out: 3 x uint64 sums, then 3 x uint32 mins, 3 x uint32 maxs (device, 48 bytes). */
int vkb_channel_stats(const uint8_t* src, int64_t n_pixels, int32_t channels, void* out,
                      void* stream);

/* cv.equalizeHist building blocks (photometric/color.py:264-298): per-channel 256-bin
 * histogram (out: 3 x 256 uint32, device) and a per-channel 256-entry LUT (3 x 256 uint8,
 * device) applied to the channels selected by channel_bits. */
int vkb_histogram_u8(const uint8_t* src, int64_t n_pixels, int32_t channels, uint32_t* out,
                     void* stream);
int vkb_apply_lut(const uint8_t* src, uint8_t* dst, int64_t n_pixels, int32_t channels,
                  const uint8_t* lut, int32_t channel_bits, void* stream);

/* cv.GaussianBlur(uint8, (k,k), sigma) with BORDER_REFLECT_101: separable 8.8 fixed-point
 * stencil staged through shared memory (photometric/blur.py:49-76). kernel_host: k integer
 * taps summing to 256 (k odd, k <= 17). */
int vkb_gaussian_blur_u8(const uint8_t* src, uint8_t* dst, int32_t h, int32_t w, int32_t channels,
                         const int32_t* kernel_host, int32_t ksize, void* stream);

/* cv.filter2D(uint8, -1, float32 kernel) with BORDER_REFLECT_101 and the anchor at the kernel
 * centre: float32 accumulation in row-major tap order, round half to even, saturate
 * (defocus_blur / motion_blur, photometric/blur.py:79-192).  taps_dev: kh x kw float32 on the
 * device, kh and kw odd and <= 63.  src != dst. */
int vkb_filter2d_u8(const uint8_t* src, uint8_t* dst, int32_t h, int32_t w, int32_t channels,
                    const float* taps_dev, int32_t kh, int32_t kw, void* stream);

/* cv.resize(uint8): INTER_NEAREST, INTER_LINEAR (11-bit fixed-point coefficients and cv2's 8-bit
 * vertical pass), INTER_AREA (box sums / computeResizeAreaTab weights when shrinking on both axes,
 * the bilinear passes with "area mode" fractions as soon as one axis enlarges),
 * INTER_LANCZOS4, INTER_NEAREST_EXACT, INTER_LINEAR_EXACT -- all bit exact -- and INTER_CUBIC
 * (pixelation, photometric/effect.py:58-79; Image.to_resized_image element/image.py:836-852;
 * Mask.to_resized_mask element/mask.py:454-479; every interpolation page_resizing samples,
 * utility/opt.py:125-148, pipeline/text_detection/page_resizing.py:118-121). */
#define VKB_INTER_NEAREST 0
#define VKB_INTER_LINEAR 1
#define VKB_INTER_CUBIC 2 /* the wheel's IPP cubic (float64 bicubic; +-1 on < 3e-4 of the pixels); sources below 4x4: cv2's own path, bit exact */
#define VKB_INTER_AREA 3     /* bit exact (uint8) */
#define VKB_INTER_LANCZOS4 4 /* bit exact (uint8) */
#define VKB_INTER_LINEAR_EXACT 5  /* cv2's codes: bit exact */
#define VKB_INTER_NEAREST_EXACT 6
int vkb_resize_u8(const uint8_t* src, int32_t src_h, int32_t src_w, uint8_t* dst, int32_t dst_h,
                  int32_t dst_w, int32_t channels, int32_t interpolation, void* stream);

/* cv.resize(float32, one channel): ScoreMap.to_resized_score_map (element/score_map.py:616-637).
 * cv2's float paths (NEAREST, LINEAR, CUBIC, AREA, LANCZOS4, the two EXACT codes)
 * LINEAR / LINEAR_EXACT (sources of at least 2 x 2) and CUBIC (4 x 4) follow the wheel's default
 * backend, Intel IPP (coordinates and taps in double; restated in float64, within 5e-7 of the
 * wheel); the other codes and smaller sources restate cv2's own path in float32, products and sums
 * in cv2's order: bit identical to cv2.  clip01 != 0
 * fuses the np.clip(mat, 0, 1) the reference applies to probability maps. */
int vkb_resize_f32(const float* src, int32_t src_h, int32_t src_w, float* dst, int32_t dst_h,
                   int32_t dst_w, int32_t interpolation, int32_t clip01, void* stream);
/* The same with every resized value multiplied by `post_scale` (one float32 product, after the
 * clip) before it is stored: PageResizingStep scales its height score maps by the resize ratio
 * right after resizing them (pipeline/text_detection/page_resizing.py:160-161, 177-180). */
int vkb_resize_f32_scaled(const float* src, int32_t src_h, int32_t src_w, float* dst, int32_t dst_h,
                          int32_t dst_w, int32_t interpolation, int32_t clip01, float post_scale,
                          void* stream);

/* Mask.to_resized_mask (element/mask.py:454-479) in ONE pass: the source reads as (v > 0) * 255,
 * cv.resize with `interpolation` (same codes and exactness as vkb_resize_u8), the result is stored
 * as (resized > binarization_threshold) in {0, 1}. */
int vkb_resize_mask_u8(const uint8_t* src, int32_t src_h, int32_t src_w, uint8_t* dst, int32_t dst_h,
                       int32_t dst_w, int32_t interpolation, int32_t binarization_threshold,
                       void* stream);

/* dst = src > threshold ? high : low -- the two binarisations of Mask.to_resized_mask
 * (element/mask.py:454-479: mask * 255 before cv.resize, > threshold after).  In place allowed. */
int vkb_threshold_u8(const uint8_t* src, uint8_t* dst, int64_t n, int32_t threshold, int32_t low,
                     int32_t high, void* stream);

/* zoom_in_blur (photometric/blur.py:278-330): the page averaged with n_levels cubic enlargements
 * of itself (centre crops), blended with the page by alpha.  levels: device array; the level's
 * enlargement has (src dims / scale) pixels and the crop starts at (up, left). */
typedef struct vkb_zoom_level {
    double scale_x, scale_y; /* 1 / (resized / original), as cv::resize computes it */
    int32_t up, left;
} vkb_zoom_level;
int vkb_zoom_in_blur_u8(const uint8_t* src, uint8_t* dst, int32_t h, int32_t w, int32_t channels,
                        const vkb_zoom_level* levels_dev, int32_t n_levels, double alpha,
                        void* stream);

/* The alpha field of `fog` (photometric/effect.py:89-208: generate_diamond_square_mask and the
 * normalisation of fog_image) computed on the device FROM THE CALLER'S NumPy GENERATOR STREAM:
 * (state, inc) is the PCG64 state after the four scalar corner draws; the array draws of the
 * levels are regenerated on the device (draw i = one 64-bit output, reached by jump-ahead), the
 * host advances its generator by the `count` of vkb_fog_draws(size, &count) outputs and then draws the crop offsets.
 * weight[l] = roughness ** l as the host's Python float computes it.  Workspaces: field
 * size x size float32, centres ((size - 1) / 2)^2 doubles, draws `count` doubles,
 * minmax 2 x uint32.  alpha: height x width float32 = the blend weights of the fog colour. */
#define VKB_FOG_MAX_LEVELS 16
typedef struct vkb_fog_params {
    uint64_t state_hi, state_lo, inc_hi, inc_lo;
    double weight[VKB_FOG_MAX_LEVELS];
    float corners[4]; /* field[0][0], field[0][-1], field[-1][-1], field[-1][0] */
    int32_t size;     /* 2^k + 1 */
    int32_t up, left, height, width;
    float ratio_span, ratio_min; /* float32(ratio_max - ratio_min), float32(ratio_min) */
    int32_t _pad;
} vkb_fog_params;
int vkb_fog_draws(int32_t size, int64_t* count);
int vkb_fog_mask(const vkb_fog_params* p, float* field, double* centres, double* draws,
                 uint32_t* minmax, float* alpha, void* stream);

/* The pixel permutation of glass_blur (photometric/blur.py:232-262) built on the device from the
 * caller's NumPy PCG64 stream.  vkb_glass_init: pos_y / pos_x = identity maps (h x w int32),
 * owner = -1 (h x w int32 scratch), flag = 0.  vkb_glass_round: one round -- the host has drawn the
 * round's two offsets (row0, col0) and passes the generator's state after them ((state, inc) and
 * the pending 32-bit half, if any); the kernel regenerates the round's 2 * n bounded integers
 * (n = centres of the round; Lemire's method on the 32-bit halves, as NumPy does), swaps every
 * centre with its target (duplicate targets: the last centre in C order wins, like NumPy's fancy
 * assignment) and sets *flag when NumPy would have REJECTED a draw (2^-32 per draw): the maps are
 * then invalid and the caller repeats the call on the host.  target: n int32, at_centre /
 * at_target: 2 * n int32 scratch (n <= ceil(h / period) * ceil(w / period)). */
int vkb_glass_init(int32_t* pos_y, int32_t* pos_x, int32_t* owner, int32_t h, int32_t w,
                   int32_t* flag, void* stream);
int vkb_glass_round(int32_t* pos_y, int32_t* pos_x, int32_t* owner, int32_t h, int32_t w,
                    int32_t row0, int32_t col0, int32_t delta, int32_t round, uint64_t state_hi,
                    uint64_t state_lo, uint64_t inc_hi, uint64_t inc_lo, int32_t has_cached,
                    uint32_t cached, int32_t* target, int32_t* at_centre, int32_t* at_target,
                    int32_t* flag, void* stream);

/* dst[y, x] = src[pos_y[y, x], pos_x[y, x]] (uint8 HWC): the pixel permutation of glass_blur
 * (photometric/blur.py:216-264); the index maps are the host-drawn random field. */
int vkb_gather_pixels_u8(const uint8_t* src, uint8_t* dst, int32_t h, int32_t w, int32_t channels,
                         const int32_t* pos_y, const int32_t* pos_x, void* stream);

/* Noise (photometric/noise.py:25-190).  kind: 0 gaussian (p0 = std), 1 poisson, 2 impulse
 * (p0 = prob_salt, p1 = prob_pepper), 3 speckle (p0 = std).
 * Philox variant: counter-based device RNG keyed by (seed, element index); distributional
 * parity.  Field variant: the host supplies the random field the reference would draw
 * (gaussian: int16 rounded noise; poisson: int64 samples; impulse: int64 categories 0/1/2 per
 * PIXEL; speckle: float64 normal) and the result is bit-exact. */
int vkb_noise_philox(const uint8_t* src, uint8_t* dst, int64_t n_pixels, int32_t channels,
                     int32_t kind, double p0, double p1, uint64_t seed, void* stream);
int vkb_noise_field(const uint8_t* src, uint8_t* dst, int64_t n_pixels, int32_t channels,
                    int32_t kind, const void* field, void* stream);

/* Streaks (photometric/streak.py:44-337).  line: analytic periodic masks; rect: the masks are
 * rasterised from the host-computed bar lists; both blend `color` with `alpha`, vertical mask
 * first, horizontal second (crossings get alpha twice, streak.py:96-98). In place on `image`. */
/* jpeg_quality (vkit/mechanism/distortion/photometric/effect.py:26-55): the pixels of
 * cv.imdecode(cv.imencode('.jpeg', mat, [IMWRITE_JPEG_QUALITY, quality])) without the (lossless)
 * entropy coding: libjpeg's RGB -> YCbCr, 4:2:0 downsampling, islow DCT, quantisation with the
 * quality-scaled Annex K tables, dequantisation, islow IDCT, fancy upsampling, YCbCr -> RGB, all in
 * libjpeg's integer arithmetic -- bit exact vs cv2 4.13 / libjpeg-turbo.  src / dst: dense h x w x
 * channels uint8 (channels 3: channel 0 plays blue like in cv2; channels 1: GRAYSCALE);
 * planes: device workspace, 16-byte aligned, 1.5 bytes per pixel of the page padded to 16 x 16. */
int vkb_jpeg_round_trip_u8(const uint8_t* src, uint8_t* dst, int32_t h, int32_t w, int32_t channels,
                           int32_t quality, uint8_t* planes, int64_t planes_bytes, void* stream);

/* ellipse_streak (vkit/mechanism/distortion/photometric/streak.py:282-337): cv.ellipse(mask,
 * center, axes, 0, 0, 360, 1, thickness) for a list of ellipses, ORed into a uint8 h x w mask
 * (caller zeroes it).  ellipses_host: n x 4 int32 (center x, center y, half axis x, half axis y)
 * on the HOST: the vertex lists of OpenCV's EllipseEx / ellipse2Poly are prepared there, the
 * pixels (Bresenham thin lines; for thickness > 1 convex quads with Line2 outlines + round caps,
 * all bit exact vs cv2 4.13) are drawn on the device, one thread per polyline segment.
 * seg_workspace: device scratch of at least 74 * 40 bytes per ellipse. */
int vkb_draw_ellipses(uint8_t* mask, int32_t h, int32_t w, const int32_t* ellipses_host,
                      int32_t n_ellipses, int32_t thickness, void* seg_workspace,
                      int64_t workspace_bytes, void* stream);

typedef struct vkb_rect { int32_t up, down, left, right; } vkb_rect;
int vkb_streak_line(uint8_t* image, int32_t h, int32_t w, int32_t channels, int32_t thickness,
                    int32_t gap, int32_t dash_thickness, int32_t dash_gap, int32_t enable_vert,
                    int32_t enable_hori, const float* color_host, float alpha, void* stream);
int vkb_fill_rects(uint8_t* mask, int32_t h, int32_t w, const vkb_rect* rects, int32_t n_rects,
                   void* stream);
int vkb_streak_masks(uint8_t* image, int32_t h, int32_t w, int32_t channels,
                     const uint8_t* mask_vert, const uint8_t* mask_hori, int32_t dash_thickness,
                     int32_t dash_gap, const float* color_host, float alpha, void* stream);

/* Batched photometric chain over a ragged batch of pages: optional cv.GaussianBlur (uint8,
 * 8.8 fixed point, BORDER_REFLECT_101; photometric/blur.py:49-76) followed by a per-pixel op
 * list (photometric/color.py:32-439) in ONE pass, per-page shapes / taps / ops.  The batched
 * form of vkb_gaussian_blur_u8 + vkb_color_ops for chains such as
 * similarity_mls -> gaussian_blur -> color_shift (random_distortion.py:350-392).
 * `pages`: device array, `pages_host`: the same records on the host (validation and launch
 * shape).  src != dst when blur_radius > 0.  blur_taps: 2*radius+1 integers summing to 256. */
typedef struct vkb_photo_page {
    const uint8_t* src;
    uint8_t* dst;
    int32_t h, w;
    int32_t blur_radius; /* 0: no blur */
    int32_t n_ops;
    int32_t blur_taps[17];
    int32_t pad_;
    vkb_color_op ops[VKB_MAX_COLOR_OPS];
} vkb_photo_page;
int vkb_photo_chain_batched(const vkb_photo_page* pages, const vkb_photo_page* pages_host,
                            int32_t n_pages, int32_t channels, void* stream);
/* Philox noise over a ragged batch (the batched form of vkb_noise_philox): ops[0] of every page
 * is a VKB_OP_NOISE record (pages without one are skipped); keyed by (seed, pixel index of the
 * page), so a page gives the same result alone or in a batch.  In place allowed. */
int vkb_noise_philox_batched(const vkb_photo_page* pages, int32_t n_pages, int32_t channels,
                             int32_t blocks_per_page, void* stream);
/* Per-page channel statistics of the `src` planes of a ragged batch.
 * out (device): per page 3 x uint64 sums, then 3 x uint32 mins, 3 x uint32 maxs (48 bytes). */
int vkb_channel_stats_batched(const vkb_photo_page* pages, int32_t n_pages, int32_t channels,
                              void* out, void* stream);

/* ---------------------------------------------------------------------------------------
 * The step before compositing.
 *
 * Background synthesis -- ImageCombinerEngine.synthesize_image
 * (vkit/engine/image/combiner.py:178-333): texture segments pasted into an h x w canvas in list
 * order (later segments overwrite earlier ones, uncovered pixels stay 0), then the pixels inside
 * the edge bands of every segment (fill_np_edge_mask, combiner.py:147-176; `band` =
 * gaussian_blur_kernel_size // 2 + 1) are replaced by cv.GaussianBlur(canvas, (ksize, ksize),
 * sigma) with BORDER_REFLECT_101 -- `kernel_host`: the ksize 8.8 fixed-point taps (sum 256) cv2
 * derives for uint8.  One pass, one launch; the canvas is written once.  items: device array;
 * `src` of an item points at the texture pixel that lands on (up, left), src_pitch in pixels.
 * ------------------------------------------------------------------------------------- */
typedef struct vkb_paste_item {
    const uint8_t* src;
    int32_t src_pitch;
    int32_t up, down, left, right; /* inclusive canvas rectangle */
    int32_t pad_;
} vkb_paste_item;
int vkb_background_compose(uint8_t* dst, int32_t h, int32_t w, int32_t channels,
                           const vkb_paste_item* items, int32_t n_items, int32_t band,
                           const int32_t* kernel_host, int32_t ksize, void* stream);

/* Glyph atlas -- the planes render_char_glyphs_in_text_line blends from
 * (vkit/engine/font/freetype.py:136-221 build_char_glyph, :314-380 the renderer,
 * engine/font/type.py:423-451 get_glyph_mask), derived on the device from FreeType coverage
 * bitmaps uploaded once:  mask = bitmap > 0 (any channel for H x W x 3 LCD bitmaps);
 * alpha = alpha_lut[bitmap] (float32; the caller evaluates np.power(v / 255, gamma) for the 256
 * byte values, so the numbers are NumPy's own); lcd_image = lcd_lut[bitmap] per channel.
 * One block per glyph.  items: device array, items_host: the same records for validation. */
typedef struct vkb_glyph_item {
    const uint8_t* bitmap; /* n_pixels x channels */
    uint8_t* mask;         /* n_pixels */
    float* alpha;          /* n_pixels or NULL (LCD glyphs have no score map) */
    uint8_t* lcd_image;    /* n_pixels x 3 or NULL */
    const float* alpha_lut; /* 256 float32 */
    const uint8_t* lcd_lut; /* 256 uint8 */
    int32_t n_pixels;
    int32_t channels; /* 1 or 3 */
} vkb_glyph_item;
int vkb_glyph_prepare(const vkb_glyph_item* items, const vkb_glyph_item* items_host,
                      int32_t n_items, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VKIT_B200_H_ */
