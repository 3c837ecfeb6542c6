"""ctypes binding of libvkit_b200.so (the C ABI declared in include/vkit_b200.h).

There is no CPU fallback: if the shared library is missing or CUDA is unavailable every op
raises.  Struct layouts are mirrored both as `ctypes.Structure` (single records) and as NumPy
structured dtypes (batched parameter blocks that are uploaded in one copy).
"""
import ctypes
import os
from ctypes import POINTER, c_double, c_float, c_int32, c_uint8, c_uint16, c_uint32, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# VKB_LIB: alternative build of the same ABI (kernel experiments); the default is the in-tree build
LIB_PATH = os.environ.get('VKB_LIB') or os.path.join(_HERE, 'csrc', 'libvkit_b200.so')

CELL_MASK_WORDS = 32
TILE = 32
TILE_CAP = 64
TILE_SLOT_BYTES = 64
TILE_HEADER_BYTES = 64

BLEND_FLOAT_CONST = 4
WARP_AFFINE = 0
WARP_PERSPECTIVE = 1
PROJ_CAMERA = 0
PROJ_MLS = 1
PROJ_GIVEN = 2
CAM_PLANE = 0
CAM_CUBIC = 1
CAM_LINE_FOLD = 2
CAM_LINE_CURVE = 3


class Planes(ctypes.Structure):
    _fields_ = [
        ('src_image', c_void_p), ('dst_image', c_void_p),
        ('src_mask', c_void_p), ('dst_mask', c_void_p),
        ('src_score', c_void_p), ('dst_score', c_void_p),
        ('image_channels', c_int32),
        ('src_h', c_int32), ('src_w', c_int32),
        ('dst_h', c_int32), ('dst_w', c_int32),
        ('_pad', c_int32),
    ]


class WarpPage(ctypes.Structure):
    _fields_ = [('planes', Planes), ('inv', c_double * 9), ('kind', c_int32), ('_pad', c_int32)]


class GridPage(ctypes.Structure):
    _fields_ = [
        ('src_h', c_int32), ('src_w', c_int32), ('grid_size', c_int32),
        ('rows', c_int32), ('cols', c_int32),
        ('projector', c_int32), ('strategy', c_int32), ('resize_as_src', c_int32),
        ('R', c_double * 9), ('t', c_double * 3), ('focal', c_double),
        ('poly', c_double * 4), ('curve_scale', c_double),
        ('rot2', c_float * 4), ('proj_min', c_float), ('proj_range', c_float),
        ('line_c', c_double), ('dist_max', c_double), ('line_alpha', c_double),
        ('line_ab', c_float * 2), ('perturb', c_float * 3),
        ('n_handles', c_int32),
        ('handles_src', c_void_p), ('handles_dst', c_void_p),
    ]


class GridMeta(ctypes.Structure):
    _fields_ = [
        ('dst_h', c_int32), ('dst_w', c_int32), ('shift_y', c_int32), ('shift_x', c_int32),
        ('resize_ratio_y', c_double), ('resize_ratio_x', c_double),
        ('status', c_int32), ('n_flagged_cells', c_int32),
    ]


class BlendItem(ctypes.Structure):
    _fields_ = [
        ('dst', c_void_p), ('value_arr', c_void_p), ('mask', c_void_p), ('alpha_arr', c_void_p),
        ('dst_f32', c_int32), ('channels', c_int32), ('dst_w', c_int32),
        ('box_y', c_int32), ('box_x', c_int32), ('box_h', c_int32), ('box_w', c_int32),
        ('value_pitch', c_int32), ('mask_pitch', c_int32), ('alpha_pitch', c_int32),
        ('keep_mode', c_int32), ('alpha', c_float), ('value_const', c_float * 4),
    ]


class FogParams(ctypes.Structure):
    """vkb_fog_params (include/vkit_b200.h)."""
    _fields_ = [
        ('state_hi', ctypes.c_uint64), ('state_lo', ctypes.c_uint64),
        ('inc_hi', ctypes.c_uint64), ('inc_lo', ctypes.c_uint64),
        ('weight', ctypes.c_double * 16), ('corners', c_float * 4), ('size', c_int32),
        ('up', c_int32), ('left', c_int32), ('height', c_int32), ('width', c_int32),
        ('ratio_span', c_float), ('ratio_min', c_float), ('_pad', c_int32),
    ]


class ColorOp(ctypes.Structure):
    _fields_ = [
        ('kind', c_int32), ('i0', c_int32), ('i1', c_int32), ('i2', c_int32), ('i3', c_int32),
        ('f0', c_float), ('f1', c_float), ('f2', c_float), ('f3', c_float),
        ('g0', c_float), ('g1', c_float), ('g2', c_float),
    ]


class PhotoPage(ctypes.Structure):
    _fields_ = [
        ('src', c_void_p), ('dst', c_void_p), ('h', c_int32), ('w', c_int32),
        ('blur_radius', c_int32), ('n_ops', c_int32), ('blur_taps', c_int32 * 17),
        ('pad_', c_int32), ('ops', ColorOp * 8),
    ]


class PolyItem(ctypes.Structure):
    _fields_ = [('first_pt', c_int32), ('n_pts', c_int32), ('y_min', c_int32), ('y_max', c_int32),
                ('value', c_float), ('pad_', c_int32)]


class ZoomLevel(ctypes.Structure):
    _fields_ = [('scale_x', c_double), ('scale_y', c_double), ('up', c_int32), ('left', c_int32)]


class PasteItem(ctypes.Structure):
    _fields_ = [('src', c_void_p), ('src_pitch', c_int32), ('up', c_int32), ('down', c_int32),
                ('left', c_int32), ('right', c_int32), ('pad_', c_int32)]


class GlyphItem(ctypes.Structure):
    _fields_ = [('bitmap', c_void_p), ('mask', c_void_p), ('alpha', c_void_p),
                ('lcd_image', c_void_p), ('alpha_lut', c_void_p), ('lcd_lut', c_void_p),
                ('n_pixels', c_int32), ('channels', c_int32)]


class Rect(ctypes.Structure):
    _fields_ = [('up', c_int32), ('down', c_int32), ('left', c_int32), ('right', c_int32)]


(CVT_RGB2HSV, CVT_HSV2RGB, CVT_RGB2HSL, CVT_HSL2RGB, CVT_RGB2GRAY, CVT_GRAY2RGB, CVT_RGBA2RGB,
 CVT_RGB2RGBA, CVT_GRAY2RGBA, CVT_RGBA2GRAY) = range(10)
(OP_MEAN_SHIFT, OP_HUE_SHIFT_RGB, OP_LIGHT_SHIFT_RGB, OP_STD_SHIFT, OP_COMPLEMENT, OP_POSTERIZE,
 OP_COLOR_BALANCE, OP_PERMUTE, OP_BOUNDARY_EQ, OP_NOISE, OP_LINE_STREAK) = range(11)
MAX_COLOR_OPS = 8
INTER_NEAREST, INTER_LINEAR, INTER_CUBIC = 0, 1, 2
INTER_AREA, INTER_LANCZOS4, INTER_LINEAR_EXACT, INTER_NEAREST_EXACT = 3, 4, 5, 6
NOISE_GAUSSIAN, NOISE_POISSON, NOISE_IMPULSE, NOISE_SPECKLE = range(4)


def _np_dtype(struct_cls):
    """NumPy structured dtype with the exact layout (offsets, padding) of a ctypes struct."""
    names, formats, offsets = [], [], []
    for name, ctype in struct_cls._fields_:
        field = getattr(struct_cls, name)
        names.append(name)
        offsets.append(field.offset)
        if isinstance(ctype, type) and issubclass(ctype, ctypes.Structure):
            formats.append(_np_dtype(ctype))
        elif isinstance(ctype, type) and issubclass(ctype, ctypes.Array):
            formats.append((np.dtype(ctype._type_), (ctype._length_,)))
        elif ctype is c_void_p:
            formats.append(np.uint64)
        else:
            formats.append(np.dtype(ctype))
    return np.dtype({'names': names, 'formats': formats, 'offsets': offsets,
                     'itemsize': ctypes.sizeof(struct_cls)})


PLANES_DTYPE = _np_dtype(Planes)
BLEND_ITEM_DTYPE = _np_dtype(BlendItem)
RECT_DTYPE = _np_dtype(Rect)
ZOOM_LEVEL_DTYPE = _np_dtype(ZoomLevel)
POLY_ITEM_DTYPE = _np_dtype(PolyItem)
PHOTO_PAGE_DTYPE = _np_dtype(PhotoPage)
COLOR_OP_DTYPE = _np_dtype(ColorOp)
WARP_PAGE_DTYPE = _np_dtype(WarpPage)
GRID_PAGE_DTYPE = _np_dtype(GridPage)
GRID_META_DTYPE = _np_dtype(GridMeta)
PASTE_ITEM_DTYPE = _np_dtype(PasteItem)
GLYPH_ITEM_DTYPE = _np_dtype(GlyphItem)


class NativeError(RuntimeError):
    pass


_lib = None


def _declare(lib):
    lib.vkb_version.restype = c_int32
    lib.vkb_last_error.restype = ctypes.c_char_p
    i32 = c_int32
    vp = c_void_p
    lib.vkb_warp_fused.argtypes = [vp, i32, i32, i32, vp]
    lib.vkb_affine_points.argtypes = [POINTER(c_double), i32, vp, vp, i32, i32, vp]
    lib.vkb_grid_project.argtypes = [vp, i32, i32, vp, i32, vp]
    lib.vkb_grid_finalize.argtypes = [vp, i32, i32, vp, vp, vp, vp, vp]
    lib.vkb_grid_build.argtypes = [vp, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp,
                                   vp, vp, vp, vp]
    lib.vkb_grid_remap.argtypes = [vp, vp, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp,
                                   vp, i32, i32, i32, vp]
    lib.vkb_grid_points.argtypes = [vp, i32, vp, vp, vp, i32, vp]
    lib.vkb_grid_points_batched.argtypes = [vp, vp, i32, vp, vp, vp, i32, vp]
    lib.vkb_affine_points_batched.argtypes = [vp, vp, vp, vp, vp, i32, vp]
    lib.vkb_grid_layout.argtypes = [vp, i32, vp, ctypes.c_int64, i32, vp, vp, vp]
    lib.vkb_stage_params.argtypes = [vp, vp, ctypes.c_int64, vp]
    lib.vkb_fill_polygon.argtypes = [vp, i32, i32, vp, i32, c_uint8, vp]
    i64 = ctypes.c_int64
    lib.vkb_blend_fill.argtypes = [POINTER(BlendItem), vp]
    lib.vkb_blend_draw_list.argtypes = [vp, i32, i32, i32, vp]
    lib.vkb_cvt_color.argtypes = [vp, vp, i64, i32, vp]
    lib.vkb_color_ops.argtypes = [vp, vp, i64, i32, POINTER(ColorOp), i32, vp]
    lib.vkb_channel_stats.argtypes = [vp, i64, i32, vp, vp]
    lib.vkb_histogram_u8.argtypes = [vp, i64, i32, vp, vp]
    lib.vkb_apply_lut.argtypes = [vp, vp, i64, i32, vp, i32, vp]
    lib.vkb_gaussian_blur_u8.argtypes = [vp, vp, i32, i32, i32, POINTER(c_int32), i32, vp]
    lib.vkb_noise_philox.argtypes = [vp, vp, i64, i32, i32, c_double, c_double, ctypes.c_uint64, vp]
    lib.vkb_noise_field.argtypes = [vp, vp, i64, i32, i32, vp, vp]
    lib.vkb_streak_line.argtypes = [vp, i32, i32, i32, i32, i32, i32, i32, i32, i32,
                                    POINTER(c_float), c_float, vp]
    lib.vkb_fill_rects.argtypes = [vp, i32, i32, vp, i32, vp]
    lib.vkb_draw_ellipses.argtypes = [vp, i32, i32, vp, i32, i32, vp, i64, vp]
    lib.vkb_jpeg_round_trip_u8.argtypes = [vp, vp, i32, i32, i32, i32, vp, i64, vp]
    lib.vkb_streak_masks.argtypes = [vp, i32, i32, i32, vp, vp, i32, i32, POINTER(c_float),
                                     c_float, vp]
    lib.vkb_threshold_u8.argtypes = [vp, vp, i64, i32, i32, i32, vp]
    lib.vkb_zoom_in_blur_u8.argtypes = [vp, vp, i32, i32, i32, vp, i32, c_double, vp]
    lib.vkb_gather_pixels_u8.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp]
    lib.vkb_fog_draws.argtypes = [i32, POINTER(ctypes.c_int64)]
    lib.vkb_fog_mask.argtypes = [POINTER(FogParams), vp, vp, vp, vp, vp, vp]
    u64 = ctypes.c_uint64
    lib.vkb_glass_init.argtypes = [vp, vp, vp, i32, i32, vp, vp]
    lib.vkb_glass_round.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, i32, u64, u64, u64, u64, i32,
                                    ctypes.c_uint32, vp, vp, vp, vp, vp]
    lib.vkb_resize_u8.argtypes = [vp, i32, i32, vp, i32, i32, i32, i32, vp]
    lib.vkb_resize_f32.argtypes = [vp, i32, i32, vp, i32, i32, i32, i32, vp]
    lib.vkb_resize_f32_scaled.argtypes = [vp, i32, i32, vp, i32, i32, i32, i32, c_float, vp]
    lib.vkb_resize_mask_u8.argtypes = [vp, i32, i32, vp, i32, i32, i32, i32, vp]
    lib.vkb_filter2d_u8.argtypes = [vp, vp, i32, i32, i32, vp, i32, i32, vp]
    lib.vkb_fill_polygons.argtypes = [vp, i32, i32, i32, vp, vp, vp, i32, i32, vp, vp]
    lib.vkb_photo_chain_batched.argtypes = [vp, vp, i32, i32, vp]
    lib.vkb_channel_stats_batched.argtypes = [vp, i32, i32, vp, vp]
    lib.vkb_noise_philox_batched.argtypes = [vp, i32, i32, i32, vp]
    lib.vkb_background_compose.argtypes = [vp, i32, i32, i32, vp, i32, i32, POINTER(c_int32), i32,
                                           vp]
    lib.vkb_glyph_prepare.argtypes = [vp, vp, i32, vp]
    for name in EXPORTS:
        getattr(lib, name).restype = c_int32


EXPORTS = (
    'vkb_warp_fused', 'vkb_affine_points', 'vkb_grid_project', 'vkb_grid_finalize',
    'vkb_grid_layout', 'vkb_stage_params', 'vkb_grid_build', 'vkb_grid_remap', 'vkb_grid_points', 'vkb_grid_points_batched', 'vkb_affine_points_batched', 'vkb_fill_polygon',
    'vkb_blend_fill', 'vkb_blend_draw_list', 'vkb_cvt_color', 'vkb_color_ops',
    'vkb_channel_stats', 'vkb_histogram_u8', 'vkb_apply_lut', 'vkb_gaussian_blur_u8', 'vkb_noise_philox', 'vkb_noise_field',
    'vkb_streak_line', 'vkb_fill_rects', 'vkb_draw_ellipses', 'vkb_jpeg_round_trip_u8', 'vkb_streak_masks', 'vkb_photo_chain_batched',
    'vkb_channel_stats_batched', 'vkb_fill_polygons', 'vkb_filter2d_u8', 'vkb_resize_u8', 'vkb_resize_f32', 'vkb_resize_mask_u8', 'vkb_gather_pixels_u8', 'vkb_noise_philox_batched', 'vkb_zoom_in_blur_u8', 'vkb_threshold_u8',
    'vkb_background_compose', 'vkb_glyph_prepare', 'vkb_resize_f32_scaled', 'vkb_fog_draws', 'vkb_fog_mask', 'vkb_glass_init', 'vkb_glass_round',
)


def lib():
    """The loaded shared library; raises NativeError when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeError(
                f'{LIB_PATH} not found: build it with `python -m vkit_b200.build` '
                '(there is no CPU fallback for the distortion path).')
        loaded = ctypes.CDLL(LIB_PATH)
        _declare(loaded)
        _lib = loaded
    return _lib


def check(rc, what=''):
    if rc != 0:
        msg = lib().vkb_last_error().decode('utf-8', 'replace')
        raise NativeError(f'{what} failed (rc={rc}): {msg}')
