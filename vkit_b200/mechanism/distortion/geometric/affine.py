"""rotate / shear_hori / shear_vert / skew_hori / skew_vert
(vkit/mechanism/distortion/geometric/affine.py).

The states build the same forward matrix and `dsize` as the reference; the pixels go through
`vkb_warp_fused`, which reproduces cv.warpAffine / cv.warpPerspective bit-exactly (fixed-point
coordinates, 1/32 px bilinear) for Image, Mask and ScoreMap in one launch.
"""
import math
from typing import Iterable, List, Optional, Sequence, Tuple, Type, TypeVar, Union

import attrs
import numpy as np
from numpy.random import Generator as RandomGenerator

from vkit_b200 import _native as nv
from vkit_b200 import device as dv
from vkit_b200.element import Image, Mask, Point, PointList, PointTuple, Polygon, ScoreMap

from ..interface import Distortion, DistortionConfig, DistortionState
from ._gridcore import planes_record
from ._hostmath import homography_4pt, invert_affine


def warp_planes(trans_mat: np.ndarray, dsize: Tuple[int, int], image=None, mask=None,
                score_map=None):
    """One fused launch over the present containers. Returns (Image, Mask, ScoreMap) tensors."""
    dst_w, dst_h = dsize
    rec = np.zeros((), dtype=nv.WARP_PAGE_DTYPE)
    planes, out_image, out_mask, out_score = planes_record(image, mask, score_map,
                                                           (dst_h, dst_w))
    rec['planes'] = planes
    if trans_mat.shape[0] == 2:
        rec['kind'] = nv.WARP_AFFINE
        rec['inv'][:6] = invert_affine(trans_mat).reshape(-1)
    else:
        assert trans_mat.shape[0] == 3
        rec['kind'] = nv.WARP_PERSPECTIVE
        rec['inv'][:] = np.linalg.inv(np.asarray(trans_mat, dtype=np.float64)).reshape(-1)
    pages_dev = dv.upload_structs(rec.reshape(1))
    nv.check(nv.lib().vkb_warp_fused(dv.ptr(pages_dev), 1, dst_h, dst_w, dv.stream_ptr()),
             'vkb_warp_fused')
    return out_image, out_mask, out_score


def affine_np_points(trans_mat: np.ndarray, np_points: np.ndarray) -> np.ndarray:
    """(N, 2) float32 (x, y) through the forward matrix (affine.py:46-64), on the device."""
    n = int(np_points.shape[0])
    if n == 0:
        return np.zeros((0, 2), dtype=np.float32 if trans_mat.shape[0] == 2 else np.float64)
    rows = int(trans_mat.shape[0])
    f32_math = int(trans_mat.dtype == np.float32 and rows == 2)
    mat = (ctypes_doubles(np.asarray(trans_mat, dtype=np.float64).reshape(-1)))
    xy_dev = dv.to_device(np.ascontiguousarray(np_points, dtype=np.float64))
    out = dv.empty((n, 2), np.float64)
    nv.check(nv.lib().vkb_affine_points(mat, rows, dv.ptr(xy_dev), dv.ptr(out), n, f32_math,
                                        dv.stream_ptr()), 'vkb_affine_points')
    return dv.to_host(out)


def ctypes_doubles(values):
    import ctypes
    arr = (ctypes.c_double * len(values))(*[float(v) for v in values])
    return arr


def affine_points(trans_mat: np.ndarray, points: PointTuple):
    return PointTuple.from_np_array(affine_np_points(trans_mat, points.to_smooth_np_array()))


def affine_polygons(trans_mat: np.ndarray, polygons: Sequence[Polygon]) -> Sequence[Polygon]:
    ranges: List[Tuple[int, int]] = []
    points = PointList()
    for polygon in polygons:
        ranges.append((len(points), len(points) + polygon.num_points))
        points.extend(polygon.points)
    new_np_points = affine_np_points(trans_mat, points.to_smooth_np_array())
    return [Polygon.from_np_array(new_np_points[begin:end]) for begin, end in ranges]


def convert_dsize_to_result_shape(dsize: Optional[Tuple[int, int]]):
    if dsize:
        return dsize[1], dsize[0]


@attrs.define
class ShearHoriConfig(DistortionConfig):
    # (-90, 90), positive = rightward
    angle: int

    @property
    def is_nop(self):
        return self.angle == 0


class ShearHoriState(DistortionState[ShearHoriConfig]):

    def __init__(self, config: ShearHoriConfig, shape: Tuple[int, int],
                 rng: Optional[RandomGenerator]):
        tan_phi = math.tan(math.radians(config.angle))
        height, width = shape
        shift_x = abs(height * tan_phi)
        self.dsize = (math.ceil(width + shift_x), height)
        if config.angle < 0:
            self.trans_mat = np.asarray([(1, -tan_phi, 0), (0, 1, 0)], dtype=np.float32)
        elif config.angle > 0:
            self.trans_mat = np.asarray([(1, -tan_phi, shift_x), (0, 1, 0)], dtype=np.float32)
        else:
            self.trans_mat = None
            self.dsize = None

    @property
    def result_shape(self):
        return convert_dsize_to_result_shape(self.dsize)


@attrs.define
class ShearVertConfig(DistortionConfig):
    # (-90, 90), positive = downward
    angle: int

    @property
    def is_nop(self):
        return self.angle == 0


class ShearVertState(DistortionState[ShearVertConfig]):

    def __init__(self, config: ShearVertConfig, shape: Tuple[int, int],
                 rng: Optional[RandomGenerator]):
        tan_abs_phi = math.tan(math.radians(abs(config.angle)))
        height, width = shape
        shift_y = width * tan_abs_phi
        self.dsize = (width, math.ceil(height + shift_y))
        if config.angle < 0:
            self.trans_mat = np.asarray([(1, 0, 0), (-tan_abs_phi, 1, shift_y)], dtype=np.float32)
        elif config.angle > 0:
            self.trans_mat = np.asarray([(1, 0, 0), (tan_abs_phi, 1, 0)], dtype=np.float32)
        else:
            self.trans_mat = None
            self.dsize = None

    @property
    def result_shape(self):
        return convert_dsize_to_result_shape(self.dsize)


@attrs.define
class RotateConfig(DistortionConfig):
    # [0, 360], clockwise
    angle: int

    @property
    def is_nop(self):
        return self.angle == 0


class RotateState(DistortionState[RotateConfig]):

    def __init__(self, config: RotateConfig, shape: Tuple[int, int],
                 rng: Optional[RandomGenerator]):
        height, width = shape
        rad = math.radians(config.angle % 360)
        sin_r, cos_r = math.sin(rad), math.cos(rad)
        shift_x = shift_y = 0
        # bounding box of the rotated page, quadrant by quadrant (affine.py:224-258)
        if rad <= math.pi / 2:
            shift_x = height * sin_r
            dst_width = height * sin_r + width * cos_r
            dst_height = height * cos_r + width * sin_r
        elif rad <= math.pi:
            t = rad - math.pi / 2
            shift_x = width * math.sin(t) + height * math.cos(t)
            shift_y = height * math.sin(t)
            dst_width = shift_x
            dst_height = shift_y + width * math.cos(t)
        elif rad < math.pi * 3 / 2:
            t = rad - math.pi
            shift_x = width * math.cos(t)
            shift_y = width * math.sin(t) + height * math.cos(t)
            dst_width = shift_x + height * math.sin(t)
            dst_height = shift_y
        else:
            t = rad - math.pi * 3 / 2
            shift_y = width * math.cos(t)
            dst_width = width * math.sin(t) + height * math.cos(t)
            dst_height = shift_y + height * math.sin(t)
        shift_x = math.ceil(shift_x)
        shift_y = math.ceil(shift_y)
        self.trans_mat = np.asarray([(cos_r, -sin_r, shift_x), (sin_r, cos_r, shift_y)],
                                    dtype=np.float32)
        self.dsize = (math.ceil(dst_width), math.ceil(dst_height))

    @property
    def result_shape(self):
        return convert_dsize_to_result_shape(self.dsize)


@attrs.define
class SkewHoriConfig(DistortionConfig):
    # (-1, 0]: shrink the left side; [0, 1): shrink the right side
    ratio: float

    @property
    def is_nop(self):
        return self.ratio == 0


class SkewHoriState(DistortionState[SkewHoriConfig]):

    def __init__(self, config: SkewHoriConfig, shape: Tuple[int, int],
                 rng: Optional[RandomGenerator]):
        height, width = shape
        src = [(0, 0), (width - 1, 0), (width - 1, height - 1), (0, height - 1)]
        shrink = round(height * abs(config.ratio))
        shrink_up = shrink // 2
        shrink_down = shrink - shrink_up
        if config.ratio < 0:
            dst = [(0, shrink_up), (width - 1, 0), (width - 1, height - 1),
                   (0, height - shrink_down - 1)]
        else:
            dst = [(0, 0), (width - 1, shrink_up), (width - 1, height - shrink_down - 1),
                   (0, height - 1)]
        self.trans_mat = homography_4pt(src, dst)
        self.dsize = (width, height)

    @property
    def result_shape(self):
        return convert_dsize_to_result_shape(self.dsize)


@attrs.define
class SkewVertConfig(DistortionConfig):
    # (-1, 0]: shrink the up side; [0, 1): shrink the down side
    ratio: float

    @property
    def is_nop(self):
        return self.ratio == 0


class SkewVertState(DistortionState[SkewVertConfig]):

    def __init__(self, config: SkewVertConfig, shape: Tuple[int, int],
                 rng: Optional[RandomGenerator]):
        height, width = shape
        src = [(0, 0), (width - 1, 0), (width - 1, height - 1), (0, height - 1)]
        shrink = round(width * abs(config.ratio))
        shrink_left = shrink // 2
        shrink_right = shrink - shrink_left
        if config.ratio < 0:
            dst = [(shrink_left, 0), (width - shrink_right - 1, 0), (width - 1, height - 1),
                   (0, height - 1)]
        else:
            # bottom-left x uses shrink_right, like the reference (affine.py:383)
            dst = [(0, 0), (width - 1, 0), (width - shrink_right - 1, height - 1),
                   (shrink_right, height - 1)]
        self.trans_mat = homography_4pt(src, dst)
        self.dsize = (width, height)

    @property
    def result_shape(self):
        return convert_dsize_to_result_shape(self.dsize)


_T_AFFINE_CONFIG = TypeVar('_T_AFFINE_CONFIG', ShearHoriConfig, ShearVertConfig, RotateConfig,
                           SkewHoriConfig, SkewVertConfig)
_T_AFFINE_STATE = TypeVar('_T_AFFINE_STATE', ShearHoriState, ShearVertState, RotateState,
                          SkewHoriState, SkewVertState)


def affine_trait_func_planes(config, state, image, mask, score_map, rng):
    assert state
    if config.is_nop:
        return image, mask, score_map
    assert state.trans_mat is not None and state.dsize is not None
    out_image, out_mask, out_score = warp_planes(state.trans_mat, state.dsize, image, mask,
                                                 score_map)
    return (
        Image(mat=out_image) if image is not None else None,  # mode is re-inferred (affine.py:436)
        Mask(mat=out_mask) if mask is not None else None,
        ScoreMap(mat=out_score, skip_prob_check=True) if score_map is not None else None,
    )


def affine_trait_func_image(config, state, image: Image, rng: Optional[RandomGenerator]):
    return affine_trait_func_planes(config, state, image, None, None, rng)[0]


def affine_trait_func_score_map(config, state, score_map: ScoreMap,
                                rng: Optional[RandomGenerator]):
    return affine_trait_func_planes(config, state, None, None, score_map, rng)[2]


def affine_trait_func_mask(config, state, mask: Mask, rng: Optional[RandomGenerator]):
    return affine_trait_func_planes(config, state, None, mask, None, rng)[1]


def affine_trait_func_points(config, state, shape: Tuple[int, int],
                             points: Union[PointList, PointTuple, Iterable[Point]],
                             rng: Optional[RandomGenerator]):
    assert state
    points = PointTuple(points)
    if config.is_nop:
        return points
    assert state.trans_mat is not None
    return affine_points(state.trans_mat, points)


def affine_trait_func_polygons(config, state, shape: Tuple[int, int],
                               polygons: Iterable[Polygon], rng: Optional[RandomGenerator]):
    assert state
    polygons = tuple(polygons)
    if config.is_nop:
        return polygons
    assert state.trans_mat is not None
    return affine_polygons(state.trans_mat, polygons)


class DistortionAffine(Distortion[_T_AFFINE_CONFIG, _T_AFFINE_STATE]):

    def __init__(self, config_cls: Type[_T_AFFINE_CONFIG], state_cls: Type[_T_AFFINE_STATE]):
        super().__init__(
            config_cls=config_cls,
            state_cls=state_cls,
            func_image=affine_trait_func_image,
            func_mask=affine_trait_func_mask,
            func_score_map=affine_trait_func_score_map,
            func_points=affine_trait_func_points,
            func_polygons=affine_trait_func_polygons,
        )
        self.func_planes = affine_trait_func_planes


shear_hori = DistortionAffine(config_cls=ShearHoriConfig, state_cls=ShearHoriState)
shear_vert = DistortionAffine(config_cls=ShearVertConfig, state_cls=ShearVertState)
rotate = DistortionAffine(config_cls=RotateConfig, state_cls=RotateState)
skew_hori = DistortionAffine(config_cls=SkewHoriConfig, state_cls=SkewHoriState)
skew_vert = DistortionAffine(config_cls=SkewVertConfig, state_cls=SkewVertState)
