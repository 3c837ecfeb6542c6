"""Polygon: a tuple of points with a lazily rasterised coverage mask (vkit/element/polygon.py).

The coverage is what cv.fillPoly produces on the polygon's own bounding-box canvas
(polygon.py:70-77); here it comes from `vkb_fill_polygon` on the device.  Operations that
need shapely / pyclipper in the reference (area, vatti clipping, unions) are not part of the
distortion path and are not provided.
"""
import math
from typing import Iterable, Optional, Sequence, Tuple, Union

import attrs
import numpy as np

from .. import _native
from .. import device as dv
from ..utility import attrs_lazy_field
from .type import ElementSetOperationMode, Shapable

_T = Union[float, str]


def rasterize_polygon(np_xy: np.ndarray, shape: Tuple[int, int]):
    """uint8 CUDA tensor (h, w) with cv.fillPoly(canvas, [np_xy], 1) semantics."""
    height, width = shape
    canvas = dv.zeros((height, width), np.uint8)
    pts = dv.to_device(np.ascontiguousarray(np_xy, dtype=np.int32))
    _native.check(
        _native.lib().vkb_fill_polygon(dv.ptr(canvas), height, width, dv.ptr(pts),
                                       int(np_xy.shape[0]), 1, dv.stream_ptr()),
        'vkb_fill_polygon')
    return canvas


@attrs.define
class PolygonInternals:
    bounding_box: 'Box'
    np_self_relative_points: np.ndarray

    _self_relative_polygon: Optional['Polygon'] = attrs_lazy_field()
    _np_mask: Optional[np.ndarray] = attrs_lazy_field()
    _mask: Optional['Mask'] = attrs_lazy_field()

    @property
    def self_relative_polygon(self):
        if self._self_relative_polygon is None:
            self._self_relative_polygon = Polygon.from_np_array(self.np_self_relative_points)
        return self._self_relative_polygon

    @property
    def mask(self):
        if self._mask is None:
            coverage = rasterize_polygon(self.self_relative_polygon.to_np_array(),
                                         self.bounding_box.shape)
            self._mask = Mask(mat=coverage).to_box_attached(self.bounding_box)
        return self._mask

    @property
    def np_mask(self):
        if self._np_mask is None:
            self._np_mask = self.mask.mat.astype(np.bool_)
        return self._np_mask


@attrs.define(frozen=True, eq=False)
class Polygon:
    points: 'PointTuple'

    _internals: Optional[PolygonInternals] = attrs_lazy_field()

    def __attrs_post_init__(self):
        assert self.points

    @property
    def internals(self):
        if self._internals is None:
            # PointTuple.to_smooth_np_array holds ROUNDED coordinates (point.py:251-252).
            rel = self.to_smooth_np_array()
            y_min, y_max = rel[:, 1].min(), rel[:, 1].max()
            x_min, x_max = rel[:, 0].min(), rel[:, 0].max()
            rel[:, 0] -= x_min
            rel[:, 1] -= y_min
            bounding_box = Box(up=round(y_min), down=round(y_max), left=round(x_min),
                               right=round(x_max))
            object.__setattr__(self, '_internals',
                               PolygonInternals(bounding_box=bounding_box,
                                                np_self_relative_points=rel))
        return self._internals

    @classmethod
    def create(cls, points: Union['PointList', 'PointTuple', Iterable['Point']]):
        return cls(points=PointTuple(points))

    @property
    def num_points(self):
        return len(self.points)

    @property
    def bounding_box(self):
        return self.internals.bounding_box

    @property
    def self_relative_polygon(self):
        return self.internals.self_relative_polygon

    @property
    def mask(self):
        return self.internals.mask

    # ---- conversions -------------------------------------------------------------------
    @classmethod
    def from_xy_pairs(cls, xy_pairs: Iterable[Tuple[_T, _T]]):
        return cls(points=PointTuple.from_xy_pairs(xy_pairs))

    def to_xy_pairs(self):
        return self.points.to_xy_pairs()

    def to_smooth_xy_pairs(self):
        return self.points.to_smooth_xy_pairs()

    @classmethod
    def from_flatten_xy_pairs(cls, flatten_xy_pairs: Sequence[_T]):
        return cls(points=PointTuple.from_flatten_xy_pairs(flatten_xy_pairs))

    def to_flatten_xy_pairs(self):
        return self.points.to_flatten_xy_pairs()

    def to_smooth_flatten_xy_pairs(self):
        return self.points.to_smooth_flatten_xy_pairs()

    @classmethod
    def from_np_array(cls, np_points: np.ndarray):
        return cls(points=PointTuple.from_np_array(np_points))

    def to_np_array(self):
        return self.points.to_np_array()

    def to_smooth_np_array(self):
        return self.points.to_smooth_np_array()

    # ---- operators ---------------------------------------------------------------------
    def get_rectangular_height(self):
        assert self.num_points == 4
        up_left, up_right, down_right, down_left = self.points
        left = math.hypot(up_left.smooth_y - down_left.smooth_y,
                          up_left.smooth_x - down_left.smooth_x)
        right = math.hypot(up_right.smooth_y - down_right.smooth_y,
                           up_right.smooth_x - down_right.smooth_x)
        return (left + right) / 2

    def get_rectangular_width(self):
        assert self.num_points == 4
        up_left, up_right, down_right, down_left = self.points
        up = math.hypot(up_left.smooth_y - up_right.smooth_y, up_left.smooth_x - up_right.smooth_x)
        down = math.hypot(down_left.smooth_y - down_right.smooth_y,
                          down_left.smooth_x - down_right.smooth_x)
        return (up + down) / 2

    def to_clipped_points(self, shapable_or_shape):
        return self.points.to_clipped_points(shapable_or_shape)

    def to_clipped_polygon(self, shapable_or_shape):
        return Polygon(points=self.to_clipped_points(shapable_or_shape))

    def to_shifted_points(self, offset_y: int = 0, offset_x: int = 0):
        return self.points.to_shifted_points(offset_y=offset_y, offset_x=offset_x)

    def to_relative_points(self, origin_y: int, origin_x: int):
        return self.points.to_relative_points(origin_y=origin_y, origin_x=origin_x)

    def to_shifted_polygon(self, offset_y: int = 0, offset_x: int = 0):
        return Polygon(points=self.to_shifted_points(offset_y=offset_y, offset_x=offset_x))

    def to_relative_polygon(self, origin_y: int, origin_x: int):
        return Polygon(points=self.to_relative_points(origin_y=origin_y, origin_x=origin_x))

    def to_conducted_resized_polygon(self, shapable_or_shape, resized_height=None,
                                     resized_width=None):
        return Polygon(points=self.points.to_conducted_resized_points(
            shapable_or_shape, resized_height, resized_width))

    def to_resized_polygon(self, resized_height=None, resized_width=None):
        return self.to_conducted_resized_polygon(self.bounding_box.shape, resized_height,
                                                 resized_width)

    def to_bounding_box(self):
        return self.bounding_box

    # ---- fills / extraction delegate to the boxed coverage mask (polygon.py:439-502) -----
    def fill_np_array(self, mat, value, alpha=1.0, keep_max_value=False, keep_min_value=False):
        self.mask.fill_np_array(mat=mat, value=value, alpha=alpha, keep_max_value=keep_max_value,
                                keep_min_value=keep_min_value)

    def extract_mask(self, mask: 'Mask'):
        return self.mask.extract_mask(mask)

    def fill_mask(self, mask: 'Mask', value=1, keep_max_value=False, keep_min_value=False):
        self.mask.fill_mask(mask=mask, value=value, keep_max_value=keep_max_value,
                            keep_min_value=keep_min_value)

    def extract_score_map(self, score_map: 'ScoreMap'):
        return self.mask.extract_score_map(score_map)

    def fill_score_map(self, score_map: 'ScoreMap', value, keep_max_value=False,
                       keep_min_value=False):
        self.mask.fill_score_map(score_map=score_map, value=value, keep_max_value=keep_max_value,
                                 keep_min_value=keep_min_value)

    def extract_image(self, image: 'Image'):
        return self.mask.extract_image(image)

    def fill_image(self, image: 'Image', value, alpha=1.0):
        self.mask.fill_image(image=image, value=value, alpha=alpha)


def generate_fill_by_polygons_mask(shape: Tuple[int, int], polygons: Iterable[Polygon],
                                   mode: ElementSetOperationMode):
    if mode == ElementSetOperationMode.UNION:
        return None
    return Mask.from_polygons(shape, polygons, mode)


from .point import Point, PointList, PointTuple  # noqa: E402
from .box import Box  # noqa: E402
from .mask import Mask  # noqa: E402
from .score_map import ScoreMap  # noqa: E402
from .image import Image  # noqa: E402
