"""Grid-based rendering shared by camera_* and similarity_mls
(vkit/mechanism/distortion/geometric/grid_rendering/{type,grid_creator,grid_blender,interface}.py).

The reference projects a sparse lattice, fits one homography per cell, rasterises every
destination cell in Python to fill dense map_x / map_y arrays and calls cv.remap three times.
Here the state holds a `GridBatch` (device plan); the dense maps are never materialised.
"""
from typing import Generic, List, Optional, Tuple, Type, TypeVar

import numpy as np
from numpy.random import Generator as RandomGenerator

from vkit_b200 import _native as nv
from vkit_b200 import device as dv
from vkit_b200.element import Image, Mask, Point, PointList, Polygon, ScoreMap

from ..interface import Distortion, DistortionConfig, DistortionState
from ._gridcore import GridBatch, planes_record
from ._hostmath import lattice_axis

_T_CONFIG = TypeVar('_T_CONFIG', bound=DistortionConfig)


class ImageGrid:
    """Read-only view of a lattice for reference-style consumers (type.py:25-143)."""

    def __init__(self, points_xy: np.ndarray, grid_size: Optional[int] = None):
        self._xy = points_xy  # (rows, cols, 2) as (x, y)
        self.grid_size = grid_size
        self._points_2d: Optional[List[PointList]] = None

    @property
    def points_2d(self) -> List[PointList]:
        if self._points_2d is None:
            self._points_2d = [
                PointList(Point.create(y=int(y), x=int(x)) for x, y in row) for row in self._xy
            ]
        return self._points_2d

    @property
    def num_rows(self):
        return int(self._xy.shape[0])

    @property
    def num_cols(self):
        return int(self._xy.shape[1])

    @property
    def shape(self):
        return self.num_rows, self.num_cols

    @property
    def image_height(self):
        return int(self._xy[..., 1].max()) + 1

    @property
    def image_width(self):
        return int(self._xy[..., 0].max()) + 1

    @property
    def image_shape(self):
        return self.image_height, self.image_width

    @property
    def flatten_points(self):
        return PointList(point for row in self.points_2d for point in row)

    def border_xy(self) -> np.ndarray:
        """Clockwise border of the lattice (generate_border_polygon, type.py:130-143)."""
        g = self._xy
        parts = [g[0, :], g[1:, -1], g[-1, -2::-1], g[-2:0:-1, 0]]
        return np.concatenate(parts, axis=0)

    def generate_border_polygon(self):
        return Polygon.from_np_array(self.border_xy())


def create_src_image_grid(height: int, width: int, grid_size: int):
    ys = lattice_axis(height, grid_size)
    xs = lattice_axis(width, grid_size)
    xy = np.stack(np.meshgrid(xs, ys), axis=-1).astype(np.int32)
    return ImageGrid(xy, grid_size=grid_size)


class DistortionStateImageGridBased(DistortionState[_T_CONFIG]):
    """State of a grid op: a one-page device plan plus the reference's public attributes."""

    def initialize_grid_plan(self, page_record: np.ndarray, keepalive=(),
                             given_lattice: Optional[np.ndarray] = None):
        self.plan = GridBatch(page_record.reshape(1), keepalive=keepalive,
                              given_lattice=given_lattice)
        meta = self.plan.meta[0]
        self.shift_amount_y = int(meta['shift_y'])
        self.shift_amount_x = int(meta['shift_x'])
        self.resize_ratio_y = float(meta['resize_ratio_y'])
        self.resize_ratio_x = float(meta['resize_ratio_x'])
        self._result_shape = (int(meta['dst_h']), int(meta['dst_w']))
        self._src_image_grid = None
        self._dst_image_grid = None

    @property
    def src_image_grid(self) -> ImageGrid:
        if self._src_image_grid is None:
            page = self.plan.pages[0]
            self._src_image_grid = create_src_image_grid(int(page['src_h']), int(page['src_w']),
                                                         int(page['grid_size']))
        return self._src_image_grid

    @property
    def dst_image_grid(self) -> ImageGrid:
        if self._dst_image_grid is None:
            self._dst_image_grid = ImageGrid(self.plan.lattice_points(0))
        return self._dst_image_grid

    def shift_and_resize_point(self, point: Point):
        return Point.create(
            y=(point.smooth_y - self.shift_amount_y) * self.resize_ratio_y,
            x=(point.smooth_x - self.shift_amount_x) * self.resize_ratio_x,
        )

    @property
    def result_shape(self):
        return self._result_shape


_T_STATE = TypeVar('_T_STATE', bound=DistortionStateImageGridBased)


class FuncImageGridBased(Generic[_T_CONFIG, _T_STATE]):

    @classmethod
    def func_planes(cls, config, state, image, mask, score_map, rng):
        """Image + Mask + ScoreMap through ONE fused remap launch."""
        assert state
        rec, out_image, out_mask, out_score = planes_record(image, mask, score_map,
                                                            state.result_shape)
        state.plan.remap(rec.reshape(1))
        return (
            Image(mat=out_image, mode=image.mode) if image is not None else None,
            Mask(mat=out_mask) if mask is not None else None,
            ScoreMap(mat=out_score, skip_prob_check=True) if score_map is not None else None,
        )

    @classmethod
    def func_image(cls, config, state, image: Image, rng: Optional[RandomGenerator]):
        return cls.func_planes(config, state, image, None, None, rng)[0]

    @classmethod
    def func_score_map(cls, config, state, score_map: ScoreMap, rng: Optional[RandomGenerator]):
        return cls.func_planes(config, state, None, None, score_map, rng)[2]

    @classmethod
    def func_mask(cls, config, state, mask: Mask, rng: Optional[RandomGenerator]):
        return cls.func_planes(config, state, None, mask, None, rng)[1]

    @classmethod
    def func_active_mask(cls, config, state, shape: Tuple[int, int],
                         rng: Optional[RandomGenerator]):
        # filled border polygon of the dst lattice (interface.py:177-192)
        assert state
        height, width = state.result_shape
        border = np.ascontiguousarray(state.dst_image_grid.border_xy(), dtype=np.int32)
        canvas = dv.zeros((height, width), np.uint8)
        pts = dv.to_device(border)
        nv.check(nv.lib().vkb_fill_polygon(dv.ptr(canvas), height, width, dv.ptr(pts),
                                           int(border.shape[0]), 1, dv.stream_ptr()),
                 'vkb_fill_polygon')
        return Mask(mat=canvas)

    @classmethod
    def func_points(cls, config, state, shape: Tuple[int, int], points,
                    rng: Optional[RandomGenerator]):
        """Batched form of func_point (interface.py:194-216): the cell comes from the ROUNDED
        point, the transform acts on the smooth coordinates."""
        assert state
        points = list(points)
        grid_size = int(state.plan.pages[0]['grid_size'])
        rows = int(state.plan.pages[0]['rows'])
        cols = int(state.plan.pages[0]['cols'])
        xy = np.asarray([(p.smooth_x, p.smooth_y) for p in points], dtype=np.float64).reshape(-1, 2)
        rc = np.asarray([(p.y // grid_size, p.x // grid_size) for p in points],
                        dtype=np.int32).reshape(-1, 2)
        if len(points) and (rc.min() < 0 or rc[:, 0].max() >= rows - 1 or rc[:, 1].max() >= cols - 1):
            raise IndexError('point outside the source lattice')  # the reference raises too
        out = state.plan.transform_points(0, xy, rc)
        return [Point.create(y=float(y), x=float(x)) for x, y in out]

    @classmethod
    def func_point(cls, config, state, shape: Tuple[int, int], point: Point,
                   rng: Optional[RandomGenerator]):
        return cls.func_points(config, state, shape, [point], rng)[0]


class DistortionImageGridBased(Distortion[_T_CONFIG, _T_STATE]):

    def __init__(self, config_cls: Type[_T_CONFIG], state_cls: Type[_T_STATE]):
        func_cls = FuncImageGridBased
        super().__init__(
            config_cls=config_cls,
            state_cls=state_cls,
            func_image=func_cls.func_image,
            func_mask=func_cls.func_mask,
            func_score_map=func_cls.func_score_map,
            func_active_mask=func_cls.func_active_mask,
            func_point=func_cls.func_point,
        )
        self.func_planes = func_cls.func_planes
        self._func_points_batched = func_cls.func_points

    def distort_points_based_on_internals(self, internals, points):
        # one kernel launch for all points instead of a Python call per point
        from vkit_b200.element import PointTuple
        internals.restore_rng_if_supported()
        return PointTuple(self._func_points_batched(internals.config, internals.state,
                                                    internals.shape, PointList(points),
                                                    internals.rng))

    def distort_polygons_based_on_internals(self, internals, polygons):
        # the vertices of ALL polygons through one launch (the reference falls back to one
        # func_point call per vertex, interface.py:694-715)
        from vkit_b200.element import Polygon, PointTuple
        internals.restore_rng_if_supported()
        polygons = list(polygons)
        flat = PointList()
        for polygon in polygons:
            flat.extend(polygon.points)
        moved = self._func_points_batched(internals.config, internals.state, internals.shape, flat,
                                          internals.rng)
        out, begin = [], 0
        for polygon in polygons:
            end = begin + polygon.num_points
            out.append(Polygon.create(points=moved[begin:end]))
            begin = end
        return out

    def distort_polygon_based_on_internals(self, internals, polygon):
        return self.distort_polygons_based_on_internals(internals, [polygon])[0]
