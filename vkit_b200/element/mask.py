"""Mask: uint8 HxW, active where mat > 0 (vkit/element/mask.py)."""
from typing import Iterable, Optional, Tuple, Union

import attrs
import numpy as np

from .. import _native
from .. import device as dv
from ._storage import DualStorage
from .type import ElementSetOperationMode, Shapable


@attrs.define
class MaskSetItemConfig:
    value: Union['Mask', np.ndarray, int] = 1
    keep_max_value: bool = False
    keep_min_value: bool = False


@attrs.define(frozen=True, eq=False)
class Mask(DualStorage, Shapable):
    _mat: object = attrs.field(alias='mat')
    box: Optional['Box'] = None

    _alt: object = attrs.field(default=None, init=False, repr=False)
    _np_mask: Optional[np.ndarray] = attrs.field(default=None, init=False, repr=False)

    def __attrs_post_init__(self):
        self._adopt(self._mat)
        if self.mat_dtype != np.uint8:
            raise RuntimeError('mat.dtype != np.uint8')
        if self.mat_ndim != 2:
            raise RuntimeError('ndim should == 2.')
        if self.box and self.shape != self.box.shape:
            raise RuntimeError('self.shape != box.shape.')

    # ---- constructors ------------------------------------------------------------------
    @classmethod
    def from_shape(cls, shape: Tuple[int, int], value: int = 0):
        height, width = shape
        assert value in (0, 1)
        init = np.zeros if value == 0 else np.ones
        return cls(mat=init((height, width), dtype=np.uint8))

    @classmethod
    def from_shapable(cls, shapable: Shapable, value: int = 0):
        return cls.from_shape(shape=shapable.shape, value=value)

    @classmethod
    def _unpack_shape_or_box(cls, shape_or_box):
        if isinstance(shape_or_box, Box):
            return shape_or_box.shape, shape_or_box
        return shape_or_box, None

    @classmethod
    def _from_np_active_count(cls, shape, mode, np_active_count, attached_box):
        if mode == ElementSetOperationMode.UNION:
            mat = np_active_count > 0
        elif mode == ElementSetOperationMode.DISTINCT:
            mat = np_active_count == 1
        elif mode == ElementSetOperationMode.INTERSECT:
            mat = np_active_count > 1
        else:
            raise NotImplementedError()
        mask = cls(mat=mat.astype(np.uint8))
        if attached_box:
            mask = mask.to_box_attached(attached_box)
        return mask

    @classmethod
    def from_boxes(cls, shape_or_box, boxes: Iterable['Box'],
                   mode: ElementSetOperationMode = ElementSetOperationMode.UNION):
        # Set-operation bookkeeping (mask.py:154-175): counts on the host, tiny and label-side.
        shape, attached_box = cls._unpack_shape_or_box(shape_or_box)
        count = np.zeros(shape, dtype=np.int32)
        for box in boxes:
            if attached_box:
                box = box.to_relative_box(origin_y=attached_box.up, origin_x=attached_box.left)
            box.extract_np_array(count)[...] += 1
        return cls._from_np_active_count(shape, mode, count, attached_box)

    @classmethod
    def from_polygons(cls, shape_or_box, polygons: Iterable['Polygon'],
                      mode: ElementSetOperationMode = ElementSetOperationMode.UNION):
        shape, attached_box = cls._unpack_shape_or_box(shape_or_box)
        count = np.zeros(shape, dtype=np.int32)
        for polygon in polygons:
            box = polygon.bounding_box
            if attached_box:
                box = box.to_relative_box(origin_y=attached_box.up, origin_x=attached_box.left)
            box.extract_np_array(count)[polygon.internals.np_mask] += 1
        return cls._from_np_active_count(shape, mode, count, attached_box)

    @classmethod
    def from_masks(cls, shape_or_box, masks: Iterable['Mask'],
                   mode: ElementSetOperationMode = ElementSetOperationMode.UNION):
        shape, attached_box = cls._unpack_shape_or_box(shape_or_box)
        count = np.zeros(shape, dtype=np.int32)
        for mask in masks:
            view = count
            if mask.box:
                box = mask.box
                if attached_box:
                    box = box.to_relative_box(origin_y=attached_box.up, origin_x=attached_box.left)
                view = box.extract_np_array(count)
            view[mask.np_mask] += 1
        return cls._from_np_active_count(shape, mode, count, attached_box)

    @classmethod
    def from_score_maps(cls, shape_or_box, score_maps: Iterable['ScoreMap'],
                        mode: ElementSetOperationMode = ElementSetOperationMode.UNION):
        return cls.from_masks(shape_or_box, (sm.to_mask() for sm in score_maps), mode)

    # ---- properties --------------------------------------------------------------------
    @property
    def equivalent_box(self):
        return self.box or Box.from_shapable(self)

    @property
    def np_mask(self):
        if self._np_mask is None:
            object.__setattr__(self, '_np_mask', self.mat > 0)
        return self._np_mask

    def set_np_mask_out_of_date(self):
        object.__setattr__(self, '_np_mask', None)

    def _after_host_write(self):
        self.set_np_mask_out_of_date()

    def _after_device_write(self):
        super()._after_device_write()
        self.set_np_mask_out_of_date()

    # ---- operators ---------------------------------------------------------------------
    def copy(self):
        return attrs.evolve(self, mat=self._clone_storage())

    def assign_mat(self, mat: np.ndarray):
        self._adopt(mat)
        self.set_np_mask_out_of_date()

    def fill_by_boxes(self, boxes: Iterable['Box'], value=1,
                      mode: ElementSetOperationMode = ElementSetOperationMode.UNION,
                      keep_max_value: bool = False, keep_min_value: bool = False):
        boxes = list(boxes)
        boxes_mask = generate_fill_by_boxes_mask(self.shape, boxes, mode)
        if boxes_mask is None:
            for box in boxes:
                box.fill_mask(self, value, keep_max_value=keep_max_value,
                              keep_min_value=keep_min_value)
        else:
            boxes_mask.fill_mask(self, value, keep_max_value=keep_max_value,
                                 keep_min_value=keep_min_value)

    def fill_by_polygons(self, polygons: Iterable['Polygon'], value=1,
                         mode: ElementSetOperationMode = ElementSetOperationMode.UNION,
                         keep_max_value: bool = False, keep_min_value: bool = False):
        polygons = list(polygons)
        polygons_mask = generate_fill_by_polygons_mask(self.shape, polygons, mode)
        if polygons_mask is None:
            for polygon in polygons:
                polygon.fill_mask(self, value, keep_max_value=keep_max_value,
                                  keep_min_value=keep_min_value)
        else:
            polygons_mask.fill_mask(self, value, keep_max_value=keep_max_value,
                                    keep_min_value=keep_min_value)

    def __setitem__(self, element, config):
        if isinstance(config, MaskSetItemConfig):
            value, keep_max, keep_min = config.value, config.keep_max_value, config.keep_min_value
        else:
            value, keep_max, keep_min = config, False, False
        element.fill_mask(mask=self, value=value, keep_max_value=keep_max, keep_min_value=keep_min)

    def __getitem__(self, element):
        return element.extract_mask(self)

    def to_inverted_mask(self):
        if self.on_device:
            return attrs.evolve(self, mat=(self.dev == 0).to(self.dev.dtype))
        return attrs.evolve(self, mat=(~self.np_mask).astype(np.uint8))

    def to_resized_mask(self, resized_height: Optional[int] = None,
                        resized_width: Optional[int] = None, cv_resize_interpolation: int = 2,
                        binarization_threshold: int = 0):
        """element/mask.py:454-479: (mask > 0) * 255 -> cv.resize -> > threshold.  NEAREST (0),
        LINEAR (1), AREA (3), LANCZOS4 (4), LINEAR_EXACT (5), NEAREST_EXACT (6) are bit exact, CUBIC (2, the default)
        follows the wheel's IPP cubic (see Image.to_resized_image)."""
        from .opt import generate_resized_shape
        assert not self.box
        resized_height, resized_width = generate_resized_shape(self.height, self.width,
                                                               resized_height, resized_width)
        if cv_resize_interpolation not in (0, 1, 2, 3, 4, 5, 6):
            raise NotImplementedError(
                'to_resized_mask: cv.INTER_NEAREST / LINEAR / CUBIC / AREA / LANCZOS4 / LINEAR_EXACT / '
                'NEAREST_EXACT have device kernels')
        src = self.dev
        dst = dv.empty((resized_height, resized_width), np.uint8)
        # both binarisations are fused into the resize kernel (one launch instead of three)
        _native.check(_native.lib().vkb_resize_mask_u8(
            dv.ptr(src), self.height, self.width, dv.ptr(dst), resized_height, resized_width,
            cv_resize_interpolation, int(binarization_threshold), dv.stream_ptr()),
            'vkb_resize_mask_u8')
        return Mask(mat=dst)

    def to_conducted_resized_mask(self, shapable_or_shape, resized_height: Optional[int] = None,
                                  resized_width: Optional[int] = None,
                                  cv_resize_interpolation: int = 2, binarization_threshold: int = 0):
        """element/mask.py:481-503: resize the attached box and the mask it holds together."""
        assert self.box
        resized_box = self.box.to_conducted_resized_box(
            shapable_or_shape=shapable_or_shape, resized_height=resized_height,
            resized_width=resized_width)
        resized = self.to_box_detached().to_resized_mask(
            resized_height=resized_box.height, resized_width=resized_box.width,
            cv_resize_interpolation=cv_resize_interpolation,
            binarization_threshold=binarization_threshold)
        return resized.to_box_attached(resized_box)

    def to_shifted_mask(self, offset_y: int = 0, offset_x: int = 0):
        assert self.box
        return attrs.evolve(self, box=self.box.to_shifted_box(offset_y=offset_y,
                                                              offset_x=offset_x))

    def to_cropped_mask(self, up=None, down=None, left=None, right=None):
        assert not self.box
        up = up or 0
        down = down or self.height - 1
        left = left or 0
        right = right or self.width - 1
        return attrs.evolve(self, mat=self._crop_storage(up, down, left, right))

    def to_box_attached(self, box: 'Box'):
        assert self.height == box.height
        assert self.width == box.width
        return attrs.evolve(self, box=box)

    def to_box_detached(self):
        assert self.box
        return attrs.evolve(self, box=None)

    def fill_np_array(self, mat, value, alpha=1.0, keep_max_value=False, keep_min_value=False):
        self.equivalent_box.fill_np_array(mat=mat, value=value, np_mask=self.np_mask, alpha=alpha,
                                          keep_max_value=keep_max_value,
                                          keep_min_value=keep_min_value)

    def extract_mask(self, mask: 'Mask'):
        mask = self.equivalent_box.extract_mask(mask).copy()
        self.to_inverted_mask().fill_mask(mask, value=0)
        return mask

    def fill_mask(self, mask: 'Mask', value=1, keep_max_value=False, keep_min_value=False):
        self.equivalent_box.fill_mask(mask=mask, value=value, mask_mask=self,
                                      keep_max_value=keep_max_value,
                                      keep_min_value=keep_min_value)

    def extract_score_map(self, score_map: 'ScoreMap'):
        score_map = self.equivalent_box.extract_score_map(score_map).copy()
        self.to_inverted_mask().fill_score_map(score_map, value=0.0)
        return score_map

    def fill_score_map(self, score_map: 'ScoreMap', value, keep_max_value=False,
                       keep_min_value=False):
        self.equivalent_box.fill_score_map(score_map=score_map, value=value, score_map_mask=self,
                                           keep_max_value=keep_max_value,
                                           keep_min_value=keep_min_value)

    def to_score_map(self):
        if self.on_device:
            return ScoreMap(mat=(self.dev > 0).to(dv.torch().float32), box=self.box,
                            skip_prob_check=True)
        return ScoreMap(mat=self.np_mask.astype(np.float32), box=self.box)

    def extract_image(self, image: 'Image'):
        image = self.equivalent_box.extract_image(image).copy()
        self.to_inverted_mask().fill_image(image, value=0)
        return image

    def fill_image(self, image: 'Image', value, alpha=1.0):
        self.equivalent_box.fill_image(image=image, value=value, image_mask=self, alpha=alpha)

    def to_external_box(self):
        np_mask = self.np_mask
        rows = np.nonzero(np_mask.any(axis=1))[0]
        cols = np.nonzero(np_mask.any(axis=0))[0]
        if len(rows) == 0 or len(cols) == 0:
            raise RuntimeError('to_external_box: empty np_mask.')
        return Box(up=int(rows[0]), down=int(rows[-1]), left=int(cols[0]), right=int(cols[-1]))


def generate_fill_by_masks_mask(shape: Tuple[int, int], masks: Iterable[Mask],
                                mode: ElementSetOperationMode):
    if mode == ElementSetOperationMode.UNION:
        return None
    return Mask.from_masks(shape, masks, mode)


from .image import Image  # noqa: E402
from .box import Box, generate_fill_by_boxes_mask  # noqa: E402
from .polygon import Polygon, generate_fill_by_polygons_mask  # noqa: E402
from .score_map import ScoreMap  # noqa: E402
