"""ScoreMap: float32 HxW, optionally a probability map in [0, 1] (vkit/element/score_map.py)."""
from typing import Iterable, Optional, Tuple, Union

import attrs
import numpy as np

from .. import device as dv
from ._storage import DualStorage
from .type import ElementSetOperationMode, Shapable


@attrs.define
class ScoreMapSetItemConfig:
    value: Union['ScoreMap', np.ndarray, float] = 1.0
    keep_max_value: bool = False
    keep_min_value: bool = False


@attrs.define(frozen=True, eq=False)
class ScoreMap(DualStorage, Shapable):
    _mat: object = attrs.field(alias='mat')
    box: Optional['Box'] = None
    is_prob: bool = True
    # Results of our own kernels are convex combinations of validated inputs; re-scanning
    # them would force a device reduction + sync per op.
    skip_prob_check: bool = attrs.field(default=False, repr=False)

    _alt: object = attrs.field(default=None, init=False, repr=False)

    def __attrs_post_init__(self):
        self._adopt(self._mat)
        if self.mat_ndim != 2:
            raise RuntimeError('ndim should == 2.')
        if self.box and self.shape != self.box.shape:
            raise RuntimeError('self.shape != box.shape.')
        if self.mat_dtype != np.float32:
            raise RuntimeError('mat.dtype != np.float32')
        num_elements = self._mat.numel() if self.on_device else self._mat.size
        if self.is_prob and not self.skip_prob_check and num_elements:
            # score_map.py:104-108
            score_min = float(self._mat.min())
            score_max = float(self._mat.max())
            if score_min < 0.0 or score_max > 1.0:
                raise RuntimeError('score not in range [0.0, 1.0]')

    @classmethod
    def from_shape(cls, shape: Tuple[int, int], value: float = 0.0, is_prob: bool = True):
        height, width = shape
        if is_prob:
            assert 0.0 <= value <= 1.0
        return cls(mat=np.full((height, width), fill_value=value, dtype=np.float32),
                   is_prob=is_prob)

    @classmethod
    def from_shapable(cls, shapable: Shapable, value: float = 0.0, is_prob: bool = True):
        return cls.from_shape(shape=shapable.shape, value=value, is_prob=is_prob)

    @property
    def equivalent_box(self):
        return self.box or Box.from_shapable(self)

    def copy(self):
        return attrs.evolve(self, mat=self._clone_storage(), skip_prob_check=True)

    def assign_mat(self, mat: np.ndarray):
        self._adopt(mat)

    def fill_by_boxes(self, boxes: Iterable['Box'], value=1.0,
                      mode: ElementSetOperationMode = ElementSetOperationMode.UNION,
                      keep_max_value: bool = False, keep_min_value: bool = False):
        boxes = list(boxes)
        boxes_mask = generate_fill_by_boxes_mask(self.shape, boxes, mode)
        if boxes_mask is None:
            for box in boxes:
                box.fill_score_map(self, value, keep_max_value=keep_max_value,
                                   keep_min_value=keep_min_value)
        else:
            boxes_mask.fill_score_map(self, value, keep_max_value=keep_max_value,
                                      keep_min_value=keep_min_value)

    def fill_by_polygons(self, polygons: Iterable['Polygon'], value=1.0,
                         mode: ElementSetOperationMode = ElementSetOperationMode.UNION,
                         keep_max_value: bool = False, keep_min_value: bool = False):
        polygons = list(polygons)
        polygons_mask = generate_fill_by_polygons_mask(self.shape, polygons, mode)
        if polygons_mask is None:
            for polygon in polygons:
                polygon.fill_score_map(self, value, keep_max_value=keep_max_value,
                                       keep_min_value=keep_min_value)
        else:
            polygons_mask.fill_score_map(self, value, keep_max_value=keep_max_value,
                                         keep_min_value=keep_min_value)

    def __setitem__(self, element, config):
        if isinstance(config, ScoreMapSetItemConfig):
            value, keep_max, keep_min = config.value, config.keep_max_value, config.keep_min_value
        else:
            value, keep_max, keep_min = config, False, False
        element.fill_score_map(score_map=self, value=value, keep_max_value=keep_max,
                               keep_min_value=keep_min)

    def __getitem__(self, element):
        return element.extract_score_map(self)

    def to_shifted_score_map(self, offset_y: int = 0, offset_x: int = 0):
        assert self.box
        return attrs.evolve(self, box=self.box.to_shifted_box(offset_y=offset_y, offset_x=offset_x),
                            skip_prob_check=True)

    # cv.INTER_NEAREST / LINEAR / CUBIC / AREA / LANCZOS4 / LINEAR_EXACT (= LINEAR for float) / NEAREST_EXACT
    _CV_INTER = {0: 0, 1: 1, 2: 2, 3: 3, 4: 4, 5: 5, 6: 6}

    def to_conducted_resized_score_map(self, shapable_or_shape, resized_height: Optional[int] = None,
                                       resized_width: Optional[int] = None,
                                       cv_resize_interpolation: int = 2):
        """element/score_map.py:594-614: resize the box and the scores it holds together."""
        assert self.box
        resized_box = self.box.to_conducted_resized_box(
            shapable_or_shape=shapable_or_shape, resized_height=resized_height,
            resized_width=resized_width)
        resized = self.to_box_detached().to_resized_score_map(
            resized_height=resized_box.height, resized_width=resized_box.width,
            cv_resize_interpolation=cv_resize_interpolation)
        return resized.to_box_attached(resized_box)

    # the reference spells this method `to_conducted_resized_polygon` (element/score_map.py:595)
    to_conducted_resized_polygon = to_conducted_resized_score_map

    def to_resized_score_map(self, resized_height: Optional[int] = None,
                             resized_width: Optional[int] = None, cv_resize_interpolation: int = 2,
                             post_scale: float = 1.0):
        """element/score_map.py:616-637: cv.resize of the float32 map (INTER_CUBIC by default),
        clipped to [0, 1] when it is a probability map -- the clip is fused into the kernel.
        LINEAR / LINEAR_EXACT / CUBIC follow the wheel's default backend (Intel IPP: coordinates and
        taps in double, within 5e-7 of the wheel); the other codes restate cv2's own float32 path
        bit for bit (vkb_resize_f32).  `post_scale` (not in the reference
        signature) multiplies the resized float32 values in the same kernel -- page_resizing scales
        its height maps by the resize ratio right after resizing them (page_resizing.py:160-161)."""
        from .. import _native
        from .opt import generate_shape_and_resized_shape
        assert not self.box
        _, _, resized_height, resized_width = generate_shape_and_resized_shape(
            self, resized_height, resized_width)
        if cv_resize_interpolation not in self._CV_INTER:
            raise NotImplementedError(
                'to_resized_score_map: cv.INTER_NEAREST / LINEAR / CUBIC / AREA / LANCZOS4 / '
                'LINEAR_EXACT / NEAREST_EXACT have device kernels')
        src = self.dev
        dst = dv.empty((resized_height, resized_width), np.float32)
        _native.check(_native.lib().vkb_resize_f32_scaled(
            dv.ptr(src), self.height, self.width, dv.ptr(dst), resized_height, resized_width,
            self._CV_INTER[cv_resize_interpolation], int(self.is_prob), float(post_scale),
            dv.stream_ptr()), 'vkb_resize_f32_scaled')
        return attrs.evolve(self, mat=dst, skip_prob_check=True)

    def to_cropped_score_map(self, up=None, down=None, left=None, right=None):
        assert not self.box
        up = up or 0
        down = down or self.height - 1
        left = left or 0
        right = right or self.width - 1
        return attrs.evolve(self, mat=self._crop_storage(up, down, left, right),
                            skip_prob_check=True)

    def to_box_attached(self, box: 'Box'):
        assert self.height == box.height
        assert self.width == box.width
        return attrs.evolve(self, box=box, skip_prob_check=True)

    def to_box_detached(self):
        assert self.box
        return attrs.evolve(self, box=None, skip_prob_check=True)

    def fill_np_array(self, mat, value, keep_max_value=False, keep_min_value=False):
        self.equivalent_box.fill_np_array(mat=mat, value=value, alpha=self,
                                          keep_max_value=keep_max_value,
                                          keep_min_value=keep_min_value)

    def fill_image(self, image: 'Image', value):
        # score map = mask (alpha > 0) + alpha in one (score_map.py:678-687, box.py:329-331)
        self.equivalent_box.fill_image(image=image, value=value, alpha=self)

    def to_mask(self, threshold: float = 0.0):
        if self.on_device:
            return Mask(mat=(self.dev > threshold).to(dv.torch().uint8), box=self.box)
        return Mask(mat=(self.mat > threshold).astype(np.uint8), box=self.box)


def generate_fill_by_score_maps_mask(shape: Tuple[int, int], score_maps: Iterable[ScoreMap],
                                     mode: ElementSetOperationMode):
    if mode == ElementSetOperationMode.UNION:
        return None
    return Mask.from_score_maps(shape, score_maps, mode)


from .image import Image  # noqa: E402
from .point import Point  # noqa: E402
from .box import Box, generate_fill_by_boxes_mask  # noqa: E402
from .mask import Mask, generate_fill_by_masks_mask  # noqa: E402
from .polygon import Polygon, generate_fill_by_polygons_mask  # noqa: E402
