"""Image: uint8 HxW / HxWx3 / HxWx4 (or float32 GCN modes) plus a colour mode
(vkit/element/image.py).  Mode conversions run as device kernels (`vkb_cvt_color`) with the
integer / float32 arithmetic of the cv.cvtColor codes the reference uses (image.py:183-207).
"""
from collections import abc
from enum import Enum, unique
from typing import Iterable, Optional, Tuple, Union

import attrs
import numpy as np

from .. import _native
from .. import device as dv
from ._storage import DualStorage
from .type import ElementSetOperationMode, Shapable


@unique
class ImageMode(Enum):
    RGB = 'rgb'
    RGB_GCN = 'rgb_gcn'
    RGBA = 'rgba'
    HSV = 'hsv'
    HSV_GCN = 'hsv_gcn'
    HSL = 'hsl'
    HSL_GCN = 'hsl_gcn'
    GRAYSCALE = 'grayscale'
    GRAYSCALE_GCN = 'grayscale_gcn'
    NONE = 'none'

    def to_ndim(self):
        if self in (ImageMode.GRAYSCALE, ImageMode.GRAYSCALE_GCN):
            return 2
        if self is ImageMode.NONE:
            raise NotImplementedError()
        return 3

    def to_dtype(self):
        if self in _GCN_TO_NON_GCN:
            return np.float32
        if self is ImageMode.NONE:
            raise NotImplementedError()
        return np.uint8

    def to_num_channels(self):
        if self is ImageMode.RGBA:
            return 4
        if self in (ImageMode.GRAYSCALE, ImageMode.GRAYSCALE_GCN):
            return None
        if self is ImageMode.NONE:
            raise NotImplementedError
        return 3

    def supports_gcn_mode(self):
        return self not in _NON_GCN_TO_GCN

    def to_gcn_mode(self):
        if not self.supports_gcn_mode():
            raise RuntimeError(f'image_mode={self} not supported.')
        return _NON_GCN_TO_GCN[self]

    def in_gcn_mode(self):
        return self in _GCN_TO_NON_GCN

    def to_non_gcn_mode(self):
        if not self.in_gcn_mode():
            raise RuntimeError(f'image_mode={self} not in gcn mode.')
        return _GCN_TO_NON_GCN[self]


_NON_GCN_TO_GCN = {
    ImageMode.RGB: ImageMode.RGB_GCN,
    ImageMode.HSV: ImageMode.HSV_GCN,
    ImageMode.HSL: ImageMode.HSL_GCN,
    ImageMode.GRAYSCALE: ImageMode.GRAYSCALE_GCN,
}
_GCN_TO_NON_GCN = {val: key for key, val in _NON_GCN_TO_GCN.items()}

# (src mode, dst mode) -> (native conversion code, dst channels [0 = 2-D])
_CVT = {
    (ImageMode.RGB, ImageMode.HSV): (_native.CVT_RGB2HSV, 3),
    (ImageMode.HSV, ImageMode.RGB): (_native.CVT_HSV2RGB, 3),
    (ImageMode.RGB, ImageMode.HSL): (_native.CVT_RGB2HSL, 3),
    (ImageMode.HSL, ImageMode.RGB): (_native.CVT_HSL2RGB, 3),
    (ImageMode.RGB, ImageMode.GRAYSCALE): (_native.CVT_RGB2GRAY, 0),
    (ImageMode.GRAYSCALE, ImageMode.RGB): (_native.CVT_GRAY2RGB, 3),
    (ImageMode.RGBA, ImageMode.RGB): (_native.CVT_RGBA2RGB, 3),
    (ImageMode.RGB, ImageMode.RGBA): (_native.CVT_RGB2RGBA, 4),
    (ImageMode.GRAYSCALE, ImageMode.RGBA): (_native.CVT_GRAY2RGBA, 4),
    (ImageMode.RGBA, ImageMode.GRAYSCALE): (_native.CVT_RGBA2GRAY, 0),
}


def convert_color(src_dev, code: int, dst_channels: int):
    """One cv.cvtColor-equivalent pass on the device; returns a new CUDA tensor."""
    height, width = int(src_dev.shape[0]), int(src_dev.shape[1])
    shape = (height, width) if dst_channels == 0 else (height, width, dst_channels)
    dst = dv.empty(shape, np.uint8)
    _native.check(
        _native.lib().vkb_cvt_color(dv.ptr(src_dev), dv.ptr(dst), height * width, code,
                                    dv.stream_ptr()), 'vkb_cvt_color')
    return dst


@attrs.define
class ImageSetItemConfig:
    value: Union['Image', np.ndarray, Tuple[int, ...], int]
    alpha: Union[np.ndarray, float] = 1.0


@attrs.define(frozen=True, eq=False)
class Image(DualStorage, Shapable):
    _mat: object = attrs.field(alias='mat')
    mode: ImageMode = ImageMode.NONE
    box: Optional['Box'] = None

    _alt: object = attrs.field(default=None, init=False, repr=False)

    def __attrs_post_init__(self):
        self._adopt(self._mat)
        dtype, ndim, shape = self.mat_dtype, self.mat_ndim, self.mat_shape
        if self.mode != ImageMode.NONE:
            assert self.mode.to_dtype() == dtype
            assert self.mode.to_ndim() == ndim
        else:
            # infer the mode from the array (image.py:228-252)
            if dtype == np.float32:
                raise NotImplementedError('mode is None and mat.dtype == np.float32.')
            if dtype != np.uint8:
                raise NotImplementedError(f'Invalid mat.dtype={dtype}.')
            if ndim == 2:
                mode = ImageMode.GRAYSCALE
            elif ndim == 3 and shape[2] == 4:
                mode = ImageMode.RGBA
            elif ndim == 3 and shape[2] == 3:
                mode = ImageMode.RGB
            elif ndim == 3:
                raise NotImplementedError(f'Invalid num_channels={shape[2]}.')
            else:
                raise NotImplementedError(f'mat.ndim={ndim} not supported.')
            object.__setattr__(self, 'mode', mode)
        if self.box and self.shape != self.box.shape:
            raise RuntimeError('self.shape != box.shape.')

    @classmethod
    def from_shape(cls, shape: Tuple[int, int], num_channels: int = 3,
                   value: Union[Tuple[int, ...], int] = 255):
        height, width = shape
        if num_channels == 0:
            mat_shape = (height, width)
        else:
            assert num_channels > 0
            if isinstance(value, tuple):
                assert len(value) == num_channels
            mat_shape = (height, width, num_channels)
        return cls(mat=np.full(mat_shape, fill_value=value, dtype=np.uint8))

    @classmethod
    def from_shapable(cls, shapable: Shapable, num_channels: int = 3,
                      value: Union[Tuple[int, ...], int] = 255):
        return cls.from_shape(shape=shapable.shape, num_channels=num_channels, value=value)

    @property
    def num_channels(self):
        return 0 if self.mat_ndim == 2 else self.mat_shape[2]

    def copy(self):
        return attrs.evolve(self, mat=self._clone_storage())

    def assign_mat(self, mat):
        self._adopt(mat)

    # ---- fills -------------------------------------------------------------------------
    def fill_by_boxes(self, boxes: Iterable['Box'], value, alpha=1.0,
                      mode: ElementSetOperationMode = ElementSetOperationMode.UNION):
        boxes = list(boxes)
        boxes_mask = generate_fill_by_boxes_mask(self.shape, boxes, mode)
        if boxes_mask is None:
            for box in boxes:
                box.fill_image(image=self, value=value, alpha=alpha)
        else:
            boxes_mask.fill_image(image=self, value=value, alpha=alpha)

    def fill_by_polygons(self, polygons: Iterable['Polygon'], value, alpha=1.0,
                         mode: ElementSetOperationMode = ElementSetOperationMode.UNION):
        polygons = list(polygons)
        polygons_mask = generate_fill_by_polygons_mask(self.shape, polygons, mode)
        if polygons_mask is None:
            for polygon in polygons:
                polygon.fill_image(image=self, value=value, alpha=alpha)
        else:
            polygons_mask.fill_image(image=self, value=value, alpha=alpha)

    def fill_by_masks(self, masks: Iterable['Mask'], value, alpha=1.0,
                      mode: ElementSetOperationMode = ElementSetOperationMode.UNION):
        masks = list(masks)
        masks_mask = generate_fill_by_masks_mask(self.shape, masks, mode)
        if masks_mask is None:
            for mask in masks:
                mask.fill_image(image=self, value=value, alpha=alpha)
        else:
            masks_mask.fill_image(image=self, value=value, alpha=alpha)

    def fill_by_score_maps(self, score_maps: Iterable['ScoreMap'], value,
                           mode: ElementSetOperationMode = ElementSetOperationMode.UNION):
        score_maps = list(score_maps)
        if mode != ElementSetOperationMode.UNION:
            raise NotImplementedError('only UNION is provided for score-map fills')
        for score_map in score_maps:
            score_map.fill_image(image=self, value=value)

    def __setitem__(self, element, config):
        if isinstance(config, ImageSetItemConfig):
            value, alpha = config.value, config.alpha
        else:
            value, alpha = config, 1.0
        if isinstance(value, tuple):
            assert value and isinstance(value[0], int)
        else:
            assert isinstance(value, (Image, np.ndarray)) or not isinstance(value, abc.Iterable)
        if isinstance(element, ScoreMap):
            element.fill_image(image=self, value=value)
        else:
            element.fill_image(image=self, value=value, alpha=alpha)

    def __getitem__(self, element):
        return element.extract_image(self)

    def to_box_attached(self, box: 'Box'):
        assert self.height == box.height
        assert self.width == box.width
        return attrs.evolve(self, box=box)

    def to_box_detached(self):
        assert self.box
        return attrs.evolve(self, box=None)

    # ---- mode conversion (image.py:771-814) --------------------------------------------
    def to_target_mode_image(self, target_mode: ImageMode):
        if target_mode == self.mode:
            return self
        if self.mode.in_gcn_mode() or target_mode.in_gcn_mode():
            raise NotImplementedError('GCN modes are outside the distortion path')
        if (self.mode, target_mode) in _CVT:
            code, channels = _CVT[(self.mode, target_mode)]
            return Image(mat=convert_color(self.dev, code, channels), mode=target_mode)
        # two hops through RGB, like the reference
        code, channels = _CVT[(self.mode, ImageMode.RGB)]
        rgb = convert_color(self.dev, code, channels)
        code, channels = _CVT[(ImageMode.RGB, target_mode)]
        return Image(mat=convert_color(rgb, code, channels), mode=target_mode)

    def to_grayscale_image(self):
        return self.to_target_mode_image(ImageMode.GRAYSCALE)

    def to_rgb_image(self):
        return self.to_target_mode_image(ImageMode.RGB)

    def to_rgba_image(self):
        return self.to_target_mode_image(ImageMode.RGBA)

    def to_hsv_image(self):
        return self.to_target_mode_image(ImageMode.HSV)

    def to_hsl_image(self):
        return self.to_target_mode_image(ImageMode.HSL)

    def to_shifted_image(self, offset_y: int = 0, offset_x: int = 0):
        assert self.box
        return attrs.evolve(self, box=self.box.to_shifted_box(offset_y=offset_y,
                                                              offset_x=offset_x))

    def to_cropped_image(self, up=None, down=None, left=None, right=None):
        assert not self.box
        up = up or 0
        down = down or self.height - 1
        left = left or 0
        right = right or self.width - 1
        return attrs.evolve(self, mat=self._crop_storage(up, down, left, right))


    # cv2 interpolation codes (cv.INTER_NEAREST / INTER_LINEAR / INTER_CUBIC)
    _CV_INTER = {0: _native.INTER_NEAREST, 1: _native.INTER_LINEAR, 2: _native.INTER_CUBIC,
                 3: _native.INTER_AREA, 4: _native.INTER_LANCZOS4, 5: _native.INTER_LINEAR_EXACT, 6: _native.INTER_NEAREST_EXACT}

    def to_resized_image(self, resized_height: Optional[int] = None,
                         resized_width: Optional[int] = None, cv_resize_interpolation: int = 2):
        """element/image.py:836-852.  The device resize reproduces cv.resize bit for bit for
        INTER_NEAREST (0), INTER_LINEAR (1), INTER_AREA (3), INTER_LANCZOS4 (4), INTER_LINEAR_EXACT (5) and
        INTER_NEAREST_EXACT (6).
        INTER_CUBIC (2, the reference's default) follows the cv2 wheel: Intel IPP's cubic for
        sources of at least 4 x 4 pixels (the float64 bicubic; +-1 on < 3e-4 of the pixels at near
        ties), cv2's own path below that, bit for bit."""
        from .opt import generate_shape_and_resized_shape
        _, _, resized_height, resized_width = generate_shape_and_resized_shape(
            self, resized_height, resized_width)
        if cv_resize_interpolation not in self._CV_INTER:
            raise NotImplementedError(
                'to_resized_image: cv.INTER_NEAREST / LINEAR / CUBIC / AREA / LANCZOS4 / LINEAR_EXACT '
                '/ NEAREST_EXACT have device kernels')
        if self.mat_dtype != np.uint8:
            raise NotImplementedError('to_resized_image is provided for uint8 images')
        src = self.dev
        channels = self.num_channels or 1
        shape = (resized_height, resized_width) + ((channels,) if self.num_channels else ())
        dst = dv.empty(shape, np.uint8)
        _native.check(_native.lib().vkb_resize_u8(
            dv.ptr(src), self.height, self.width, dv.ptr(dst), resized_height, resized_width,
            channels, self._CV_INTER[cv_resize_interpolation], dv.stream_ptr()), 'vkb_resize_u8')
        return attrs.evolve(self, mat=dst)

    def to_conducted_resized_image(self, shapable_or_shape, resized_height: Optional[int] = None,
                                   resized_width: Optional[int] = None,
                                   cv_resize_interpolation: int = 2):
        """element/image.py:854-873: resize the attached box and the pixels it holds together."""
        assert self.box
        resized_box = self.box.to_conducted_resized_box(
            shapable_or_shape=shapable_or_shape, resized_height=resized_height,
            resized_width=resized_width)
        resized = self.to_box_detached().to_resized_image(
            resized_height=resized_box.height, resized_width=resized_box.width,
            cv_resize_interpolation=cv_resize_interpolation)
        return resized.to_box_attached(resized_box)


from .box import Box, generate_fill_by_boxes_mask  # noqa: E402
from .polygon import Polygon, generate_fill_by_polygons_mask  # noqa: E402
from .mask import Mask, generate_fill_by_masks_mask  # noqa: E402
from .score_map import ScoreMap, generate_fill_by_score_maps_mask  # noqa: E402
