"""Streak ops (vkit/mechanism/distortion/photometric/streak.py).

line_streak evaluates its periodic masks analytically per pixel (no mask buffer);
rectangle_streak rasterises the host-computed bar lists into two device masks; both blend
`color` with `alpha`, vertical mask first, horizontal second -- crossings get alpha twice,
like the reference (streak.py:96-98)."""
import ctypes
from typing import List, Optional, Tuple

import attrs
import numpy as np
from numpy.random import Generator as RandomGenerator

from vkit_b200 import _native as nv
from vkit_b200 import device as dv
from vkit_b200.element import Box, Image

from ..interface import Distortion, DistortionConfig, DistortionNopState


def _color_array(image: Image, color):
    channels = image.num_channels or 1
    if isinstance(color, (tuple, list)):
        if channels > 1 and len(color) != channels:
            raise RuntimeError('value is tuple but len(value) != num_channels.')
        values = list(color)
    else:
        values = [color] * channels
    values = list(np.asarray(values).astype(np.uint8).astype(np.float32)) + [0.0] * 4
    return (ctypes.c_float * 4)(*[float(v) for v in values[:4]])


def _check_alpha(alpha: float):
    alpha = float(alpha)
    if alpha < 0.0 or alpha > 1.0:
        raise RuntimeError(f'alpha={alpha} is invalid.')
    return alpha


@attrs.define
class LineStreakConfig(DistortionConfig):
    thickness: int = 1
    gap: int = 4
    dash_thickness: int = 0
    dash_gap: int = 0
    color: Tuple[int, int, int] = (0, 0, 0)
    alpha: float = 1.0
    enable_vert: bool = True
    enable_hori: bool = True


def line_streak_image(config: LineStreakConfig, state, image: Image,
                      rng: Optional[RandomGenerator]):
    image = image.copy()
    alpha = _check_alpha(config.alpha)
    if alpha == 0.0 or not (config.enable_vert or config.enable_hori):
        return image
    dst = image.dev
    nv.check(nv.lib().vkb_streak_line(
        dv.ptr(dst), image.height, image.width, image.num_channels or 1, config.thickness,
        config.gap, config.dash_thickness, config.dash_gap, int(config.enable_vert),
        int(config.enable_hori), _color_array(image, config.color), alpha, dv.stream_ptr()),
        'vkb_streak_line')
    image._after_device_write()
    return image


line_streak = Distortion(config_cls=LineStreakConfig,
                         state_cls=DistortionNopState[LineStreakConfig],
                         func_image=line_streak_image)


def generate_centered_boxes(height: int, width: int, aspect_ratio: float, short_side_min: int,
                            short_side_step: int):
    # streak.py:109-145
    center_y = height // 2
    center_x = width // 2
    boxes: List[Box] = []
    idx = 0
    while True:
        short_side = short_side_min + idx * short_side_step
        if aspect_ratio >= 1:
            height_min = short_side
            width_min = round(height_min * aspect_ratio)
        elif 0 < aspect_ratio < 1:
            width_min = short_side
            height_min = round(width_min / aspect_ratio)
        else:
            raise NotImplementedError()
        up = center_y - height_min // 2
        down = up + height_min - 1
        left = center_x - width_min // 2
        right = left + width_min - 1
        if (0 <= up and down < height) or (0 <= left and right < width):
            boxes.append(Box(up=up, down=down, left=left, right=right))
            idx += 1
        else:
            break
    return boxes


@attrs.define
class RectangleStreakConfig(DistortionConfig):
    thickness: int = 1
    aspect_ratio: Optional[float] = None
    dash_thickness: int = 0
    dash_gap: int = 0
    short_side_min: int = 10
    short_side_step: int = 10
    color: Tuple[int, int, int] = (0, 0, 0)
    alpha: float = 1.0


def _rectangle_bars(boxes, thickness: int, height: int, width: int):
    """Bar lists of the concentric rectangles (streak.py:170-236)."""
    vert, hori = [], []
    for box in boxes:
        inner_up = box.down - thickness + 1
        inner_down = box.up + thickness - 1
        inner_left = box.right - thickness + 1
        inner_right = box.left + thickness - 1
        bar_up, bar_down = max(0, box.up), min(height - 1, box.down)
        if 0 <= inner_right < width and bar_up <= bar_down:
            vert.append((bar_up, bar_down, max(0, box.left), inner_right))
        if 0 <= inner_left < width and bar_up <= bar_down:
            vert.append((bar_up, bar_down, inner_left, min(width - 1, box.right)))
        bar_left, bar_right = max(0, inner_right + 1), min(width - 1, inner_left - 1)
        if 0 <= inner_down < height and bar_left <= bar_right:
            hori.append((max(0, box.up), inner_down, bar_left, bar_right))
        if 0 <= inner_up < height and bar_left <= bar_right:
            hori.append((inner_up, min(height - 1, box.down), bar_left, bar_right))
    return vert, hori


def _rects_mask(rects, height: int, width: int):
    mask = dv.zeros((height, width), np.uint8)
    if rects:
        arr = np.asarray(rects, dtype=np.int32).reshape(-1, 4)
        rects_dev = dv.to_device(arr)
        nv.check(nv.lib().vkb_fill_rects(dv.ptr(mask), height, width, dv.ptr(rects_dev),
                                         int(arr.shape[0]), dv.stream_ptr()), 'vkb_fill_rects')
    return mask


def rectangle_streak_image(config: RectangleStreakConfig, state, image: Image,
                           rng: Optional[RandomGenerator]):
    aspect_ratio = config.aspect_ratio
    if aspect_ratio is None:
        aspect_ratio = image.width / image.height
    boxes = generate_centered_boxes(image.height, image.width, aspect_ratio,
                                    config.short_side_min, config.short_side_step)
    vert, hori = _rectangle_bars(boxes, config.thickness, image.height, image.width)
    image = image.copy()
    alpha = _check_alpha(config.alpha)
    if alpha == 0.0:
        return image
    mask_vert = _rects_mask(vert, image.height, image.width)
    mask_hori = _rects_mask(hori, image.height, image.width)
    nv.check(nv.lib().vkb_streak_masks(
        dv.ptr(image.dev), image.height, image.width, image.num_channels or 1, dv.ptr(mask_vert),
        dv.ptr(mask_hori), config.dash_thickness, config.dash_gap,
        _color_array(image, config.color), alpha, dv.stream_ptr()), 'vkb_streak_masks')
    image._after_device_write()
    return image


rectangle_streak = Distortion(config_cls=RectangleStreakConfig,
                              state_cls=DistortionNopState[RectangleStreakConfig],
                              func_image=rectangle_streak_image)


@attrs.define
class EllipseStreakConfig(DistortionConfig):
    thickness: int = 1
    aspect_ratio: Optional[float] = None
    short_side_min: int = 10
    short_side_step: int = 10
    color: Tuple[int, int, int] = (0, 0, 0)
    alpha: float = 1.0


def ellipse_streak_image(config: EllipseStreakConfig, state, image: Image,
                         rng: Optional[RandomGenerator]):
    # streak.py:292-330.  (Under cv2 >= 4.13 the reference itself raises inside cv.ellipse because
    # Mask.mat is read-only; the fixtures come from the reference with that array made writable.)
    aspect_ratio = config.aspect_ratio
    if aspect_ratio is None:
        aspect_ratio = image.width / image.height
    boxes = generate_centered_boxes(image.height, image.width, aspect_ratio,
                                    config.short_side_min, config.short_side_step)
    if config.thickness < 1:
        raise NotImplementedError('ellipse_streak: filled ellipses (thickness < 1) are not provided')
    image = image.copy()
    alpha = _check_alpha(config.alpha)
    if alpha == 0.0 or not boxes:
        return image
    height, width = image.height, image.width
    ellipses = np.asarray([(width // 2, height // 2, box.width // 2, box.height // 2)
                           for box in boxes], dtype=np.int32)
    mask = dv.zeros((height, width), np.uint8)
    workspace = dv.empty((int(ellipses.shape[0]) * 74 * 40,), np.uint8)
    nv.check(nv.lib().vkb_draw_ellipses(
        dv.ptr(mask), height, width, ellipses.ctypes.data_as(ctypes.c_void_p),
        int(ellipses.shape[0]), int(config.thickness), dv.ptr(workspace),
        int(workspace.numel()), dv.stream_ptr()), 'vkb_draw_ellipses')
    # mask.fill_image(image, color, alpha): one blend over the drawn pixels
    no_mask = ctypes.c_void_p(0)
    nv.check(nv.lib().vkb_streak_masks(
        dv.ptr(image.dev), height, width, image.num_channels or 1, dv.ptr(mask), no_mask, 0, 0,
        _color_array(image, config.color), alpha, dv.stream_ptr()), 'vkb_streak_masks')
    image._after_device_write()
    return image


ellipse_streak = Distortion(config_cls=EllipseStreakConfig,
                            state_cls=DistortionNopState[EllipseStreakConfig],
                            func_image=ellipse_streak_image)
