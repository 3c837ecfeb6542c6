"""Helpers of the photometric ops (vkit/mechanism/distortion/photometric/opt.py) and the
launcher of the fused per-pixel op list (`vkb_color_ops`)."""
import ctypes
from enum import Enum, unique
from typing import Optional, Sequence

import attrs
import numpy as np

from vkit_b200 import _native as nv
from vkit_b200 import device as dv
from vkit_b200.element import Image, ImageMode


@unique
class OutOfBoundBehavior(Enum):
    CLIP = 'clip'
    CYCLE = 'cycle'


def to_rgb_image(image: Image, mode: ImageMode):
    if mode not in (ImageMode.GRAYSCALE, ImageMode.RGB):
        image = image.to_rgb_image()
    return image


def to_original_image(image: Image, mode: ImageMode):
    if mode not in (ImageMode.GRAYSCALE, ImageMode.RGB):
        image = image.to_target_mode_image(mode)
    return image


def channel_bits(image: Image, channels: Optional[Sequence[int]]):
    num = image.num_channels or 1
    if not channels:
        return (1 << num) - 1
    bits = 0
    for channel in channels:
        if channel < 0 or channel >= num:
            raise IndexError(f'channel {channel} out of range')
        bits |= 1 << channel
    return bits


def make_op(kind, i0=0, i1=0, i2=0, i3=0, f0=0.0, f1=0.0, f2=0.0, f3=0.0, g0=0.0, g1=0.0, g2=0.0):
    return nv.ColorOp(kind, int(i0), int(i1), int(i2), int(i3), float(f0), float(f1), float(f2),
                      float(f3), float(g0), float(g1), float(g2))


def run_color_ops(image: Image, ops, keep_mode: bool = True) -> Image:
    """Apply an op list to a uint8 image in one pass; returns a new device-backed Image."""
    if image.mat_dtype != np.uint8:
        raise NotImplementedError('photometric ops expect uint8 images')
    channels = image.num_channels or 1
    src = image.dev
    dst = dv.empty(tuple(src.shape), np.uint8)
    arr = (nv.ColorOp * len(ops))(*ops)
    nv.check(nv.lib().vkb_color_ops(dv.ptr(src), dv.ptr(dst), image.height * image.width, channels,
                                    arr, len(ops), dv.stream_ptr()), 'vkb_color_ops')
    if keep_mode:
        return attrs.evolve(image, mat=dst)
    return Image(mat=dst)


def channel_stats(image: Image):
    """(sums[3] uint64, mins[3], maxs[3]) of a uint8 image, computed on the device."""
    channels = image.num_channels or 1
    out = dv.empty((48,), np.uint8)
    nv.check(nv.lib().vkb_channel_stats(dv.ptr(image.dev), image.height * image.width, channels,
                                        dv.ptr(out), dv.stream_ptr()), 'vkb_channel_stats')
    raw = dv.to_host(out).tobytes()
    sums = np.frombuffer(raw[:24], dtype=np.uint64)
    mins = np.frombuffer(raw[24:36], dtype=np.uint32)
    maxs = np.frombuffer(raw[36:48], dtype=np.uint32)
    return sums[:channels], mins[:channels], maxs[:channels]
