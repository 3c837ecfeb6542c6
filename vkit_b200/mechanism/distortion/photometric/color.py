"""Colour ops (vkit/mechanism/distortion/photometric/color.py), each a short op list for the
fused per-pixel kernel.  Integer ops (mean_shift, complement, posterization, forward HSV) are
bit-exact; ops that pass through HSV->RGB or RGB<->HLS are within +-1 of cv2, which is itself
backend dependent there."""
from typing import Any, Mapping, Optional, Sequence

import attrs
import numpy as np
from numpy.random import Generator as RandomGenerator

from vkit_b200 import _native as nv
from vkit_b200.element import Image, ImageMode

from ..interface import Distortion, DistortionConfig, DistortionNopState
from .opt import OutOfBoundBehavior, channel_bits, channel_stats, make_op, run_color_ops


def _mean_shift_op(image: Image, channels, delta: int, threshold: Optional[int],
                   oob_behavior: OutOfBoundBehavior):
    return make_op(nv.OP_MEAN_SHIFT, i0=delta, i1=-1 if threshold is None else threshold,
                   i2=channel_bits(image, channels),
                   i3=1 if oob_behavior == OutOfBoundBehavior.CYCLE else 0)


@attrs.define
class MeanShiftConfig(DistortionConfig):
    delta: int
    threshold: Optional[int] = None
    channels: Optional[Sequence[int]] = None
    oob_behavior: OutOfBoundBehavior = OutOfBoundBehavior.CLIP


def mean_shift_image(config: MeanShiftConfig, state, image: Image,
                     rng: Optional[RandomGenerator]):
    if config.delta == 0:
        return image
    return run_color_ops(image, [_mean_shift_op(image, config.channels, config.delta,
                                                config.threshold, config.oob_behavior)])


mean_shift = Distortion(config_cls=MeanShiftConfig, state_cls=DistortionNopState[MeanShiftConfig],
                        func_image=mean_shift_image)


@attrs.define
class ColorShiftConfig(DistortionConfig):
    delta: int


def color_shift_image(config: ColorShiftConfig, state, image: Image,
                      rng: Optional[RandomGenerator]):
    mode = image.mode
    if mode in (ImageMode.HSV, ImageMode.HSL):
        if config.delta == 0:
            return image
        return run_color_ops(image, [_mean_shift_op(image, [0], config.delta, None,
                                                    OutOfBoundBehavior.CYCLE)])
    # RGB -> HSV_FULL, hue += delta (cyclic), -> RGB, fused in one pass (color.py:93-116)
    rgb = image.to_rgb_image()
    out = run_color_ops(rgb, [make_op(nv.OP_HUE_SHIFT_RGB, i0=config.delta)])
    return out.to_target_mode_image(mode)


color_shift = Distortion(config_cls=ColorShiftConfig,
                         state_cls=DistortionNopState[ColorShiftConfig],
                         func_image=color_shift_image)


@attrs.define
class BrightnessShiftConfig(DistortionConfig):
    delta: int
    intermediate_image_mode: ImageMode = ImageMode.HSL


def brightness_shift_image(config: BrightnessShiftConfig, state, image: Image,
                           rng: Optional[RandomGenerator]):
    mode = image.mode
    if mode in (ImageMode.HSV, ImageMode.HSL):
        if config.delta == 0:
            return image
        return run_color_ops(image, [_mean_shift_op(image, [2], config.delta, None,
                                                    OutOfBoundBehavior.CLIP)])
    assert config.intermediate_image_mode in (ImageMode.HSV, ImageMode.HSL)
    via_hsv = 1 if config.intermediate_image_mode == ImageMode.HSV else 0
    rgb = image.to_rgb_image()
    out = run_color_ops(rgb, [make_op(nv.OP_LIGHT_SHIFT_RGB, i0=config.delta, i1=via_hsv)])
    return out.to_target_mode_image(mode)


brightness_shift = Distortion(config_cls=BrightnessShiftConfig,
                              state_cls=DistortionNopState[BrightnessShiftConfig],
                              func_image=brightness_shift_image)


@attrs.define
class StdShiftConfig(DistortionConfig):
    scale: float
    channels: Optional[Sequence[int]] = None


def std_shift_image(config: StdShiftConfig, state, image: Image, rng: Optional[RandomGenerator]):
    assert config.scale > 0
    num = image.num_channels or 1
    sums, _, _ = channel_stats(image)
    count = image.height * image.width
    scale32 = np.float32(config.scale)
    if image.num_channels == 0:
        means = [np.float32(float(sums[0]) / count)]
    else:
        means = [np.float32(float(s) / count) for s in sums]
    # mat * scale - mean * (scale - 1) in float32 (color.py:165-184)
    sub = [np.float32(m) * np.float32(config.scale - 1) for m in means] + [0.0, 0.0]
    op = make_op(nv.OP_STD_SHIFT, i2=channel_bits(image, config.channels), f0=scale32, f1=sub[0],
                 f2=sub[1] if num > 1 else 0.0, f3=sub[2] if num > 2 else 0.0)
    return run_color_ops(image, [op])


std_shift = Distortion(config_cls=StdShiftConfig, state_cls=DistortionNopState[StdShiftConfig],
                       func_image=std_shift_image)


@attrs.define
class BoundaryEqualizationConfig(DistortionConfig):
    channels: Optional[Sequence[int]] = None


def boundary_equalization_image(config: BoundaryEqualizationConfig, state, image: Image,
                                rng: Optional[RandomGenerator]):
    num = image.num_channels or 1
    _, mins, maxs = channel_stats(image)
    bits = channel_bits(image, config.channels)
    mn = [0.0, 0.0, 0.0]
    sc = [0.0, 0.0, 0.0]
    active = 0
    for c in range(min(num, 3)):
        if not (bits >> c) & 1:
            continue
        delta = np.float32(maxs[c]) - np.float32(mins[c])
        if delta > 0:
            active |= 1 << c
            mn[c] = float(mins[c])
            sc[c] = np.float32(255.0) / delta
    if not active:
        return image
    op = make_op(nv.OP_BOUNDARY_EQ, i2=active, f0=mn[0], f1=mn[1], f2=mn[2], g0=sc[0], g1=sc[1],
                 g2=sc[2])
    return run_color_ops(image, [op])


boundary_equalization = Distortion(config_cls=BoundaryEqualizationConfig,
                                   state_cls=DistortionNopState[BoundaryEqualizationConfig],
                                   func_image=boundary_equalization_image)


@attrs.define
class HistogramEqualizationConfig(DistortionConfig):
    channels: Optional[Sequence[int]] = None


def equalize_hist_lut(hist: np.ndarray) -> np.ndarray:
    """cv.equalizeHist's LUT from a 256-bin histogram: lut[first] = 0, then
    saturate_cast<uchar>(cumulative * scale) with float32 scale = 255 / (total - hist[first])."""
    total = int(hist.sum())
    first = int(np.nonzero(hist)[0][0])
    if int(hist[first]) == total:
        return np.full(256, first, dtype=np.uint8)
    scale = np.float32(255.0) / np.float32(total - int(hist[first]))
    lut = np.zeros(256, dtype=np.uint8)
    cumulative = np.cumsum(hist[first + 1:].astype(np.int64))
    values = np.rint(cumulative.astype(np.float32) * scale)
    lut[first + 1:] = np.clip(values, 0, 255).astype(np.uint8)
    return lut


def histogram_equalization_image(config: HistogramEqualizationConfig, state, image: Image,
                                 rng: Optional[RandomGenerator]):
    from vkit_b200 import device as dv
    if image.mat_dtype != np.uint8:
        raise NotImplementedError('photometric ops expect uint8 images')
    channels = image.num_channels or 1
    n_pixels = image.height * image.width
    hist_dev = dv.empty((3, 256), np.uint32)
    nv.check(nv.lib().vkb_histogram_u8(dv.ptr(image.dev), n_pixels, min(channels, 3),
                                       dv.ptr(hist_dev), dv.stream_ptr()), 'vkb_histogram_u8')
    hist = dv.to_host(hist_dev)
    bits = channel_bits(image, config.channels)
    lut = np.zeros((3, 256), dtype=np.uint8)
    for c in range(min(channels, 3)):
        if (bits >> c) & 1:
            lut[c] = equalize_hist_lut(hist[c])
    lut_dev = dv.to_device(lut)
    dst = dv.empty(tuple(image.dev.shape), np.uint8)
    nv.check(nv.lib().vkb_apply_lut(dv.ptr(image.dev), dv.ptr(dst), n_pixels, channels,
                                    dv.ptr(lut_dev), bits, dv.stream_ptr()), 'vkb_apply_lut')
    return attrs.evolve(image, mat=dst)


histogram_equalization = Distortion(config_cls=HistogramEqualizationConfig,
                                    state_cls=DistortionNopState[HistogramEqualizationConfig],
                                    func_image=histogram_equalization_image)


@attrs.define
class ComplementConfig(DistortionConfig):
    threshold: Optional[int] = None
    enable_threshold_lte: bool = False
    channels: Optional[Sequence[int]] = None


def complement_image(config: ComplementConfig, state, image: Image,
                     rng: Optional[RandomGenerator]):
    if config.threshold is not None:
        assert 0 <= config.threshold <= 255
    op = make_op(nv.OP_COMPLEMENT, i1=-1 if config.threshold is None else config.threshold,
                 i2=channel_bits(image, config.channels), i3=int(config.enable_threshold_lte))
    return run_color_ops(image, [op])


complement = Distortion(config_cls=ComplementConfig,
                        state_cls=DistortionNopState[ComplementConfig],
                        func_image=complement_image)


@attrs.define
class PosterizationConfig(DistortionConfig):
    num_bits: int
    channels: Optional[Sequence[int]] = None


def posterization_image(config: PosterizationConfig, state, image: Image,
                        rng: Optional[RandomGenerator]):
    assert 0 <= config.num_bits < 8
    if config.num_bits == 0:
        return image
    op = make_op(nv.OP_POSTERIZE, i0=(0xFF >> config.num_bits) << config.num_bits,
                 i2=channel_bits(image, config.channels))
    return run_color_ops(image, [op])


posterization = Distortion(config_cls=PosterizationConfig,
                           state_cls=DistortionNopState[PosterizationConfig],
                           func_image=posterization_image)


@attrs.define
class ColorBalanceConfig(DistortionConfig):
    ratio: float


def color_balance_image(config: ColorBalanceConfig, state, image: Image,
                        rng: Optional[RandomGenerator]):
    if image.mode == ImageMode.GRAYSCALE:
        return image
    if image.mode != ImageMode.RGB:
        raise NotImplementedError('color_balance is provided for RGB / GRAYSCALE images')
    assert 0.0 <= config.ratio <= 1.0
    # (1 - ratio) * gray + ratio * mat in float32, clipped, truncated (color.py:390-391)
    op = make_op(nv.OP_COLOR_BALANCE, f0=np.float32(config.ratio),
                 f1=np.float32(1 - config.ratio))
    return run_color_ops(image, [op])


color_balance = Distortion(config_cls=ColorBalanceConfig,
                           state_cls=DistortionNopState[ColorBalanceConfig],
                           func_image=color_balance_image)


@attrs.define
class ChannelPermutationConfig(DistortionConfig):
    _rng_state: Optional[Mapping[str, Any]] = None

    @property
    def supports_rng_state(self) -> bool:
        return True

    @property
    def rng_state(self) -> Optional[Mapping[str, Any]]:
        return self._rng_state

    @rng_state.setter
    def rng_state(self, val: Mapping[str, Any]):
        self._rng_state = val


def channel_permutation_image(config: ChannelPermutationConfig, state, image: Image,
                              rng: Optional[RandomGenerator]):
    assert rng
    indices = rng.permutation(image.num_channels)
    packed = 0
    for c, idx in enumerate(indices):
        packed |= int(idx) << (4 * c)
    return run_color_ops(image, [make_op(nv.OP_PERMUTE, i0=packed)])


channel_permutation = Distortion(config_cls=ChannelPermutationConfig,
                                 state_cls=DistortionNopState[ChannelPermutationConfig],
                                 func_image=channel_permutation_image)
