"""Effect ops (vkit/mechanism/distortion/photometric/effect.py): jpeg_quality, pixelation, fog.

jpeg_quality: the JPEG round trip without entropy coding, libjpeg's integer arithmetic (csrc/jpeg.cu).
pixelation: cv.resize INTER_LINEAR down + INTER_NEAREST up, both bit exact on the device.
fog: the diamond-square field is drawn on the host from the caller's NumPy generator (it is the
random field the reference would draw, a few vector operations per level), the blend runs on the
device."""
import ctypes
from typing import Any, Mapping, Optional, Tuple

import attrs
import numpy as np
from numpy.random import Generator as RandomGenerator

from vkit_b200 import _native as nv
from vkit_b200 import device as dv
from vkit_b200.element import Image, ImageMode

from ..interface import Distortion, DistortionConfig, DistortionNopState
from .opt import to_original_image, to_rgb_image


@attrs.define
class JpegQualityConfig(DistortionConfig):
    quality: int


def jpeg_quality_image(config: JpegQualityConfig, state, image: Image,
                       rng: Optional[RandomGenerator]):
    # effect.py:35-48: cv.imencode('.jpeg', mat, quality) + cv.imdecode.  The round trip runs on
    # the device without the (lossless) entropy coding, in libjpeg's integer arithmetic.
    mode = image.mode
    image = to_rgb_image(image, mode)
    assert 0 <= config.quality <= 100
    if image.mat_dtype != np.uint8:
        raise NotImplementedError('jpeg_quality is provided for uint8 images')
    channels = image.num_channels or 1
    height, width = image.height, image.width
    padded = ((height + 15) // 16 * 16) * ((width + 15) // 16 * 16)
    planes = dv.empty((padded + padded // 2,), np.uint8)
    shape = (height, width) if image.num_channels == 0 else (height, width, channels)
    dst = dv.empty(shape, np.uint8)
    nv.check(nv.lib().vkb_jpeg_round_trip_u8(dv.ptr(image.dev), dv.ptr(dst), height, width, channels,
                                             int(config.quality), dv.ptr(planes),
                                             int(planes.numel()), dv.stream_ptr()),
             'vkb_jpeg_round_trip_u8')
    image = attrs.evolve(image, mat=dst)
    return to_original_image(image, mode)


jpeg_quality = Distortion(config_cls=JpegQualityConfig,
                          state_cls=DistortionNopState[JpegQualityConfig],
                          func_image=jpeg_quality_image)


@attrs.define
class PixelationConfig(DistortionConfig):
    ratio: float


def resize_device(image: Image, resized_height: int, resized_width: int, interpolation: int) -> Image:
    """cv.resize(image.mat, (w, h), interpolation) for uint8 images, NEAREST / LINEAR."""
    if image.mat_dtype != np.uint8:
        raise NotImplementedError('resize is provided for uint8 images')
    src = image.dev
    channels = image.num_channels or 1
    shape = (resized_height, resized_width) + ((channels,) if image.num_channels else ())
    dst = dv.empty(shape, np.uint8)
    nv.check(nv.lib().vkb_resize_u8(dv.ptr(src), image.height, image.width, dv.ptr(dst),
                                    resized_height, resized_width, channels, interpolation,
                                    dv.stream_ptr()), 'vkb_resize_u8')
    return attrs.evolve(image, mat=dst)


def pixelation_image(config: PixelationConfig, state, image: Image,
                     rng: Optional[RandomGenerator]):
    # effect.py:58-79
    assert 0 < config.ratio < 1
    small = resize_device(image, round(image.height * config.ratio),
                          round(image.width * config.ratio), nv.INTER_LINEAR)
    return resize_device(small, image.height, image.width, nv.INTER_NEAREST)


pixelation = Distortion(config_cls=PixelationConfig,
                        state_cls=DistortionNopState[PixelationConfig],
                        func_image=pixelation_image)


def generate_diamond_square_mask(shape: Tuple[int, int], roughness: float, rng: RandomGenerator):
    """Diamond-square plasma field cropped to `shape` (effect.py:89-146); consumes `rng` exactly
    like the reference: 4 corner draws, then per level the diamond draw, the two square draws,
    and finally the crop offsets.  Every displacement is `(1 - weight) * sums / 4 + weight * U(0, 1)`
    with the reference's dtype sequence (float32 corner sums, float64 centres and draws, float32
    field); the neighbour sums are slices instead of np.roll copies, same additions.  One quirk is
    kept on purpose: for the column midpoints the reference closes the wrapped sums with the first
    ROW of `left_right` (`left_right[0].reshape(-1, 1)`, effect.py:131-133), not its first column."""
    assert 0.0 <= roughness <= 1.0
    height, width = shape
    size = int(2**np.ceil(np.log2(max(height, width))) + 1)
    field = np.zeros((size, size), dtype=np.float32)
    for corner in ((0, 0), (0, -1), (-1, -1), (-1, 0)):
        field[corner] = rng.uniform(0.0, 1.0)

    step, level = size - 1, 0
    while step >= 2:
        weight = roughness**level
        keep = 1 - weight
        half = step // 2
        corners = field[0:size:step, 0:size:step]
        down = corners[:-1] + corners[1:]          # corner + the corner below
        right = corners[:, :-1] + corners[:, 1:]   # corner + the corner to the right
        m = down.shape[0]

        # centres of the squares
        centres = keep * (down[:, :-1] + right[:-1]) / 4 + weight * rng.uniform(0, 1, (m, m))
        field[half:size:step, half:size:step] = centres

        # edge midpoints on the corner rows: left / right corners + centres above / below (wrapping)
        above_below = np.empty((m + 1, m), dtype=centres.dtype)
        above_below[1:m] = centres[1:] + centres[:-1]
        above_below[0] = centres[0] + centres[-1]
        above_below[m] = above_below[0]
        field[0:size:step, half:size:step] = (keep * (right + above_below) / 4
                                              + weight * rng.uniform(0, 1, above_below.shape))

        # edge midpoints on the corner columns
        left_right = np.empty((m, m + 1), dtype=centres.dtype)
        left_right[:, 1:m] = centres[:, 1:] + centres[:, :-1]
        left_right[:, 0] = centres[:, 0] + centres[:, -1]
        left_right[:, m] = left_right[0, :m]  # (sic) the first row, see the docstring
        field[half:size:step, 0:size:step] = (keep * (down + left_right) / 4
                                              + weight * rng.uniform(0, 1, left_right.shape))
        level += 1
        step = half

    up = rng.integers(0, size - height + 1)
    left = rng.integers(0, size - width + 1)
    return field[up:up + height, left:left + width]


def diamond_square_alpha_device(shape: Tuple[int, int], roughness: float, ratio_min: float,
                                ratio_max: float, rng: RandomGenerator):
    """The fog alpha (generate_diamond_square_mask + fog_image's normalisation, effect.py:89-190) as
    a float32 CUDA tensor, computed on the device from the caller's generator stream.  NumPy's
    PCG64 can be advanced by any number of steps and every uniform double of the field is one
    64-bit output, so the device regenerates the array draws itself (`vkb_fog_mask`); the host makes
    the four scalar corner draws, advances the generator past the array draws and makes the two
    crop draws -- the generator ends in the state the reference leaves it in.  Returns None when
    the generator is not a PCG64 (the caller then draws the field on the host)."""
    bit_generator = rng.bit_generator
    if type(bit_generator).__name__ != 'PCG64':
        return None
    assert 0.0 <= roughness <= 1.0
    height, width = shape
    size = int(2**np.ceil(np.log2(max(height, width))) + 1)
    if size < 3:
        return None
    params = nv.FogParams()
    for k in range(4):  # (0, 0), (0, -1), (-1, -1), (-1, 0), stored as float32 like the field
        params.corners[k] = float(np.float32(rng.uniform(0.0, 1.0)))
    state = bit_generator.state
    s128, inc128 = state['state']['state'], state['state']['inc']
    mask64 = (1 << 64) - 1
    params.state_hi, params.state_lo = s128 >> 64, s128 & mask64
    params.inc_hi, params.inc_lo = inc128 >> 64, inc128 & mask64
    step, level = size - 1, 0
    while step >= 2:
        params.weight[level] = roughness**level
        level += 1
        step //= 2
    # past the array draws; the cached half of a 32-bit draw survives them in the reference
    count = ctypes.c_int64(0)
    nv.check(nv.lib().vkb_fog_draws(size, ctypes.byref(count)), 'vkb_fog_draws')
    n_draws = int(count.value)
    bit_generator.advance(n_draws)
    after = bit_generator.state
    after['has_uint32'], after['uinteger'] = state['has_uint32'], state['uinteger']
    bit_generator.state = after
    params.size = size
    params.up = int(rng.integers(0, size - height + 1))
    params.left = int(rng.integers(0, size - width + 1))
    params.height, params.width = height, width
    params.ratio_span = float(np.float32(ratio_max - ratio_min))
    params.ratio_min = float(np.float32(ratio_min))
    field = dv.empty((size, size), np.float32)
    centres = dv.empty((max(1, (size - 1) // 2) ** 2,), np.float64)
    draws = dv.empty((max(1, n_draws),), np.float64)
    minmax = dv.empty((2,), np.uint32)
    alpha = dv.empty((height, width), np.float32)
    nv.check(nv.lib().vkb_fog_mask(ctypes.byref(params), dv.ptr(field), dv.ptr(centres),
                                   dv.ptr(draws), dv.ptr(minmax), dv.ptr(alpha), dv.stream_ptr()),
             'vkb_fog_mask')
    return alpha


@attrs.define
class FogConfig(DistortionConfig):
    roughness: float
    fog_rgb: Tuple[int, int, int] = (226, 238, 234)
    ratio_max: float = 1.0
    ratio_min: float = 0.0

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


def fog_image(config: FogConfig, state, image: Image, rng: Optional[RandomGenerator]):
    # effect.py:169-208
    mode = image.mode
    image = to_rgb_image(image, mode)
    assert rng is not None
    assert config.ratio_min < config.ratio_max
    if config.ratio_min < 0.0 or config.ratio_max > 1.0:
        raise NotImplementedError('fog ratios outside [0, 1] are not provided')
    alpha = diamond_square_alpha_device(image.shape, config.roughness, config.ratio_min,
                                        config.ratio_max, rng)
    if alpha is None:  # a bit generator whose stream cannot be split: the field is drawn on the host
        mask = generate_diamond_square_mask(image.shape, config.roughness, rng)
        mask -= mask.min()
        mask /= mask.max()
        mask *= (config.ratio_max - config.ratio_min)
        mask += config.ratio_min
        alpha = dv.to_device(np.ascontiguousarray(mask, dtype=np.float32))

    channels = image.num_channels or 1
    if image.mode == ImageMode.GRAYSCALE:
        fog_value = [np.float32(0.2126 * config.fog_rgb[0] + 0.7152 * config.fog_rgb[1]
                                + 0.0722 * config.fog_rgb[2])]
    else:
        assert image.mode == ImageMode.RGB
        fog_value = [np.float32(v) for v in config.fog_rgb]

    # (1 - mask) * mat + mask * fog in float32, clipped and truncated: the device blend
    dst = image.dev.clone()
    item = nv.BlendItem()
    item.dst = dst.data_ptr()
    item.dst_f32 = 0
    item.channels = channels
    item.dst_w = image.width
    item.box_y, item.box_x, item.box_h, item.box_w = 0, 0, image.height, image.width
    item.value_arr = None
    for i, v in enumerate(fog_value):
        item.value_const[i] = float(v)
    item.mask = None
    item.alpha_arr = alpha.data_ptr()
    item.alpha_pitch = image.width
    item.keep_mode = nv.BLEND_FLOAT_CONST  # the GRAYSCALE fog value is fractional
    item.alpha = 1.0
    nv.check(nv.lib().vkb_blend_fill(ctypes.byref(item), dv.stream_ptr()), 'vkb_blend_fill')
    image = attrs.evolve(image, mat=dst)
    return to_original_image(image, mode)


fog = Distortion(config_cls=FogConfig, state_cls=DistortionNopState[FogConfig],
                 func_image=fog_image)
