"""Blur ops (vkit/mechanism/distortion/photometric/blur.py).  gaussian_blur runs as a separable
8.8 fixed-point shared-memory stencil that reproduces cv.GaussianBlur on uint8 bit-exactly."""
import ctypes
from typing import Any, Mapping, Optional

import attrs
import numpy as np
from numpy.random import Generator as RandomGenerator

from vkit_b200 import _native as nv
from vkit_b200 import device as dv
from vkit_b200.element import Image

from ..interface import Distortion, DistortionConfig, DistortionNopState
from .opt import to_original_image, to_rgb_image


def _estimate_gaussian_kernel_size(sigma: float):
    kernel_size = max(3, round(3 * sigma) + 1)
    if kernel_size % 2 == 0:
        kernel_size += 1
    return kernel_size


def gaussian_kernel_u8(ksize: int, sigma: float):
    """Integer taps (sum 256) cv2 derives for uint8 images: float Gaussian, normalised, then
    rounded to 8.8 fixed point from the outside in with error diffusion, centre = remainder."""
    r = ksize // 2
    xs = np.arange(ksize, dtype=np.float64) - r
    k = np.exp(-(xs * xs) / (2.0 * sigma * sigma))
    k = k / k.sum()
    taps = [0] * ksize
    err = 0.0
    for i in range(r):
        adj = k[i] * 256.0 + err
        v = float(np.rint(adj))
        err = adj - v
        taps[i] = taps[ksize - 1 - i] = int(v)
    taps[r] = 256 - sum(taps)
    return taps


def gaussian_blur_device(image: Image, sigma: float) -> Image:
    ksize = _estimate_gaussian_kernel_size(sigma)
    if ksize > 17:
        raise NotImplementedError('gaussian_blur kernels wider than 17 taps are not provided')
    taps = gaussian_kernel_u8(ksize, sigma)
    src = image.dev
    dst = dv.empty(tuple(src.shape), np.uint8)
    arr = (ctypes.c_int32 * ksize)(*taps)
    nv.check(nv.lib().vkb_gaussian_blur_u8(dv.ptr(src), dv.ptr(dst), image.height, image.width,
                                           image.num_channels or 1, arr, ksize, dv.stream_ptr()),
             'vkb_gaussian_blur_u8')
    return attrs.evolve(image, mat=dst)


@attrs.define
class GaussianBlurConfig(DistortionConfig):
    sigma: float


def gaussian_blur_image(config: GaussianBlurConfig, state, image: Image,
                        rng: Optional[RandomGenerator]):
    mode = image.mode
    image = to_rgb_image(image, mode)
    image = gaussian_blur_device(image, config.sigma)
    return to_original_image(image, mode)


gaussian_blur = Distortion(config_cls=GaussianBlurConfig,
                           state_cls=DistortionNopState[GaussianBlurConfig],
                           func_image=gaussian_blur_image)


def _next_row(name):
    def func(config, state, image, rng):
        raise NotImplementedError(
            f'{name} is a "next" row of the scope table (SURVEY.md section 8f) and has no device '
            'kernel yet; disable it via RandomDistortionFactoryConfig.disabled_policy_names.')
    return func


@attrs.define
class DefocusBlurConfig(DistortionConfig):
    radius: int
    anti_aliasing_sigma: float = 0.5


defocus_blur = Distortion(config_cls=DefocusBlurConfig,
                          state_cls=DistortionNopState[DefocusBlurConfig],
                          func_image=_next_row('defocus_blur'))


@attrs.define
class MotionBlurConfig(DistortionConfig):
    radius: int
    angle: int
    anti_aliasing_sigma: float = 0.5


motion_blur = Distortion(config_cls=MotionBlurConfig,
                         state_cls=DistortionNopState[MotionBlurConfig],
                         func_image=_next_row('motion_blur'))


@attrs.define
class GlassBlurConfig(DistortionConfig):
    sigma: float
    delta: int = 1
    loop: int = 5

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


glass_blur = Distortion(config_cls=GlassBlurConfig,
                        state_cls=DistortionNopState[GlassBlurConfig],
                        func_image=_next_row('glass_blur'))


@attrs.define
class ZoomInBlurConfig(DistortionConfig):
    ratio: float = 0.1
    step: float = 0.01
    alpha: float = 0.5


zoom_in_blur = Distortion(config_cls=ZoomInBlurConfig,
                          state_cls=DistortionNopState[ZoomInBlurConfig],
                          func_image=_next_row('zoom_in_blur'))
