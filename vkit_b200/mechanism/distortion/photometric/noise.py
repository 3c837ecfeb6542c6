"""Noise ops (vkit/mechanism/distortion/photometric/noise.py).

Two modes.  Default (`VKIT_B200_NOISE=philox` or unset): a counter-based Philox generator on
the device keyed by a seed drawn from the op's private rng and the element index --
distributional parity with the reference, no host work per pixel.  `VKIT_B200_NOISE=field`
(or `use_host_field(True)`): the host draws exactly the field the reference would draw from
the same generator state and the device applies it -- bit-exact, but bounded by NumPy's
sequential sampler (it exists for parity tests)."""
import os
from typing import Any, Mapping, Optional

import attrs
import numpy as np
from numpy.random import Generator as RandomGenerator

from vkit_b200 import _native as nv
from vkit_b200 import device as dv
from vkit_b200.element import Image

from ..interface import Distortion, DistortionConfig, DistortionNopState

_HOST_FIELD = os.environ.get('VKIT_B200_NOISE', 'philox') == 'field'


def use_host_field(flag: bool):
    global _HOST_FIELD
    _HOST_FIELD = bool(flag)


def _apply(image: Image, kind: int, rng: RandomGenerator, p0=0.0, p1=0.0, field_fn=None):
    if image.mat_dtype != np.uint8:
        raise NotImplementedError('noise ops expect uint8 images')
    src = image.dev
    dst = dv.empty(tuple(src.shape), np.uint8)
    n_pixels = image.height * image.width
    channels = image.num_channels or 1
    if _HOST_FIELD:
        field = dv.to_device(np.ascontiguousarray(field_fn()))
        nv.check(nv.lib().vkb_noise_field(dv.ptr(src), dv.ptr(dst), n_pixels, channels, kind,
                                          dv.ptr(field), dv.stream_ptr()), 'vkb_noise_field')
    else:
        seed = int(rng.integers(0, 2**63 - 1))
        nv.check(nv.lib().vkb_noise_philox(dv.ptr(src), dv.ptr(dst), n_pixels, channels, kind,
                                           float(p0), float(p1), seed, dv.stream_ptr()),
                 'vkb_noise_philox')
    return Image(mat=dst)  # mode is dropped, like the reference (noise.py:54, 90, 144, 183)


class _RngStateConfig(DistortionConfig):

    @property
    def supports_rng_state(self) -> bool:
        return True

    @property
    def rng_state(self) -> Optional[Mapping[str, Any]]:
        return self._rng_state

    @rng_state.setter
    def rng_state(self, val: Mapping[str, Any]):
        self._rng_state = val


@attrs.define
class GaussionNoiseConfig(_RngStateConfig):
    std: float
    _rng_state: Optional[Mapping[str, Any]] = None


def gaussion_noise_image(config: GaussionNoiseConfig, state, image: Image,
                         rng: Optional[RandomGenerator]):
    assert rng
    shape = image.mat_shape
    return _apply(image, nv.NOISE_GAUSSIAN, rng, p0=config.std,
                  field_fn=lambda: np.round(rng.normal(0, config.std, shape)).astype(np.int16))


gaussion_noise = Distortion(config_cls=GaussionNoiseConfig,
                            state_cls=DistortionNopState[GaussionNoiseConfig],
                            func_image=gaussion_noise_image)


@attrs.define
class PoissonNoiseConfig(_RngStateConfig):
    _rng_state: Optional[Mapping[str, Any]] = None


def poisson_noise_image(config: PoissonNoiseConfig, state, image: Image,
                        rng: Optional[RandomGenerator]):
    assert rng
    return _apply(image, nv.NOISE_POISSON, rng,
                  field_fn=lambda: rng.poisson(image.mat.astype(np.float32)).astype(np.int64))


poisson_noise = Distortion(config_cls=PoissonNoiseConfig,
                           state_cls=DistortionNopState[PoissonNoiseConfig],
                           func_image=poisson_noise_image)


@attrs.define
class ImpulseNoiseConfig(_RngStateConfig):
    prob_salt: float
    prob_pepper: float
    _rng_state: Optional[Mapping[str, Any]] = None


def impulse_noise_image(config: ImpulseNoiseConfig, state, image: Image,
                        rng: Optional[RandomGenerator]):
    assert rng
    prob_presv = 1 - config.prob_salt - config.prob_pepper
    assert prob_presv >= 0.0
    return _apply(
        image, nv.NOISE_IMPULSE, rng, p0=config.prob_salt, p1=config.prob_pepper,
        field_fn=lambda: rng.choice((0, 1, 2), size=image.shape,
                                    p=[prob_presv, config.prob_salt, config.prob_pepper]).astype(
                                        np.int64))


impulse_noise = Distortion(config_cls=ImpulseNoiseConfig,
                           state_cls=DistortionNopState[ImpulseNoiseConfig],
                           func_image=impulse_noise_image)


@attrs.define
class SpeckleNoiseConfig(_RngStateConfig):
    std: float
    _rng_state: Optional[Mapping[str, Any]] = None


def speckle_noise_image(config: SpeckleNoiseConfig, state, image: Image,
                        rng: Optional[RandomGenerator]):
    assert rng
    shape = image.mat_shape
    return _apply(image, nv.NOISE_SPECKLE, rng, p0=config.std,
                  field_fn=lambda: rng.normal(0, config.std, shape).astype(np.float64))


speckle_noise = Distortion(config_cls=SpeckleNoiseConfig,
                           state_cls=DistortionNopState[SpeckleNoiseConfig],
                           func_image=speckle_noise_image)
