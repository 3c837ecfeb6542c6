"""Effect ops (vkit/mechanism/distortion/photometric/effect.py): jpeg_quality, pixelation, fog.
All three are "next" rows of the scope table (libjpeg codec, cv.resize models, sequential
diamond-square RNG recursion); the config classes exist so policies / configs stay
interchangeable with the reference."""
from typing import Any, Mapping, Optional, Tuple

import attrs

from ..interface import Distortion, DistortionConfig, DistortionNopState
from .blur import _next_row


@attrs.define
class JpegQualityConfig(DistortionConfig):
    quality: int


jpeg_quality = Distortion(config_cls=JpegQualityConfig,
                          state_cls=DistortionNopState[JpegQualityConfig],
                          func_image=_next_row('jpeg_quality'))


@attrs.define
class PixelationConfig(DistortionConfig):
    ratio: float


pixelation = Distortion(config_cls=PixelationConfig,
                        state_cls=DistortionNopState[PixelationConfig],
                        func_image=_next_row('pixelation'))


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


fog = Distortion(config_cls=FogConfig, state_cls=DistortionNopState[FogConfig],
                 func_image=_next_row('fog'))
