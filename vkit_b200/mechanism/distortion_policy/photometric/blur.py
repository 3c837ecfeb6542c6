"""Config generators of the blur ops (distortion_policy/photometric/blur.py)."""
from typing import Tuple

import attrs
from numpy.random import Generator as RandomGenerator

from vkit_b200.mechanism import distortion

from ..opt import sample_float, sample_int
from ..type import DistortionConfigGenerator, DistortionPolicyFactory


def _level_float(generator, rng, lo, hi, **kwargs):
    return sample_float(level=generator.level, value_min=lo, value_max=hi, prob_reciprocal=None,
                        rng=rng, **kwargs)


def _level_int(generator, rng, lo, hi, **kwargs):
    return sample_int(level=generator.level, value_min=lo, value_max=hi, prob_negative=None,
                      rng=rng, **kwargs)


@attrs.define
class GaussianBlurConfigGeneratorConfig:
    sigma_min: float = 0.5
    sigma_max: float = 1.0


class GaussianBlurConfigGenerator(
        DistortionConfigGenerator[GaussianBlurConfigGeneratorConfig,
                                  distortion.GaussianBlurConfig]):

    def __call__(self, shape: Tuple[int, int], rng: RandomGenerator):
        return distortion.GaussianBlurConfig(
            sigma=_level_float(self, rng, self.config.sigma_min, self.config.sigma_max))


gaussian_blur_policy_factory = DistortionPolicyFactory(distortion.gaussian_blur,
                                                       GaussianBlurConfigGenerator)


@attrs.define
class DefocusBlurConfigGeneratorConfig:
    radius_min: int = 1
    radius_max: int = 2


class DefocusBlurConfigGenerator(
        DistortionConfigGenerator[DefocusBlurConfigGeneratorConfig,
                                  distortion.DefocusBlurConfig]):

    def __call__(self, shape: Tuple[int, int], rng: RandomGenerator):
        return distortion.DefocusBlurConfig(
            radius=_level_int(self, rng, self.config.radius_min, self.config.radius_max))


defocus_blur_policy_factory = DistortionPolicyFactory(distortion.defocus_blur,
                                                      DefocusBlurConfigGenerator)


@attrs.define
class MotionBlurConfigGeneratorConfig:
    radius_min: int = 1
    radius_max: int = 2


class MotionBlurConfigGenerator(
        DistortionConfigGenerator[MotionBlurConfigGeneratorConfig, distortion.MotionBlurConfig]):

    def __call__(self, shape: Tuple[int, int], rng: RandomGenerator):
        radius = _level_int(self, rng, self.config.radius_min, self.config.radius_max)
        angle = rng.integers(0, 360)
        return distortion.MotionBlurConfig(radius=radius, angle=angle)


motion_blur_policy_factory = DistortionPolicyFactory(distortion.motion_blur,
                                                     MotionBlurConfigGenerator)


@attrs.define
class GlassBlurConfigGeneratorConfig:
    sigma_min: float = 0.5
    sigma_max: float = 1.0
    delta_min: int = 1
    delta_max: int = 1
    loop_min: int = 1
    loop_max: int = 4


class GlassBlurConfigGenerator(
        DistortionConfigGenerator[GlassBlurConfigGeneratorConfig, distortion.GlassBlurConfig]):

    def __call__(self, shape: Tuple[int, int], rng: RandomGenerator):
        cfg = self.config
        sigma = _level_float(self, rng, cfg.sigma_min, cfg.sigma_max)
        delta = _level_int(self, rng, cfg.delta_min, cfg.delta_max)
        loop = _level_int(self, rng, cfg.loop_min, cfg.loop_max)
        return distortion.GlassBlurConfig(sigma=sigma, delta=delta, loop=loop)


glass_blur_policy_factory = DistortionPolicyFactory(distortion.glass_blur,
                                                    GlassBlurConfigGenerator)


@attrs.define
class ZoomInBlurConfigGeneratorConfig:
    ratio_min: float = 0.01
    ratio_max: float = 0.1
    step_min: float = 0.002
    step_max: float = 0.02
    alpha_min: float = 0.5
    alpha_max: float = 0.7


class ZoomInBlurConfigGenerator(
        DistortionConfigGenerator[ZoomInBlurConfigGeneratorConfig, distortion.ZoomInBlurConfig]):

    def __call__(self, shape: Tuple[int, int], rng: RandomGenerator):
        cfg = self.config
        ratio = _level_float(self, rng, cfg.ratio_min, cfg.ratio_max)
        step = _level_float(self, rng, cfg.step_min, cfg.step_max)
        alpha = rng.uniform(cfg.alpha_min, cfg.alpha_max)
        return distortion.ZoomInBlurConfig(ratio=ratio, step=step, alpha=alpha)


zoom_in_blur_policy_factory = DistortionPolicyFactory(distortion.zoom_in_blur,
                                                      ZoomInBlurConfigGenerator)
