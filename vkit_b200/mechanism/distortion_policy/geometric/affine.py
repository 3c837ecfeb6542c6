"""Config generators of the affine family (distortion_policy/geometric/affine.py)."""
from typing import Tuple

import attrs
from numpy.random import Generator as RandomGenerator

from vkit_b200.mechanism import distortion

from ..opt import sample_float, sample_int
from ..type import DistortionConfigGenerator, DistortionPolicyFactory


@attrs.define
class ShearHoriConfigGeneratorConfig:
    angle_min: int = 1
    angle_max: int = 30
    prob_negative: float = 0.5


@attrs.define
class ShearVertConfigGeneratorConfig:
    angle_min: int = 1
    angle_max: int = 30
    prob_negative: float = 0.5


@attrs.define
class RotateConfigGeneratorConfig:
    angle_min: int = 1
    angle_max: int = 180
    prob_negative: float = 0.5


@attrs.define
class SkewHoriConfigGeneratorConfig:
    ratio_min: float = 0.0
    ratio_max: float = 0.35
    prob_negative: float = 0.5


@attrs.define
class SkewVertConfigGeneratorConfig:
    ratio_min: float = 0.0
    ratio_max: float = 0.35
    prob_negative: float = 0.5


def _sample_angle(generator, rng: RandomGenerator):
    return sample_int(level=generator.level, value_min=generator.config.angle_min,
                      value_max=generator.config.angle_max,
                      prob_negative=generator.config.prob_negative, rng=rng)


def _sample_ratio(generator, rng: RandomGenerator):
    ratio = sample_float(level=generator.level, value_min=generator.config.ratio_min,
                         value_max=generator.config.ratio_max, prob_reciprocal=None, rng=rng)
    if rng.random() < generator.config.prob_negative:
        ratio *= -1
    return ratio


class ShearHoriConfigGenerator(
        DistortionConfigGenerator[ShearHoriConfigGeneratorConfig, distortion.ShearHoriConfig]):

    def __call__(self, shape: Tuple[int, int], rng: RandomGenerator):
        return distortion.ShearHoriConfig(angle=_sample_angle(self, rng))


class ShearVertConfigGenerator(
        DistortionConfigGenerator[ShearVertConfigGeneratorConfig, distortion.ShearVertConfig]):

    def __call__(self, shape: Tuple[int, int], rng: RandomGenerator):
        return distortion.ShearVertConfig(angle=_sample_angle(self, rng))


class RotateConfigGenerator(
        DistortionConfigGenerator[RotateConfigGeneratorConfig, distortion.RotateConfig]):

    def __call__(self, shape: Tuple[int, int], rng: RandomGenerator):
        return distortion.RotateConfig(angle=_sample_angle(self, rng))


class SkewHoriConfigGenerator(
        DistortionConfigGenerator[SkewHoriConfigGeneratorConfig, distortion.SkewHoriConfig]):

    def __call__(self, shape: Tuple[int, int], rng: RandomGenerator):
        return distortion.SkewHoriConfig(ratio=_sample_ratio(self, rng))


class SkewVertConfigGenerator(
        DistortionConfigGenerator[SkewVertConfigGeneratorConfig, distortion.SkewVertConfig]):

    def __call__(self, shape: Tuple[int, int], rng: RandomGenerator):
        return distortion.SkewVertConfig(ratio=_sample_ratio(self, rng))


shear_hori_policy_factory = DistortionPolicyFactory(distortion.shear_hori, ShearHoriConfigGenerator)
shear_vert_policy_factory = DistortionPolicyFactory(distortion.shear_vert, ShearVertConfigGenerator)
rotate_policy_factory = DistortionPolicyFactory(distortion.rotate, RotateConfigGenerator)
skew_hori_policy_factory = DistortionPolicyFactory(distortion.skew_hori, SkewHoriConfigGenerator)
skew_vert_policy_factory = DistortionPolicyFactory(distortion.skew_vert, SkewVertConfigGenerator)
