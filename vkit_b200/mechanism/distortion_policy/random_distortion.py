"""RandomDistortion: staged random chain of policies
(vkit/mechanism/distortion_policy/random_distortion.py:66-671).

Host orchestration only; every draw from `rng` happens in the same order as in the reference
so that identical seeds produce identical chains and configs.
"""
import logging
from collections import defaultdict
from typing import Any, Iterable, List, Mapping, Optional, Sequence, Tuple, Union

import attrs
from numpy.random import Generator as RandomGenerator

from vkit_b200.element import (Box, Image, Mask, Point, PointList, PointTuple, Polygon, ScoreMap,
                               Shapable)
from vkit_b200.utility import PathType, dyn_structure, normalize_to_probs, rng_choice_with_size

from ..distortion.interface import Distortion, DistortionResult
from .geometric import affine, camera, mls
from .opt import LEVEL_MAX, LEVEL_MIN
from .photometric import blur, color, effect, noise, streak
from .type import DistortionPolicy, DistortionPolicyFactory

logger = logging.getLogger(__name__)


@attrs.define
class RandomDistortionDebug:
    distortion_names: List[str] = attrs.field(factory=list)
    distortion_levels: List[int] = attrs.field(factory=list)
    distortion_images: List[Image] = attrs.field(factory=list)
    distortion_configs: List[Any] = attrs.field(factory=list)
    distortion_states: List[Any] = attrs.field(factory=list)


@attrs.define
class RandomDistortionStageConfig:
    distortion_policies: Sequence[DistortionPolicy]
    distortion_policy_weights: Sequence[float]
    prob_enable: float
    num_distortions_min: int
    num_distortions_max: int
    inject_corner_points: bool = False
    conflict_control_keyword_groups: Sequence[Sequence[str]] = ()
    force_sample_level_in_full_range: bool = False


class RandomDistortionStage:

    def __init__(self, config: RandomDistortionStageConfig):
        self.config = config
        self.distortion_policy_probs = normalize_to_probs(self.config.distortion_policy_weights)

    def _has_conflict(self, policies: Sequence[DistortionPolicy]):
        hits = defaultdict(int)
        for policy in policies:
            for group_idx, keywords in enumerate(self.config.conflict_control_keyword_groups):
                if any(keyword in policy.name for keyword in keywords):
                    hits[group_idx] += 1
                    break
        return any(count > 1 for count in hits.values())

    def sample_distortion_policies(self, rng: RandomGenerator) -> Sequence[DistortionPolicy]:
        num_distortions = rng.integers(self.config.num_distortions_min,
                                       self.config.num_distortions_max + 1)
        if num_distortions <= 0:
            return ()
        for _ in range(5):
            policies = rng_choice_with_size(rng, self.config.distortion_policies,
                                            size=num_distortions,
                                            probs=self.distortion_policy_probs, replace=False)
            if not self._has_conflict(policies):
                return policies
        logger.warning(f'Cannot sample distortion policies with num_distortion={num_distortions}.')
        return ()

    @classmethod
    def generate_corner_points(cls, shape: Tuple[int, int]):
        # points along the page border, used to trim the page after a geometric op
        height, width = shape
        step = min(height // 4, width // 4)
        assert step > 0
        ys = list(range(0, height, step))
        if ys[-1] < height - 1:
            ys.append(height - 1)
        xs = list(range(0, width, step))
        if xs[0] == 0:
            xs.pop(0)
        if xs[-1] == width - 1:
            xs.pop()
        corner_points = PointList()
        for x in (0, width - 1):
            corner_points.extend(Point.create(y=y, x=x) for y in ys)
        for y in (0, height - 1):
            corner_points.extend(Point.create(y=y, x=x) for x in xs)
        return corner_points.to_point_tuple()

    def apply_distortions(self, distortion_result: DistortionResult, level_min: int,
                          level_max: int, rng: RandomGenerator,
                          debug: Optional[RandomDistortionDebug] = None):
        if rng.random() > self.config.prob_enable:
            return distortion_result

        if self.config.inject_corner_points:
            distortion_result.corner_points = self.generate_corner_points(distortion_result.shape)

        if self.config.force_sample_level_in_full_range:
            level_min, level_max = LEVEL_MIN, LEVEL_MAX

        for policy in self.sample_distortion_policies(rng):
            level = rng.integers(level_min, level_max + 1)
            distortion_result = policy.distort(
                level=level,
                shapable_or_shape=distortion_result.shape,
                image=distortion_result.image,
                mask=distortion_result.mask,
                score_map=distortion_result.score_map,
                point=distortion_result.point,
                points=distortion_result.points,
                corner_points=distortion_result.corner_points,
                polygon=distortion_result.polygon,
                polygons=distortion_result.polygons,
                rng=rng,
                enable_debug=bool(debug),
            )
            if debug:
                assert distortion_result.image
                debug.distortion_images.append(distortion_result.image)
                debug.distortion_names.append(policy.name)
                debug.distortion_levels.append(level)
                debug.distortion_configs.append(distortion_result.config)
                debug.distortion_states.append(distortion_result.state)
            distortion_result.config = None
            distortion_result.state = None
        return distortion_result


class RandomDistortion:

    def __init__(self, configs: Sequence[RandomDistortionStageConfig], level_min: int,
                 level_max: int):
        self.stages = [RandomDistortionStage(config) for config in configs]
        self.level_min = level_min
        self.level_max = level_max

    @classmethod
    def get_distortion_result_all_points(cls, distortion_result: DistortionResult):
        if distortion_result.corner_points:
            yield from distortion_result.corner_points
        if distortion_result.point:
            yield distortion_result.point
        if distortion_result.points:
            yield from distortion_result.points
        if distortion_result.polygon:
            yield from distortion_result.polygon.points
        if distortion_result.polygons:
            for polygon in distortion_result.polygons:
                yield from polygon.points

    @classmethod
    def get_distortion_result_element_bounding_box(cls, distortion_result: DistortionResult):
        assert distortion_result.corner_points
        points = list(cls.get_distortion_result_all_points(distortion_result))
        return Box(up=min(p.y for p in points), down=max(p.y for p in points),
                   left=min(p.x for p in points), right=max(p.x for p in points))

    @classmethod
    def trim_distortion_result(cls, distortion_result: DistortionResult):
        # crop the page to the bounding box of everything that was tracked (:266-348)
        if not distortion_result.corner_points:
            return distortion_result
        height, width = distortion_result.shape
        box = cls.get_distortion_result_element_bounding_box(distortion_result)
        pad_up, pad_down = box.up, height - 1 - box.down
        pad_left, pad_right = box.left, width - 1 - box.right
        assert pad_up >= -1 and pad_down >= -1  # rounding slack
        assert pad_left >= -1 and pad_right >= -1
        if pad_up <= 0 and pad_down <= 0 and pad_left <= 0 and pad_right <= 0:
            return distortion_result

        up, down = max(0, box.up), min(height - 1, box.down)
        left, right = max(0, box.left), min(width - 1, box.right)
        pad_up, pad_left = max(0, pad_up), max(0, pad_left)

        if distortion_result.image:
            distortion_result.image = distortion_result.image.to_cropped_image(
                up=up, down=down, left=left, right=right)
        if distortion_result.mask:
            distortion_result.mask = distortion_result.mask.to_cropped_mask(
                up=up, down=down, left=left, right=right)
        if distortion_result.score_map:
            distortion_result.score_map = distortion_result.score_map.to_cropped_score_map(
                up=up, down=down, left=left, right=right)
        if distortion_result.point:
            distortion_result.point = distortion_result.point.to_shifted_point(
                offset_y=-pad_up, offset_x=-pad_left)
        if distortion_result.points:
            distortion_result.points = distortion_result.points.to_shifted_points(
                offset_y=-pad_up, offset_x=-pad_left)
        if distortion_result.polygon:
            distortion_result.polygon = distortion_result.polygon.to_shifted_polygon(
                offset_y=-pad_up, offset_x=-pad_left)
        if distortion_result.polygons:
            distortion_result.polygons = [
                polygon.to_shifted_polygon(offset_y=-pad_up, offset_x=-pad_left)
                for polygon in distortion_result.polygons
            ]
        return distortion_result

    def distort(
        self,
        rng: RandomGenerator,
        shapable_or_shape: Optional[Union[Shapable, Tuple[int, int]]] = None,
        image: Optional[Image] = None,
        mask: Optional[Mask] = None,
        score_map: Optional[ScoreMap] = None,
        point: Optional[Point] = None,
        points: Optional[Union[PointList, PointTuple, Iterable[Point]]] = None,
        polygon: Optional[Polygon] = None,
        polygons: Optional[Iterable[Polygon]] = None,
        debug: Optional[RandomDistortionDebug] = None,
    ):
        shape = Distortion.get_shape(shapable_or_shape=shapable_or_shape, image=image, mask=mask,
                                     score_map=score_map)
        result = DistortionResult(shape=shape)
        result.image = image
        result.mask = mask
        result.score_map = score_map
        result.point = point
        result.points = PointTuple(points) if points else None
        result.polygon = polygon
        if polygons:
            result.polygons = tuple(polygons)

        for stage in self.stages:
            result = stage.apply_distortions(distortion_result=result, level_min=self.level_min,
                                             level_max=self.level_max, rng=rng, debug=debug)
        return self.trim_distortion_result(result)


@attrs.define
class RandomDistortionFactoryConfig:
    # Photometric.
    prob_photometric: float = 1.0
    num_photometric_min: int = 0
    num_photometric_max: int = 2
    photometric_conflict_control_keyword_groups: Sequence[Sequence[str]] = attrs.field(
        factory=lambda: [['blur', 'pixelation', 'jpeg'], ['noise']])
    # Geometric.
    prob_geometric: float = 0.75
    force_post_rotate: bool = False
    # Shared.
    level_min: int = LEVEL_MIN
    level_max: int = LEVEL_MAX
    disabled_policy_names: Sequence[str] = attrs.field(factory=list)
    name_to_policy_config: Mapping[str, Any] = attrs.field(factory=dict)
    name_to_policy_weight: Mapping[str, float] = attrs.field(factory=dict)


# (factories of one group, total default weight of the group), random_distortion.py:424-501
_PHOTOMETRIC_POLICY_FACTORIES_AND_DEFAULT_WEIGHTS_SUM_PAIRS = (
    ((color.mean_shift_policy_factory, color.color_shift_policy_factory,
      color.brightness_shift_policy_factory, color.std_shift_policy_factory,
      color.boundary_equalization_policy_factory, color.histogram_equalization_policy_factory,
      color.complement_policy_factory, color.posterization_policy_factory,
      color.color_balance_policy_factory, color.channel_permutation_policy_factory), 10.0),
    ((blur.gaussian_blur_policy_factory, blur.defocus_blur_policy_factory,
      blur.motion_blur_policy_factory, blur.glass_blur_policy_factory,
      blur.zoom_in_blur_policy_factory), 1.0),
    ((noise.gaussion_noise_policy_factory, noise.poisson_noise_policy_factory,
      noise.impulse_noise_policy_factory, noise.speckle_noise_policy_factory), 3.0),
    ((effect.jpeg_quality_policy_factory, effect.pixelation_policy_factory,
      effect.fog_policy_factory), 1.0),
    ((streak.line_streak_policy_factory, streak.rectangle_streak_policy_factory,
      streak.ellipse_streak_policy_factory), 1.0),
)

_GEOMETRIC_POLICY_FACTORIES_AND_DEFAULT_WEIGHTS_SUM_PAIRS = (
    ((affine.shear_hori_policy_factory, affine.shear_vert_policy_factory,
      affine.rotate_policy_factory, affine.skew_hori_policy_factory,
      affine.skew_vert_policy_factory), 1.0),
    ((mls.similarity_mls_policy_factory,), 1.0),
    ((camera.camera_plane_only_policy_factory, camera.camera_cubic_curve_policy_factory,
      camera.camera_plane_line_fold_policy_factory,
      camera.camera_plane_line_curve_policy_factory), 1.0),
)


class RandomDistortionFactory:

    @classmethod
    def unpack_policy_factories_and_default_weights_sum_pairs(cls, pairs):
        factories: List[DistortionPolicyFactory] = []
        weights: List[float] = []
        for group, weights_sum in pairs:
            factories.extend(group)
            weights.extend([weights_sum / len(group)] * len(group))
        return factories, weights

    def __init__(
        self,
        photometric_policy_factories_and_default_weights_sum_pairs=(
            _PHOTOMETRIC_POLICY_FACTORIES_AND_DEFAULT_WEIGHTS_SUM_PAIRS),
        geometric_policy_factories_and_default_weights_sum_pairs=(
            _GEOMETRIC_POLICY_FACTORIES_AND_DEFAULT_WEIGHTS_SUM_PAIRS),
    ):
        (self.photometric_policy_factories, self.photometric_policy_default_weights
         ) = self.unpack_policy_factories_and_default_weights_sum_pairs(
             photometric_policy_factories_and_default_weights_sum_pairs)
        (self.geometric_policy_factories, self.geometric_policy_default_weights
         ) = self.unpack_policy_factories_and_default_weights_sum_pairs(
             geometric_policy_factories_and_default_weights_sum_pairs)

    @classmethod
    def create_policies_and_policy_weights(cls, policy_factories, policy_default_weights,
                                           config: RandomDistortionFactoryConfig):
        disabled = set(config.disabled_policy_names)
        policies: List[DistortionPolicy] = []
        weights: List[float] = []
        for factory, default_weight in zip(policy_factories, policy_default_weights):
            if factory.name in disabled:
                continue
            policies.append(factory.create(config.name_to_policy_config.get(factory.name)))
            weights.append(config.name_to_policy_weight.get(factory.name, default_weight))
        return policies, weights

    def create(self, config: Optional[Union[Mapping[str, Any], PathType,
                                            RandomDistortionFactoryConfig]] = None):
        config = dyn_structure(config, RandomDistortionFactoryConfig, support_path_type=True,
                               support_none_type=True)
        stage_configs: List[RandomDistortionStageConfig] = []

        policies, weights = self.create_policies_and_policy_weights(
            self.photometric_policy_factories, self.photometric_policy_default_weights, config)
        stage_configs.append(RandomDistortionStageConfig(
            distortion_policies=policies,
            distortion_policy_weights=weights,
            prob_enable=config.prob_photometric,
            num_distortions_min=config.num_photometric_min,
            num_distortions_max=config.num_photometric_max,
            conflict_control_keyword_groups=config.photometric_conflict_control_keyword_groups,
        ))

        policies, weights = self.create_policies_and_policy_weights(
            self.geometric_policy_factories, self.geometric_policy_default_weights, config)
        post_rotate_policy = None
        if config.force_post_rotate:
            idx = next(i for i, policy in enumerate(policies) if policy.name == 'rotate')
            post_rotate_policy = policies.pop(idx)
            weights.pop(idx)
        stage_configs.append(RandomDistortionStageConfig(
            distortion_policies=policies,
            distortion_policy_weights=weights,
            prob_enable=config.prob_geometric,
            num_distortions_min=1,
            num_distortions_max=1,
            inject_corner_points=config.force_post_rotate,
        ))
        if post_rotate_policy:
            stage_configs.append(RandomDistortionStageConfig(
                distortion_policies=[post_rotate_policy],
                distortion_policy_weights=[1.0],
                prob_enable=1.0,
                num_distortions_min=1,
                num_distortions_max=1,
                force_sample_level_in_full_range=True,
            ))
        return RandomDistortion(configs=stage_configs, level_min=config.level_min,
                                level_max=config.level_max)


random_distortion_factory = RandomDistortionFactory()
