"""`vkit.mechanism.distortion_policy` surface (vkit/mechanism/distortion_policy/__init__.py)."""
from .random_distortion import (RandomDistortion, RandomDistortionDebug, RandomDistortionFactory,
                                RandomDistortionFactoryConfig, random_distortion_factory)
