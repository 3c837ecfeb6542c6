"""Containers of the distortion path: the `vkit.element` surface (vkit/element/__init__.py)."""
from .type import Shapable, ElementSetOperationMode
from .point import Point, PointList, PointTuple
from .box import Box
from .polygon import Polygon
from .mask import Mask, MaskSetItemConfig
from .score_map import ScoreMap, ScoreMapSetItemConfig
from .image import Image, ImageMode, ImageSetItemConfig
