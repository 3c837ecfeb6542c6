"""Scalar helpers shared by the containers (vkit/element/opt.py:24-92)."""
from typing import Optional, Tuple, Union

from .type import Shapable


def clip_val(val, size: int):
    return max(0, min(val, size - 1))


def resize_val(val, size: int, resized_size: int):
    return clip_val(val * resized_size / size, resized_size)


def extract_shape_from_shapable_or_shape(shapable_or_shape: Union[Shapable, Tuple[int, int]]):
    if isinstance(shapable_or_shape, Shapable):
        return shapable_or_shape.shape
    height, width = shapable_or_shape
    return height, width


def generate_resized_shape(height: int, width: int, resized_height: Optional[int] = None,
                           resized_width: Optional[int] = None):
    if not resized_height and not resized_width:
        raise RuntimeError('Missing resized_height or resized_width.')
    if resized_height is None:
        resized_height = round(resized_width * height / width)
    if resized_width is None:
        resized_width = round(resized_height * width / height)
    return resized_height, resized_width


def generate_shape_and_resized_shape(shapable_or_shape, resized_height: Optional[int] = None,
                                     resized_width: Optional[int] = None):
    height, width = extract_shape_from_shapable_or_shape(shapable_or_shape)
    resized_height, resized_width = generate_resized_shape(height, width, resized_height,
                                                           resized_width)
    return height, width, resized_height, resized_width
