"""Shapable protocol and set-operation modes (vkit/element/type.py)."""
from enum import Enum, unique
from typing import Tuple


class Shapable:

    @property
    def height(self) -> int:
        raise NotImplementedError()

    @property
    def width(self) -> int:
        raise NotImplementedError()

    @property
    def area(self) -> int:
        return self.height * self.width

    @property
    def shape(self) -> Tuple[int, int]:
        return self.height, self.width


@unique
class ElementSetOperationMode(Enum):
    UNION = 'union'
    DISTINCT = 'distinct'
    INTERSECT = 'intersect'
