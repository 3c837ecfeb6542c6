"""Host-side helpers that the reference keeps in `vkit.utility` (vkit/utility/opt.py).

Only what the distortion path touches: lazily initialised attrs fields, rng helpers that must
consume the NumPy generator exactly like the reference (so configs match on identical seeds),
`dyn_structure` (dict -> attrs config; the reference uses cattrs, which this image lacks) and
the snake-case naming that policy names / conflict keywords are derived from.
"""
import enum
import json
import os
import re
import typing
from collections import abc
from os import PathLike
from typing import Any, List, Mapping, Optional, Sequence, Tuple, Type, TypeVar, Union, get_args

import attrs
from numpy.random import Generator as RandomGenerator

PathType = Union[str, PathLike]


def attrs_lazy_field():
    return attrs.field(default=None, init=False, repr=False)


_T_FIELD = TypeVar('_T_FIELD')


def unwrap_optional_field(field: Optional[_T_FIELD]) -> _T_FIELD:
    assert field is not None
    return field


def is_path_type(path: Any):
    return isinstance(path, (str, PathLike))


def read_json_file(path: PathType):
    with open(os.path.expandvars(os.path.expanduser(os.fspath(path))), 'r') as fin:
        return json.loads(os.path.expandvars(fin.read()))


_T_ITEM = TypeVar('_T_ITEM')


def rng_choice(rng: RandomGenerator, items: Sequence[_T_ITEM],
               probs: Optional[Sequence[float]] = None) -> _T_ITEM:
    # vkit/utility/opt.py:95-101
    idx = rng.choice(len(items), p=probs)
    return items[idx]


def rng_choice_with_size(rng: RandomGenerator, items: Sequence[_T_ITEM], size: int,
                         probs: Optional[Sequence[float]] = None,
                         replace: bool = True) -> Sequence[_T_ITEM]:
    # vkit/utility/opt.py:104-113
    indices = rng.choice(len(items), p=probs, size=size, replace=replace)
    return [items[idx] for idx in indices]


def rng_shuffle(rng: RandomGenerator, items: Sequence[_T_ITEM]) -> Sequence[_T_ITEM]:
    indices = list(range(len(items)))
    rng.shuffle(indices)
    return tuple(items[idx] for idx in indices)


def normalize_to_probs(weights: Sequence[float]):
    total = sum(weights)
    return [weight / total for weight in weights]


def normalize_to_keys_and_probs(key_weight_items):
    if isinstance(key_weight_items, abc.Mapping):
        pairs = list(key_weight_items.items())
    else:
        pairs = list(key_weight_items)
    keys = [key for key, _ in pairs]
    return keys, normalize_to_probs([weight for _, weight in pairs])


def convert_camel_case_name_to_snake_case_name(name: str):
    return re.sub(r'(?<!^)(?=[A-Z])', '_', name).lower()


def get_config_class_snake_case_name(class_name: str):
    # vkit/utility/opt.py:235-243: FooBarConfig -> foo_bar
    name = convert_camel_case_name_to_snake_case_name(class_name)
    if name.endswith('_config'):
        name = name[:-len('_config')]
    return name


def get_generic_classes(cls: Type[Any]):
    return get_args(cls.__orig_bases__[0])  # type: ignore


# ------------------------------------------------------------------------------------------
# dyn_structure: Mapping / Sequence / path / instance -> instance of target_cls.
# ------------------------------------------------------------------------------------------
_T_TARGET = TypeVar('_T_TARGET')


def _structure(obj: Any, tp: Any):
    if tp is None or tp is Any:
        return obj
    origin = typing.get_origin(tp)
    if origin is Union:
        args = [a for a in get_args(tp) if a is not type(None)]
        if obj is None:
            return None
        if len(args) == 1:
            return _structure(obj, args[0])
        return obj
    if isinstance(tp, type) and attrs.has(tp):
        if isinstance(obj, tp):
            return obj
        if isinstance(obj, abc.Mapping):
            return _structure_attrs(obj, tp)
        return obj
    if isinstance(tp, type) and issubclass(tp, enum.Enum):
        return obj if isinstance(obj, tp) else tp(obj)
    if origin in (list, abc.Sequence, typing.Sequence) and isinstance(obj, (list, tuple)):
        args = get_args(tp)
        return [_structure(item, args[0] if args else None) for item in obj]
    if origin is tuple and isinstance(obj, (list, tuple)):
        args = get_args(tp)
        if len(args) == 2 and args[1] is Ellipsis:
            return tuple(_structure(item, args[0]) for item in obj)
        if len(args) == len(obj):
            return tuple(_structure(item, arg) for item, arg in zip(obj, args))
        return tuple(obj)
    if tp is float and isinstance(obj, (int, float)) and not isinstance(obj, bool):
        return float(obj)
    if tp is int and isinstance(obj, (int, float)) and not isinstance(obj, bool):
        return int(obj)
    return obj


def _structure_attrs(obj: Mapping[str, Any], cls: Type[_T_TARGET]) -> _T_TARGET:
    try:
        hints = typing.get_type_hints(cls)
    except Exception:
        hints = {}
    fields = {f.alias: f for f in attrs.fields(cls) if f.init}
    extra = set(obj) - set(fields)
    if extra:
        # forbid_extra_keys=True in the reference's converter (vkit/utility/opt.py:153)
        raise TypeError(f'Extra keys {sorted(extra)} for {cls.__name__}')
    kwargs = {}
    for alias, field in fields.items():
        if alias in obj:
            kwargs[alias] = _structure(obj[alias], hints.get(field.name))
    return cls(**kwargs)


def dyn_structure(
    dyn_object: Any,
    target_cls: Type[_T_TARGET],
    support_path_type: bool = False,
    force_path_type: bool = False,
    support_none_type: bool = False,
) -> _T_TARGET:
    # vkit/utility/opt.py:162-202
    if support_none_type and dyn_object is None:
        return target_cls()

    if support_path_type or force_path_type:
        dyn_object_is_path_type = is_path_type(dyn_object)
        if force_path_type:
            assert dyn_object_is_path_type
        if dyn_object_is_path_type:
            dyn_object = read_json_file(dyn_object)

    try:
        if isinstance(dyn_object, target_cls):
            return dyn_object
    except TypeError:
        pass

    if isinstance(dyn_object, abc.Mapping):
        if isinstance(target_cls, type) and attrs.has(target_cls):
            return _structure_attrs(dyn_object, target_cls)
        return target_cls(**dyn_object)
    if isinstance(dyn_object, abc.Sequence):
        return _structure(list(dyn_object), target_cls)
    raise NotImplementedError(f'dyn_structure: cannot structure {type(dyn_object)}')
