"""Loader for the LIVE reference (vkit-x/vkit) -- test infrastructure only.

`/root/reference` exists only in the build container, never on the GPU box.  This module
is used by `tests/golden/make_golden.py` (fixture generation) and by CPU tests that are
skipped when the reference tree is absent.  Nothing under `vkit_b200/` imports it.

The reference imports `iolite`, `cattrs`, `shapely`, `pyclipper` at module import time
(vkit/element/{box,polygon,mask}.py:19-28, vkit/utility/opt.py:34-37) but never calls them
on the distortion / blend path, so inert stand-ins are registered.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get('VKIT_REFERENCE_ROOT', '/root/reference')


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'vkit'))


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _structure(obj, cls):
    # Minimal stand-in for cattrs.structure: recurse into nested attrs classes and enums.
    import attrs
    import enum
    import typing
    if isinstance(cls, type) and attrs.has(cls) and isinstance(obj, dict):
        hints = typing.get_type_hints(cls)
        kwargs = {}
        for field in attrs.fields(cls):
            key = field.alias if field.alias in obj else field.name.lstrip('_')
            if key not in obj:
                continue
            kwargs[field.alias] = _structure(obj[key], hints.get(field.name, None))
        extra = set(obj) - {f.alias for f in attrs.fields(cls)} - {f.name.lstrip('_')
                                                                    for f in attrs.fields(cls)}
        if extra:
            raise TypeError(f'extra keys {extra} for {cls}')
        return cls(**kwargs)
    if isinstance(cls, type) and issubclass(cls, enum.Enum) and not isinstance(obj, cls):
        return cls(obj)
    return obj


def _install_stubs():
    class _Anything:
        def __init__(self, *a, **k):
            pass

        def __call__(self, *a, **k):
            return _Anything()

        def __getattr__(self, name):
            return _Anything()

    if 'iolite' not in sys.modules:
        _mod('iolite', file=_Anything(), folder=_Anything(), read_json=_Anything(),
             write_json=_Anything())
    if 'cattrs' not in sys.modules:
        class _Converter:
            def __init__(self, *a, **k):
                pass

            def register_structure_hook(self, *a, **k):
                pass

            def register_unstructure_hook_factory(self, *a, **k):
                pass

            def structure(self, obj, cls):
                return _structure(obj, cls)

        errors = _mod('cattrs.errors', ClassValidationError=type('ClassValidationError',
                                                                 (Exception,), {}))
        gen = _mod('cattrs.gen', make_dict_unstructure_fn=_Anything())
        _mod('cattrs', GenConverter=_Converter, Converter=_Converter, errors=errors, gen=gen,
             override=_Anything())
    if 'shapely' not in sys.modules:
        geometry = _mod('shapely.geometry', Polygon=_Anything, MultiPolygon=_Anything,
                        GeometryCollection=_Anything, box=_Anything(), CAP_STYLE=_Anything(),
                        JOIN_STYLE=_Anything())
        strtree = _mod('shapely.strtree', STRtree=_Anything)
        validation = _mod('shapely.validation', make_valid=_Anything())
        ops = _mod('shapely.ops', unary_union=_Anything())
        _mod('shapely', geometry=geometry, strtree=strtree, validation=validation, ops=ops)
    if 'pyclipper' not in sys.modules:
        _mod('pyclipper')


def load():
    """Import and return the reference `vkit` package (raises if absent)."""
    if not available():
        raise RuntimeError(f'reference tree not found at {REFERENCE_ROOT}')
    os.environ.setdefault('DISABLE_VKIT_COLLECT_USAGE_INFORMATION', '1')
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import vkit  # noqa: F401
    import vkit.element  # noqa: F401
    import vkit.mechanism.distortion  # noqa: F401
    import vkit.mechanism.distortion_policy  # noqa: F401
    return vkit
