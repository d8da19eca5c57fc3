"""Name -> class table of the accelerated algorithms.

Same surface as the reference's registry (omnisafe/algorithms/registry.py:L25-78): `@registry.register`
on a class, `registry.get(name)` to look it up, `REGISTRY.name`; duplicate names and non-classes are
rejected with the reference's exception types.
"""
from __future__ import annotations

_TABLE_NAME = 'OmniSafe-B200'
_classes: dict[str, type] = {}


def register(cls: type) -> type:
    """Class decorator: file `cls` under its own name."""
    if not isinstance(cls, type):
        raise TypeError(f'module must be a class, but got {type(cls)}')
    key = cls.__name__
    if key in _classes:
        raise KeyError(f'{key} is already registered in {_TABLE_NAME}')
    _classes[key] = cls
    return cls


def get(name: str) -> type:
    try:
        return _classes[name]
    except KeyError:
        raise KeyError(f'{name} is not in the {_TABLE_NAME} registry') from None


class _RegistryView:
    """Object form of the table for code that expects `REGISTRY.register / .get / .name`."""

    name = _TABLE_NAME
    register = staticmethod(register)
    get = staticmethod(get)

    def __contains__(self, name: str) -> bool:
        return name in _classes

    def names(self) -> list[str]:
        return sorted(_classes)


REGISTRY = _RegistryView()
