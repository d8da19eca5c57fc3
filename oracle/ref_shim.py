"""TEST INFRASTRUCTURE ONLY -- import shim for the *unmodified* reference.

Makes `import omnisafe` (from /root/reference) work in the build container, where gymnasium,
safety_gymnasium, matplotlib, ... are not installed.  It is used by
`tests/golden/make_golden.py` to generate the golden fixtures that pin the oracle, and by nothing
else.  /root/reference does not exist on the GPU box, so nothing in `-m gpu` tests, `smoke()` or
`bench.py` may import this module.

The shim stubs the third-party roots the hot path never calls and supplies a tiny real
`gymnasium.spaces.Box/Discrete` (the only gymnasium types the on-policy path touches:
omnisafe/typing.py:L32, omnisafe/envs/core.py).
"""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import sys
import types
from unittest import mock

import numpy as np

REFERENCE_ROOT = '/root/reference'
_STUB_ROOTS = (
    'gymnasium',
    'safety_gymnasium',
    'matplotlib',
    'seaborn',
    'gdown',
    'moviepy',
    'pytorch_lightning',
    'metadrive',
    'isaacgym',
    'gpytorch',
)


class Box:
    """Minimal stand-in for gymnasium.spaces.Box (shape/low/high/dtype/sample)."""

    def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):  # noqa: ARG002
        if shape is None:
            shape = np.shape(low)
        self.shape = tuple(shape)
        self.dtype = np.dtype(dtype)
        self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape).copy()
        self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape).copy()

    def sample(self):
        return np.random.uniform(self.low, self.high).astype(self.dtype)

    def contains(self, x):
        return np.shape(x) == self.shape

    def __repr__(self):
        return f'Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})'


class Discrete:
    def __init__(self, n, seed=None, start=0):  # noqa: ARG002
        self.n = int(n)
        self.start = start
        self.shape = ()
        self.dtype = np.dtype(np.int64)

    def sample(self):
        return int(np.random.randint(self.n)) + self.start


class _StubModule(types.ModuleType):
    """Package-like module: CamelCase attrs -> fresh plain classes, others -> MagicMock."""

    __path__: list = []

    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        val = type(name, (), {}) if name[:1].isupper() else mock.MagicMock(name=name)
        setattr(self, name, val)
        return val


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):  # noqa: ARG002
        if fullname.split('.')[0] in _STUB_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        mod = _StubModule(spec.name)
        if spec.name == 'gymnasium.spaces':
            mod.Box = Box
            mod.Discrete = Discrete
        return mod

    def exec_module(self, module):  # noqa: ARG002
        return None


def install() -> None:
    """Install the stub finder and put the reference on sys.path (idempotent)."""
    if not any(isinstance(f, _StubFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, _StubFinder())
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def import_reference():
    """Return the unmodified reference package."""
    install()
    import omnisafe  # noqa: PLC0415

    return omnisafe
