"""NumPy-backed stand-in for the part of the ``jax`` API that astronomix's finite-volume
path uses -- TEST INFRASTRUCTURE ONLY (see ../README.md).

JAX is not installed in this image and cannot be installed (no network).  This package lets
the UNMODIFIED reference sources under /root/reference/astronomix execute eagerly on NumPy,
so that ``oracle/gen_reference_vectors.py`` can record golden vectors of the reference's own
code.  ``jit`` is the identity, control-flow primitives are Python control flow, arrays are
``numpy.ndarray`` subclasses with ``.at[]``.
"""
from __future__ import annotations

import functools as _functools

import numpy as _np

from . import _core
from ._core import Array  # noqa: F401
from . import numpy  # noqa: F401
from . import lax, random, debug, nn, tree_util  # noqa: F401

__shim__ = True


class _Config:
    def update(self, name, value):
        if name == "jax_enable_x64":
            _core._x64[0] = bool(value)

    @property
    def jax_enable_x64(self):
        return _core._x64[0]


config = _Config()


class _Jitted:
    """``jax.jit`` result: calls the function eagerly; ``lower().compile()`` is a no-op."""

    def __init__(self, fun):
        self._fun = fun
        _functools.update_wrapper(self, fun)

    def __call__(self, *args, **kwargs):
        return self._fun(*args, **kwargs)

    def __get__(self, obj, objtype=None):
        return self if obj is None else _functools.partial(self, obj)

    def lower(self, *args, **kwargs):
        return self

    def compile(self):
        return self

    def memory_analysis(self):
        return None


def jit(fun=None, **_kw):
    if fun is None:
        return lambda f: _Jitted(f)
    return _Jitted(fun)


def checkpoint(fun=None, **_kw):
    if fun is None:
        return lambda f: f
    return fun


remat = checkpoint


def vmap(fun, in_axes=0, out_axes=0, **_kw):
    def mapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = next(_np.shape(a)[ax] for a, ax in zip(args, axes) if ax is not None)
        outs = []
        for i in range(n):
            call = [a if ax is None else _np.take(_np.asarray(a), i, axis=ax).view(Array)
                    for a, ax in zip(args, axes)]
            outs.append(fun(*call))
        if isinstance(outs[0], tuple):
            return tuple(_np.stack([o[k] for o in outs], axis=out_axes).view(Array)
                         for k in range(len(outs[0])))
        return _np.stack(outs, axis=out_axes).view(Array)
    return mapped


def device_put(x, *_a, **_kw):
    return x


def devices(*_a):
    return ["numpy-cpu"]


def block_until_ready(x):
    return x


class NamedSharding:  # only ever used as a type annotation / None on this path
    def __init__(self, *a, **k):
        self.args = a
