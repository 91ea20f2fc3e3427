"""``jax.numpy`` stand-in: NumPy functions with jax's default dtypes and the ``Array`` result type.

TEST INFRASTRUCTURE ONLY (see ../../README.md).  Functions are resolved lazily from NumPy and
wrapped: Array arguments are passed as plain ndarray views, results come back as ``Array``.
Only creation routines differ from NumPy (default float = float64 with x64 on, else float32,
as ``jax.config.jax_enable_x64`` decides in JAX).
"""
from __future__ import annotations

import sys as _sys
import types as _types

import numpy as _np

from .._core import Array as _Array, default_float as _dflt, default_int as _dint, wrap as _wrap

ndarray = _Array
pi = _np.pi
e = _np.e
inf = _np.inf
nan = _np.nan
newaxis = None
float32, float64 = _np.float32, _np.float64
int32, int64 = _np.int32, _np.int64
bool_ = _np.bool_
complex64, complex128 = _np.complex64, _np.complex128
int = int  # the reference uses jnp.int in one annotation


def _plain(x):
    if isinstance(x, _Array):
        return x.view(_np.ndarray)
    if isinstance(x, (list, tuple)) and any(isinstance(v, (_Array, list, tuple)) for v in x):
        return type(x)(_plain(v) for v in x) if type(x) in (list, tuple) else x
    return x


def _lift(fn):
    def wrapped(*args, **kwargs):
        args = tuple(_plain(a) for a in args)
        kwargs = {k: _plain(v) for k, v in kwargs.items()}
        return _wrap(_canon_out(fn(*args, **kwargs)))
    wrapped.__name__ = getattr(fn, "__name__", "fn")
    return wrapped


def _float_default(dtype):
    return _dflt() if dtype is None else dtype


def _canon(a):
    """jax never keeps 64-bit types when x64 is off."""
    if _dflt() is _np.float32:
        if a.dtype == _np.float64:
            return a.astype(_np.float32)
        if a.dtype == _np.int64:
            return a.astype(_np.int32)
        if a.dtype == _np.complex128:
            return a.astype(_np.complex64)
    return a


def _canon_out(x):
    if isinstance(x, (_np.ndarray, _np.generic)):
        return _canon(_np.asarray(x))
    if type(x) in (tuple, list):
        return type(x)(_canon_out(v) for v in x)
    return x


def array(obj, dtype=None, copy=True, **_kw):
    a = _np.array(_plain(obj), dtype=dtype, copy=True)
    if dtype is None and not isinstance(obj, _np.ndarray):
        a = _canon(a)
    return a.view(_Array)


def asarray(obj, dtype=None, **_kw):
    return array(obj, dtype=dtype)


def zeros(shape, dtype=None):
    return _np.zeros(shape, _float_default(dtype)).view(_Array)


def ones(shape, dtype=None):
    return _np.ones(shape, _float_default(dtype)).view(_Array)


def full(shape, fill_value, dtype=None):
    if dtype is None:
        dtype = _canon(_np.asarray(_plain(fill_value))).dtype
    return _np.full(shape, _plain(fill_value), dtype).view(_Array)


def empty(shape, dtype=None):
    return zeros(shape, dtype)


def arange(*args, dtype=None, **kw):
    return _canon(_np.arange(*args, dtype=dtype, **kw)).view(_Array)


def linspace(start, stop, num=50, endpoint=True, retstep=False, dtype=None, axis=0):
    out = _np.linspace(_plain(start), _plain(stop), num, endpoint=endpoint, retstep=retstep, axis=axis)
    if retstep:
        return _canon(out[0]).astype(_float_default(dtype)).view(_Array), out[1]
    return out.astype(_float_default(dtype)).view(_Array)


def eye(n, m=None, k=0, dtype=None):
    return _np.eye(n, m, k, _float_default(dtype)).view(_Array)


def meshgrid(*xi, **kw):
    return [_wrap(a.copy()) for a in _np.meshgrid(*[_plain(x) for x in xi], **kw)]


def where(cond, x=None, y=None, **_kw):
    if x is None:
        return _wrap(_np.where(_plain(cond)))
    return _wrap(_np.where(_plain(cond), _plain(x), _plain(y)))


def vectorize(pyfunc, **kw):
    return _lift(_np.vectorize(pyfunc, **kw))


def _module(name, src):
    m = _types.ModuleType(__name__ + "." + name)

    def _getattr(attr, _src=src):
        if attr.startswith("__"):
            raise AttributeError(attr)
        return _lift(getattr(_src, attr))
    m.__getattr__ = _getattr
    _sys.modules[m.__name__] = m
    return m


linalg = _module("linalg", _np.linalg)
fft = _module("fft", _np.fft)


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    try:
        fn = getattr(_np, name)
    except AttributeError:
        raise AttributeError(f"jax.numpy stand-in has no attribute {name!r}") from None
    if callable(fn) and not isinstance(fn, type):
        fn = _lift(fn)
    globals()[name] = fn
    return fn
