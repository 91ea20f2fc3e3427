"""Array type and global flags of the NumPy-backed jax stand-in (see ../README.md).

TEST INFRASTRUCTURE ONLY.  ``Array`` is ``numpy.ndarray`` plus the three behaviours of
``jax.Array`` that astronomix's finite-volume path relies on:

* functional updates ``x.at[idx].set/add/multiply/min/max(v)`` (a copy is updated);
* out-of-range *integer* indices never raise: reads clamp, scatter updates are dropped
  (astronomix relies on this in ``_geometry/boundaries.py:46-69`` when a reflective
  boundary is applied to a 3-row vector field with ``var_index == 3``);
* ``block_until_ready()``.

Everything else (elementwise IEEE arithmetic, ``roll``, ``where``, reductions) is NumPy's.
"""
from __future__ import annotations

import operator

import numpy as np

_x64 = [True]


def default_float():
    return np.float64 if _x64[0] else np.float32


def default_int():
    return np.int64 if _x64[0] else np.int32


def _has_oob_int(idx, shape):
    """True if a static integer of the index tuple is out of range for its axis."""
    if not isinstance(idx, tuple):
        idx = (idx,)
    n_specified = sum(1 for i in idx if i is not None and i is not Ellipsis)
    ax = 0
    for i in idx:
        if i is None:
            continue
        if i is Ellipsis:
            ax += len(shape) - n_specified
            continue
        if isinstance(i, (int, np.integer)) and not isinstance(i, (bool, np.bool_)):
            if ax < len(shape) and not (-shape[ax] <= i < shape[ax]):
                return True
        ax += 1
    return False


def _clamp_ints(idx, shape):
    single = not isinstance(idx, tuple)
    if single:
        idx = (idx,)
    n_specified = sum(1 for i in idx if i is not None and i is not Ellipsis)
    out, ax = [], 0
    for i in idx:
        if i is None:
            out.append(i)
            continue
        if i is Ellipsis:
            out.append(i)
            ax += len(shape) - n_specified
            continue
        if isinstance(i, (int, np.integer)) and not isinstance(i, (bool, np.bool_)) and ax < len(shape):
            n = shape[ax]
            if i >= n:
                i = n - 1
            elif i < -n:
                i = 0
        out.append(i)
        ax += 1
    return out[0] if single else tuple(out)


class _AtIndexer:
    __slots__ = ("_arr",)

    def __init__(self, arr):
        self._arr = arr

    def __getitem__(self, idx):
        return _AtRef(self._arr, idx)


class _AtRef:
    __slots__ = ("_arr", "_idx")

    def __init__(self, arr, idx):
        self._arr = arr
        self._idx = idx

    def _update(self, fn, value):
        out = np.array(self._arr, copy=True, subok=False)
        if _has_oob_int(self._idx, out.shape):
            return out.view(Array)  # jax drops out-of-bounds scatter updates
        value = np.asarray(value)
        if fn is None:
            out[self._idx] = value
        else:
            out[self._idx] = fn(out[self._idx], value)
        return out.view(Array)

    def set(self, value, **_kw):
        return self._update(None, value)

    def add(self, value, **_kw):
        return self._update(operator.add, value)

    def multiply(self, value, **_kw):
        return self._update(operator.mul, value)

    def divide(self, value, **_kw):
        return self._update(operator.truediv, value)

    def min(self, value, **_kw):
        return self._update(np.minimum, value)

    def max(self, value, **_kw):
        return self._update(np.maximum, value)

    def get(self, **_kw):
        return self._arr[self._idx]


class Array(np.ndarray):
    """numpy.ndarray with jax.Array's ``.at`` and index clamping."""

    @property
    def at(self):
        return _AtIndexer(self)

    def __getitem__(self, idx):
        try:
            out = np.ndarray.__getitem__(self, idx)
        except IndexError:
            if not _has_oob_int(idx, self.shape):
                raise
            out = np.ndarray.__getitem__(self, _clamp_ints(idx, self.shape))
        if isinstance(out, np.ndarray):
            return out
        return np.asarray(out).view(Array)  # keep 0-d results array-like

    def __setitem__(self, idx, value):  # jax arrays are immutable
        raise TypeError("jax arrays are immutable; use x.at[idx].set(value)")

    def block_until_ready(self):
        return self

    def __hash__(self):
        raise TypeError("unhashable type: Array")


def wrap(x):
    """Give every ndarray / NumPy scalar that leaves a jnp function the Array type."""
    if isinstance(x, Array):
        return x
    if isinstance(x, np.ndarray):
        return x.view(Array)
    if isinstance(x, np.generic):
        return np.asarray(x).view(Array)
    if isinstance(x, tuple) and type(x) is tuple:
        return tuple(wrap(v) for v in x)
    if isinstance(x, list):
        return [wrap(v) for v in x]
    return x


def mutable(x):
    """Plain-ndarray view for code inside the shim that writes in place."""
    return np.asarray(x).view(np.ndarray)
