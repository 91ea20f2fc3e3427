"""``jax.lax`` control flow and slicing as plain Python / NumPy (TEST INFRASTRUCTURE ONLY)."""
from __future__ import annotations

import numpy as _np

from ._core import wrap as _wrap


def cond(pred, true_fun, false_fun, *operands, **_kw):
    return true_fun(*operands) if bool(pred) else false_fun(*operands)


def switch(index, branches, *operands):
    i = min(max(int(index), 0), len(branches) - 1)
    return branches[i](*operands)


def while_loop(cond_fun, body_fun, init_val):
    val = init_val
    while bool(cond_fun(val)):
        val = body_fun(val)
    return val


def fori_loop(lower, upper, body_fun, init_val, **_kw):
    val = init_val
    for i in range(int(lower), int(upper)):
        val = body_fun(i, val)
    return val


def scan(f, init, xs=None, length=None, **_kw):
    carry, ys = init, []
    n = length if xs is None else len(xs)
    for i in range(n):
        carry, y = f(carry, None if xs is None else xs[i])
        ys.append(y)
    return carry, ys


def stop_gradient(x):
    return x


def slice_in_dim(operand, start_index, limit_index, stride=1, axis=0):
    idx = [slice(None)] * _np.ndim(operand)
    idx[axis] = slice(start_index, limit_index, stride)
    return _wrap(_np.asarray(operand)[tuple(idx)])


def concatenate(operands, dimension):
    return _wrap(_np.concatenate([_np.asarray(o) for o in operands], axis=dimension))


def with_sharding_constraint(x, *_a, **_kw):
    return x
