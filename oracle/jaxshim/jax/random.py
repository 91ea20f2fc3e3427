"""``jax.random`` stand-in (TEST INFRASTRUCTURE ONLY).  NumPy's generator, NOT Threefry: streams
differ from JAX's, so nothing random is recorded in the golden vectors."""
from __future__ import annotations

import numpy as _np

from ._core import default_float as _dflt, wrap as _wrap


def key(seed):
    return _np.random.SeedSequence(int(seed))


PRNGKey = key


def split(k, num=2):
    return list(k.spawn(num))


def normal(k, shape=(), dtype=None):
    return _wrap(_np.random.default_rng(k).standard_normal(shape).astype(dtype or _dflt()))


def uniform(k, shape=(), dtype=None, minval=0.0, maxval=1.0):
    return _wrap(_np.random.default_rng(k).uniform(minval, maxval, shape).astype(dtype or _dflt()))
