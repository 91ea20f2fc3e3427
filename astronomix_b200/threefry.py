"""jax's default PRNG for the turbulent driving: keys on the host (integer arithmetic on three words per
step), samples on the device (libamx ``amx_threefry_normal``).

The reference carries ``jax.random.key(42)`` through the time loop (time_integration.py:787-789) and every
forced step does ``key, sk1, sk2 = jax.random.split(key, 3)`` followed by two ``jax.random.normal`` draws of
shape (3, n, n, n) (_turbulent_forcing.py:48-51).  ``PARTITIONABLE`` selects jax's counter layout: True is the
default of jax >= 0.5 (``jax_threefry_partitionable``), False the older default; AMX_THREEFRY_PARTITIONABLE=0
switches.  The Threefry block function is pinned to the Random123 known-answer vectors, the layers above it to
values the jax documentation prints for both layouts (tests/test_threefry.py).
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _lib

PARTITIONABLE = os.environ.get("AMX_THREEFRY_PARTITIONABLE", "1") != "0"
_M = 0xFFFFFFFF
_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))


def _block(k0, k1, x0, x1):
    """Threefry-2x32-20 on Python ints."""
    ks = (k0, k1, k0 ^ k1 ^ 0x1BD11BDA)
    x0 = (x0 + ks[0]) & _M
    x1 = (x1 + ks[1]) & _M
    for blk in range(5):
        for r in _ROT[blk & 1]:
            x0 = (x0 + x1) & _M
            x1 = ((x1 << r) | (x1 >> (32 - r))) & _M
            x1 ^= x0
        x0 = (x0 + ks[(blk + 1) % 3]) & _M
        x1 = (x1 + ks[(blk + 2) % 3] + blk + 1) & _M
    return x0, x1


def key(seed):
    """jax.random.key(seed) as its two raw words."""
    seed = int(seed)
    return ((seed >> 32) & _M, seed & _M)


def split(k, num=2, partitionable=None):
    """jax.random.split(k, num) -> list of raw keys."""
    part = PARTITIONABLE if partitionable is None else partitionable
    if part:
        return [_block(k[0], k[1], 0, i) for i in range(num)]
    # counters 0 .. 2 num - 1: first half in word 0, second half in word 1; outputs concatenated, then paired
    words = [None] * (2 * num)
    for b in range(num):
        words[b], words[b + num] = _block(k[0], k[1], b, b + num)
    return [(words[2 * i], words[2 * i + 1]) for i in range(num)]


def normal(k, shape, dtype=torch.float32, device=None, partitionable=None, out=None):
    """jax.random.normal(k, shape, dtype) as a CUDA tensor."""
    part = PARTITIONABLE if partitionable is None else partitionable
    _lib.require_device()
    if out is None:
        out = torch.empty(tuple(shape), dtype=dtype, device=device if device is not None else "cuda")
    code = _lib.AMX_F32 if out.dtype == torch.float32 else _lib.AMX_F64
    with torch.cuda.device(out.device):
        rc = _lib.load().amx_threefry_normal(int(k[0]), int(k[1]), out.numel(), C.c_void_p(out.data_ptr()), code,
                                             1 if part else 0, C.c_void_p(torch.cuda.current_stream().cuda_stream))
    if rc:
        raise _lib.AmxError(rc, "amx_threefry_normal failed (count beyond the 32-bit counters of the old layout?)")
    return out


def normal_range(k, total, offset, out, partitionable=None):
    """Samples [offset, offset + out.numel()) of jax.random.normal(k, shape) with prod(shape) = total, into the
    contiguous CUDA tensor ``out`` (a slab rank's part of a global draw)."""
    part = PARTITIONABLE if partitionable is None else partitionable
    _lib.require_device()
    assert out.is_contiguous()
    code = _lib.AMX_F32 if out.dtype == torch.float32 else _lib.AMX_F64
    with torch.cuda.device(out.device):
        rc = _lib.load().amx_threefry_normal_range(int(k[0]), int(k[1]), int(total), int(offset), out.numel(),
                                                   C.c_void_p(out.data_ptr()), code, 1 if part else 0,
                                                   C.c_void_p(torch.cuda.current_stream().cuda_stream))
    if rc:
        raise _lib.AmxError(rc, "amx_threefry_normal_range failed")
    return out
