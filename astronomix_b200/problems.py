"""Seeded synthetic initial conditions and option builders shared by the tests, smoke() and
bench.py.  Everything is NumPy; nothing here touches the oracle or CUDA."""
from __future__ import annotations

import numpy as np

from . import option_classes as _oc_unused  # noqa: F401
import astronomix_b200 as ax
from astronomix_b200 import option_classes as oc

PERIODIC = oc.PERIODIC_BOUNDARIES


def make(dim, n, mhd=False, box=1.0, boundaries="periodic", dtype=np.float32, params=None, **cfg):
    """(config, params, registered_variables) finalized for an n^dim grid."""
    if boundaries == "periodic":
        bs = PERIODIC if dim > 1 else oc.BoundarySettings1D(oc.PERIODIC_BOUNDARY, oc.PERIODIC_BOUNDARY)
    elif boundaries == "open":
        bs = oc.BoundarySettings() if dim > 1 else oc.BoundarySettings1D()
    elif boundaries == "reflective":
        r = oc.BoundarySettings1D(oc.REFLECTIVE_BOUNDARY, oc.REFLECTIVE_BOUNDARY)
        bs = oc.BoundarySettings(r, r, r) if dim > 1 else r
    else:
        bs = boundaries
    prec = oc.DOUBLE_PRECISION if np.dtype(dtype) == np.float64 else oc.SINGLE_PRECISION
    cfg.setdefault("numerical_precision", prec)
    config = ax.SimulationConfig(dimensionality=dim, mhd=mhd, box_size=box, num_cells=n,
                                 boundary_settings=bs, **cfg)
    rv = ax.get_registered_variables(config)
    config = ax.finalize_config(config, (rv.num_vars,) + (n,) * dim)
    return config, (params or ax.SimulationParams()), rv


def centers(n, box=1.0):
    dx = box / n
    return (np.arange(n) + 0.5) * dx


def sod(n=101, dtype=np.float32):
    """README hello-world / tests/hydro_tests/shock_tube.py:54-89."""
    x = centers(n)
    s = np.zeros((3, n), dtype=dtype)
    s[0] = np.where(x < 0.5, 1.0, 0.125)
    s[2] = np.where(x < 0.5, 1.0, 0.1)
    return s


def kelvin_helmholtz(n, dtype=np.float32):
    """tests/hydro_tests/kelvin_helmholtz.py:60-74."""
    x = centers(n)
    X, Y = np.meshgrid(x, x, indexing="ij")
    s = np.zeros((4, n, n), dtype=dtype)
    inner = (Y > 0.25) & (Y < 0.75)
    s[0] = np.where(inner, 2.0, 1.0)
    s[1] = np.where(inner, -0.5, 0.5)
    s[2] = 0.01 * np.sin(2 * np.pi * X)
    s[3] = 2.5
    return s


def orszag_tang(n, dtype=np.float64):
    """tests/mhd_tests/orszag_tang.py:39-100 (box 2*pi)."""
    gamma = 5 / 3
    box = 2 * np.pi
    dx = box / n
    x = np.linspace(dx / 2, box - dx / 2, n)
    X, Y = np.meshgrid(x, x, indexing="ij")
    s = np.zeros((7, n, n), dtype=dtype)
    s[0] = gamma ** 2
    s[1] = -np.sin(Y)
    s[2] = np.sin(X)
    s[3] = gamma
    s[4] = -np.sin(Y)
    s[5] = np.sin(2 * X)
    return s


def mhd_blast3d(n, dtype=np.float32, p_in=1000.0, p_out=0.1, B0=100 / np.sqrt(4 * np.pi), r_inj=0.1):
    """tests/mhd_tests/mhd_blast3D.py:47-125 (theta = pi/2, phi = pi/4)."""
    x = centers(n)
    X, Y, Z = np.meshgrid(x, x, x, indexing="ij")
    r2 = (X - 0.5) ** 2 + (Y - 0.5) ** 2 + (Z - 0.5) ** 2
    s = np.zeros((8, n, n, n), dtype=dtype)
    s[0] = 1.0
    s[4] = np.where(r2 < r_inj ** 2, p_in, p_out)
    theta, phi = np.pi / 2, np.pi / 4
    s[5] = B0 * np.sin(theta) * np.cos(phi)
    s[6] = B0 * np.sin(theta) * np.sin(phi)
    s[7] = B0 * np.cos(theta)
    return s


def smooth_random(dim, n, mhd=False, seed=0, dtype=np.float32, amp=0.3):
    """A few random low-k Fourier modes per field: smooth, no symmetry, all limiter branches hit."""
    rng = np.random.default_rng(seed)
    x = centers(n)
    grids = np.meshgrid(*([x] * dim), indexing="ij")
    nv = dim + 2 + (3 if mhd else 0)

    def field():
        f = np.zeros((n,) * dim)
        for _ in range(4):
            k = rng.integers(1, 4, size=dim)
            ph = rng.uniform(0, 2 * np.pi)
            f += rng.uniform(-1, 1) * np.sin(2 * np.pi * sum(ki * g for ki, g in zip(k, grids)) + ph)
        return f / 4

    s = np.zeros((nv,) + (n,) * dim, dtype=dtype)
    s[0] = 1.0 + amp * field()
    for k in range(dim):
        s[1 + k] = amp * 2 * field()
    s[dim + 1] = 1.0 + amp * field()
    if mhd:
        for k in range(3):
            s[dim + 2 + k] = amp * field()
    return s


def rel_l1(a, ref, dim, mhd):
    """Worst relative L1 error over {rho, velocity (all components), p, B (all components)}."""
    groups = [[0], list(range(1, dim + 1)), [dim + 1]]
    if mhd:
        groups.append(list(range(dim + 2, dim + 5)))
    worst = 0.0
    for g in groups:
        num = float(np.sum(np.abs(a[g].astype(np.float64) - ref[g].astype(np.float64))))
        den = float(np.sum(np.abs(ref[g].astype(np.float64))))
        if num == 0.0:
            continue
        worst = max(worst, num / den if den > 0 else np.inf)
    return worst


def totals(s, dim, mhd, gamma=5 / 3):
    """(mass, momentum[dim], gas+magnetic energy) sums in float64."""
    s = s.astype(np.float64)
    rho, p = s[0], s[dim + 1]
    v2 = sum(s[1 + k] ** 2 for k in range(dim))
    E = p / (gamma - 1) + 0.5 * rho * v2
    if mhd:
        E = E + 0.5 * sum(s[dim + 2 + k] ** 2 for k in range(3))
    return rho.sum(), [float((rho * s[1 + k]).sum()) for k in range(dim)], float(E.sum())
