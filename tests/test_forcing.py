"""Turbulent driving: the torch implementation against the NumPy restatement on
identical noise, and the two properties the construction promises (solenoidal field, prescribed
kinetic-energy injection).  CPU tensors here; the GPU loop is in test_gpu_parity.py."""
import numpy as np
import pytest
import torch

import astronomix_b200 as ax
from astronomix_b200 import option_classes as oc, problems as pr
from astronomix_b200.forcing import Forcing
from oracle import fv_oracle as fo
from oracle import threefry as tf


def _noise(n, tdt, seed=42):
    """What forcing.ForcingNoise draws on the device (key(seed) -> split in three -> two normal draws), from the
    NumPy restatement of jax's PRNG"""
    npdt = np.float64 if tdt == torch.float64 else np.float32
    _, sk1, sk2 = tf.split(tf.seed_key(seed), 3)
    return torch.complex(torch.from_numpy(tf.normal(sk1, (3, n, n, n), npdt)),
                         torch.from_numpy(tf.normal(sk2, (3, n, n, n), npdt)))


def _setup(n=16, dtype=np.float64, mhd=True):
    config, params, rv = pr.make(3, n, mhd=mhd, dtype=dtype, boundary_handling=oc.PERIODIC_ROLL,
                                 turbulent_forcing_config=oc.TurbulentForcingConfig(turbulent_forcing=True))
    s = pr.smooth_random(3, n, mhd=mhd, seed=9, dtype=dtype)
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    return config, params, rv, s, tdt


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-12), (np.float32, 2e-5)])
def test_forcing_matches_oracle_on_identical_noise(dtype, tol):
    config, params, rv, s, tdt = _setup(dtype=dtype)
    f = Forcing(config, params, rv, tdt, torch.device("cpu"))
    noise = _noise(16, tdt)
    w = f.field(noise.clone())
    w_ref = fo.forcing_field(noise.numpy().copy(), 16, config.box_size)
    scale = max(np.abs(x).max() for x in w_ref)
    for a, b in zip(w, w_ref):
        assert np.abs(a.numpy() - b).max() <= tol * scale
    dt = 1.0e-3
    forced = f.apply(torch.from_numpy(s.copy()), torch.tensor(dt, dtype=tdt), noise.clone()).numpy()
    o = fo.Opts.from_config(config, params)
    ref = fo.apply_forcing(s, dtype(dt), noise.numpy().copy(), params.turbulent_forcing_params.energy_injection_rate,
                           o, config.box_size)
    assert pr.rel_l1(forced, ref, 3, True) <= tol
    assert np.array_equal(forced[0], s[0]) and np.array_equal(forced[4:], s[4:])  # only the velocity rows change


def test_forcing_field_is_solenoidal_and_injects_the_prescribed_energy():
    config, params, rv, s, tdt = _setup(n=24)
    f = Forcing(config, params, rv, tdt, torch.device("cpu"))
    noise = _noise(24, tdt, seed=3)
    w = f.field(noise.clone())
    div = sum(k * torch.fft.fftn(c) for k, c in zip((f.kx, f.ky, f.kz), w))
    spec = sum(torch.fft.fftn(c).abs().max() for c in w)
    # exact up to the Nyquist planes (taking the real part pairs k with -k, and fftfreq has only -k_Ny),
    # where the spectrum k^6 exp(-8 k / k_peak) is ~1e-8 of its peak
    assert div.abs().max() <= 1e-6 * spec * f.kx.abs().max()
    assert all(abs(float(c.mean())) < 1e-12 * float(c.abs().max()) for c in w)          # no DC mode
    dt = 2.0e-3
    st = torch.from_numpy(s.copy())
    ek0 = 0.5 * (st[0] * (st[1] ** 2 + st[2] ** 2 + st[3] ** 2)).sum() * f.dV
    forced = f.apply(st, torch.tensor(dt, dtype=tdt), noise.clone())
    ek1 = 0.5 * (forced[0] * (forced[1] ** 2 + forced[2] ** 2 + forced[3] ** 2)).sum() * f.dV
    edot = params.turbulent_forcing_params.energy_injection_rate
    assert abs(float(ek1 - ek0) - edot * dt) <= 1e-10 * edot * dt + 1e-14


def test_forcing_needs_periodic_roll_3d():
    config, params, rv = pr.make(3, 16, mhd=True, dtype=np.float64)   # GHOST_CELLS
    with pytest.raises(ValueError):
        Forcing(config, params, rv, torch.float64, torch.device("cpu"))


@pytest.mark.parametrize("dim,n,mhd", [(1, 64, False), (2, 24, True), (3, 12, True)])
def test_snapshot_totals_match_oracle(dim, n, mhd):
    """snapshots.Diagnostics (torch reductions, here on CPU tensors) against the oracle's
    total_quantities (total_quantities.py:31-176, incl. the box_size**dim factor of the total mass)."""
    from astronomix_b200.snapshots import Diagnostics
    bnd = "open" if dim == 1 else "periodic"
    config, params, rv = pr.make(dim, n, mhd=mhd, dtype=np.float64, boundaries=bnd, box=1.5, return_snapshots=True)
    s = pr.sod(n, np.float64) if dim == 1 else pr.smooth_random(dim, n, mhd=mhd, seed=4, dtype=np.float64)
    d = Diagnostics(config, params, rv, torch.float64, torch.device("cpu"))
    centers = ax.get_helper_data(config, dtype=np.float64).geometric_centers
    ref = fo.total_quantities(s, fo.Opts.from_config(config, params), config.box_size, np.float64(params.gamma),
                              centers if dim > 1 else None)
    st = torch.from_numpy(s)
    for name in ("total_mass", "total_energy", "internal_energy", "kinetic_energy"):
        assert abs(float(getattr(d, name)(st)) - ref[name]) <= 1e-12 * abs(ref[name])
    if dim > 1:
        assert abs(float(d.radial_momentum(st)) - ref["radial_momentum"]) <= 1e-12 * ref["kinetic_energy"]
    if mhd and dim == 3:
        assert abs(float(d.magnetic_divergence(st)) - ref["magnetic_divergence"]) <= 1e-12 * abs(ref["magnetic_divergence"])
    if dim > 1:
        assert abs(ref["total_mass"] - s[0].sum() * 1.5 ** dim) <= 1e-12 * ref["total_mass"]
