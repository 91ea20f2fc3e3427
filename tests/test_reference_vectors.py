"""Golden vectors recorded from the reference's own sources (``oracle/gen_reference_vectors.py``:
the unmodified files of /root/reference/astronomix executed on NumPy through ``oracle/jaxshim``).

CPU leg: the restatement ``oracle/fv_oracle.py`` must reproduce every vector (bit for bit in
almost all cases; the bound below is 1e-13 fp64 / 2e-6 fp32 relative L1 to leave room for
one-ulp differences in how a scalar was rounded).  GPU leg: the CUDA path through the C-ABI
must reproduce them within the north-star tolerance (1e-12 fp64, 1e-5 fp32)."""
import glob
import json
import os

import numpy as np
import pytest

import astronomix_b200 as ax
from astronomix_b200 import option_classes as oc
from astronomix_b200 import problems as pr
from oracle import fv_oracle as fo

HERE = os.path.dirname(os.path.abspath(__file__))
VEC_DIR = os.path.join(HERE, "golden", "reference_vectors")
FILES = sorted(glob.glob(os.path.join(VEC_DIR, "*.npz")))
NAMES = [os.path.basename(f)[:-4] for f in FILES]


def load(name):
    z = np.load(os.path.join(VEC_DIR, name + ".npz"))
    return z["initial_state"], z["final_state"], json.loads(str(z["spec"]))


def build(spec):
    """The reference-facing option classes of this package, filled from the recorded spec."""
    b1 = [oc.BoundarySettings1D(l, r) for l, r in spec["boundaries"]]
    bs = b1[0] if spec["dim"] == 1 else oc.BoundarySettings(*(b1 + [b1[-1]] * (3 - len(b1))))
    cfg = dict(spec["cfg"])
    cfg.setdefault("numerical_precision", oc.SINGLE_PRECISION)  # the reference's default
    params = ax.SimulationParams(t_end=spec["t_end"], **spec["params"])
    dtype = np.float64 if spec["dtype"] == "f64" else np.float32
    return pr.make(spec["dim"], spec["n"], mhd=spec["mhd"], box=spec["box"], boundaries=bs, dtype=dtype,
                   params=params, **cfg)


def test_vectors_present():
    with open(os.path.join(VEC_DIR, "INDEX.json")) as f:
        index = json.load(f)
    assert sorted(index["cases"]) == NAMES and len(NAMES) >= 100


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_reference_vector(name):
    s0, ref, spec = load(name)
    config, params, rv = build(spec)
    out = fo.time_integration(s0, fo.Opts.from_config(config, params))
    assert out.dtype == ref.dtype and out.shape == ref.shape
    err = pr.rel_l1(out, ref, spec["dim"], spec["mhd"])
    assert err <= (1e-13 if spec["dtype"] == "f64" else 2e-6), f"{name}: rel-L1 {err:.3e}"


# HYBRID_HLLC / AM_HLLC / VAN_ALBADA_PP run on the generic kernels only in unsplit mode; everything
# recorded here is inside the C-ABI's supported set.
# vectors whose CUDA leg has not been on hardware yet run from tests/test_z_full_size_properties.py (which
# sorts last), so that under -x they cannot hide a test that has
NOT_YET_ON_HARDWARE = tuple(n for n in NAMES if "_split_hybrid" in n or "_split_amhllc" in n
                            or "_hybrid_valbadapp" in n or "_amhllc_valbadapp" in n)


@pytest.mark.gpu
@pytest.mark.parametrize("name", [n for n in NAMES if n not in NOT_YET_ON_HARDWARE])
def test_cuda_reproduces_reference_vector(cuda, name):
    check_cuda_vector(name)


def check_cuda_vector(name):
    from astronomix_b200 import stepping as ti
    s0, ref, spec = load(name)
    config, params, rv = build(spec)
    out = ax.time_integration(s0, config, params, rv)
    out = np.asarray(out)
    assert out.dtype == ref.dtype and out.shape == ref.shape and np.isfinite(out).all()
    err = pr.rel_l1(out, ref, spec["dim"], spec["mhd"])
    assert err <= (1e-12 if spec["dtype"] == "f64" else 1e-5), f"{name}: rel-L1 {err:.3e}"


# ------------------------------------------------------------------------------------------------
# Rows next to the step, recorded by oracle/gen_reference_aux_vectors.py from
# the reference's own time_integration: SnapshotData of return_snapshots runs, and driven-turbulence
# runs together with the noise the reference drew (the stand-in's random stream is not JAX's, so the
# noise is an input here, not something to reproduce).
AUX_DIR = os.path.join(HERE, "golden", "reference_vectors_aux")
SNAP_NAMES = sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(AUX_DIR, "snap_*.npz")))
# use_specific_snapshot_timepoints (time_integration.py:542-554, :671-678, :742-750)
SNAP_NAMES += sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(AUX_DIR, "snaptp_*.npz")))
FORCED_NAMES = sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(AUX_DIR, "forced_*.npz")))
SERIES = ("total_mass", "total_energy", "internal_energy", "kinetic_energy")


def load_aux(name):
    z = np.load(os.path.join(AUX_DIR, name + ".npz"))
    return z, json.loads(str(z["spec"]))


def build_snap(spec):
    dim, mhd = spec["dim"], spec["mhd"]
    dtype = np.float64 if spec["dtype"] == "f64" else np.float32
    st = ax.SnapshotSettings(return_states=True, return_final_state=True, return_total_mass=True,
                             return_total_energy=True, return_internal_energy=True, return_kinetic_energy=True,
                             return_radial_momentum=dim > 1, return_magnetic_divergence=mhd and dim == 3)
    bnd = {0: "open", 1: "reflective", 2: "periodic"}[spec["boundary"]]
    tps = spec.get("snapshot_timepoints")
    config, params, rv = pr.make(dim, spec["n"], mhd=mhd, dtype=dtype, boundaries=bnd, return_snapshots=True,
                                 num_snapshots=spec["num_snapshots"], snapshot_settings=st,
                                 use_specific_snapshot_timepoints=tps is not None,
                                 numerical_precision=oc.SINGLE_PRECISION)  # the reference's default
    if tps is not None:
        params = params._replace(snapshot_timepoints=tuple(tps))
    return config, params._replace(t_end=spec["t_end"]), rv, dtype


def test_aux_vectors_present():
    assert len(SNAP_NAMES) >= 9 and len(FORCED_NAMES) >= 2


@pytest.mark.parametrize("name", SNAP_NAMES)
def test_oracle_reproduces_reference_snapshots(name):
    z, spec = load_aux(name)
    config, params, rv, dtype = build_snap(spec)
    dim, mhd, ns = spec["dim"], spec["mhd"], spec["num_snapshots"]
    centers = ax.get_helper_data(config, dtype=dtype).geometric_centers if dim > 1 else None
    ref, rec, steps = fo.time_integration_snapshots(z["initial_state"], fo.Opts.from_config(config, params), ns,
                                                    box_size=config.box_size, centers=centers,
                                                    timepoints=spec.get("snapshot_timepoints"))
    tol = 1e-13 if dtype == np.float64 else 2e-6
    assert steps == int(z["num_iterations"]) and rec["current_checkpoint"] == int(z["current_checkpoint"])
    np.testing.assert_allclose(rec["time_points"], z["time_points"], rtol=tol, atol=0)
    assert pr.rel_l1(ref, z["final_state"], dim, mhd) <= tol
    for k in range(min(ns, int(z["current_checkpoint"]))):
        assert pr.rel_l1(rec["states"][k], z["states"][k], dim, mhd) <= tol
    for f in SERIES + (("radial_momentum",) if dim > 1 else ()) + (("magnetic_divergence",) if mhd and dim == 3 else ()):
        scale = max(np.abs(z[f]).max(), np.abs(z["kinetic_energy"]).max())
        np.testing.assert_allclose(rec[f], z[f], rtol=tol * 10, atol=tol * 10 * scale)


@pytest.mark.parametrize("name", FORCED_NAMES)
def test_oracle_reproduces_reference_driven_turbulence(name):
    z, spec = load_aux(name)
    dtype = np.float64 if spec["dtype"] == "f64" else np.float32
    tf = oc.TurbulentForcingConfig(turbulent_forcing=True)
    config, params, rv = pr.make(3, spec["n"], mhd=True, dtype=dtype, boundary_handling=oc.PERIODIC_ROLL,
                                 turbulent_forcing_config=tf, numerical_precision=oc.SINGLE_PRECISION)
    params = params._replace(t_end=spec["t_end"])
    assert params.turbulent_forcing_params.energy_injection_rate == spec["energy_injection_rate"]
    out, steps, time = fo.time_integration_forced(z["initial_state"], fo.Opts.from_config(config, params),
                                                  list(z["noise"]), spec["energy_injection_rate"], config.box_size)
    assert steps == spec["steps"]
    assert pr.rel_l1(out, z["final_state"], 3, True) <= (1e-13 if dtype == np.float64 else 2e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("name", SNAP_NAMES)
def test_cuda_reproduces_reference_snapshots(cuda, name):
    z, spec = load_aux(name)
    config, params, rv, dtype = build_snap(spec)
    dim, mhd, ns = spec["dim"], spec["mhd"], spec["num_snapshots"]
    data = ax.time_integration(z["initial_state"], config, params, rv)
    assert isinstance(data, ax.SnapshotData)
    tol = 1e-12 if dtype == np.float64 else 1e-5  # the north-star bars (BASELINE.json)
    assert data.num_iterations == int(z["num_iterations"]) and data.current_checkpoint == int(z["current_checkpoint"])
    np.testing.assert_allclose(data.time_points, z["time_points"], rtol=tol, atol=0)
    assert pr.rel_l1(data.final_state, z["final_state"], dim, mhd) <= tol
    for k in range(min(ns, int(z["current_checkpoint"]))):
        assert pr.rel_l1(data.states[k], z["states"][k], dim, mhd) <= tol
    for f in SERIES:
        np.testing.assert_allclose(getattr(data, f), z[f], rtol=max(tol, 1e-5 if dtype == np.float32 else 0))
    if dim > 1:
        scale = np.abs(z["kinetic_energy"]).max()
        np.testing.assert_allclose(data.radial_momentum, z["radial_momentum"], atol=1e-4 * scale, rtol=1e-4)
    if mhd and dim == 3:
        np.testing.assert_allclose(data.magnetic_divergence, z["magnetic_divergence"], rtol=1e-3,
                                   atol=1e-5 * np.abs(z["initial_state"][-3:]).max() * spec["n"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", FORCED_NAMES)
def test_cuda_reproduces_reference_driven_turbulence(cuda, name):
    import torch
    from astronomix_b200 import forcing as fm
    from astronomix_b200 import stepping as ti
    z, spec = load_aux(name)
    dtype = np.float64 if spec["dtype"] == "f64" else np.float32
    tf = oc.TurbulentForcingConfig(turbulent_forcing=True)
    config, params, rv = pr.make(3, spec["n"], mhd=True, dtype=dtype, boundary_handling=oc.PERIODIC_ROLL,
                                 turbulent_forcing_config=tf, numerical_precision=oc.SINGLE_PRECISION)
    params = params._replace(t_end=spec["t_end"])
    noises = [torch.from_numpy(x).to(cuda) for x in z["noise"]]
    final, _, info = fm.integrate_per_step(torch.from_numpy(z["initial_state"]).to(cuda), "torch_cuda", config, params,
                                           rv, None, ti._from_device, noise_fn=lambda step: noises[step].clone())
    assert info[0] == spec["steps"]
    err = pr.rel_l1(final.cpu().numpy(), z["final_state"], 3, True)
    assert err <= (1e-12 if dtype == np.float64 else 1e-5), err  # the north-star bars (BASELINE.json)
