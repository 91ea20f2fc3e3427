"""Golden vectors for the rows next to the step from the reference's OWN
code -- TEST INFRASTRUCTURE ONLY.  Same mechanism as ``gen_reference_vectors.py`` (the unmodified
sources under /root/reference/astronomix executed over ``oracle/jaxshim``):

* ``snap_*.npz``   ``time_integration(..., return_snapshots=True)``: every field of the reference's
  ``SnapshotData`` (time points, states, totals, max |div B|, num_iterations, current_checkpoint).
* ``forced_*.npz`` ``time_integration`` with ``turbulent_forcing``: final state plus the complex
  noise fields the reference drew, step by step.  The stand-in's ``jax.random`` is NumPy's
  generator, not Threefry, so the noise is recorded and re-injected: the vectors pin
  ``_create_forcing_field`` / ``_apply_forcing`` and the loop order (dt from the unforced state,
  forcing, step), not the random stream.

Usage (in the build container):  python oracle/gen_reference_aux_vectors.py
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.ref_on_numpy import load_reference, ref  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "reference_vectors_aux")

SNAP_CASES = [
    # name, dim, n, mhd, dtype, t_end, num_snapshots, boundary code (0 open, 1 reflective, 2 periodic), seed
    ("snap_mhd3d_f64", 3, 16, True, "f64", 0.03, 4, 2, 5),
    ("snap_mhd2d_f64", 2, 32, True, "f64", 0.05, 5, 2, 5),
    ("snap_sod101_f64", 1, 101, False, "f64", 0.2, 7, 0, 0),
    ("snap_hydro2d_f32_sparse", 2, 32, False, "f32", 0.04, 40, 2, 5),
    ("snap_mhd3d_f32", 3, 16, True, "f32", 0.03, 4, 2, 6),
    ("snap_mhd3d_reflective_f64", 3, 12, True, "f64", 0.06, 4, 1, 110),
]
SNAPTP_CASES = [
    # use_specific_snapshot_timepoints: name, dim, n, mhd, dtype, time points (last = t_end), boundary code, seed
    ("snaptp_mhd3d_f64", 3, 16, True, "f64", [0.0, 0.01, 0.017, 0.03], 2, 5),
    ("snaptp_hydro2d_f32", 2, 32, False, "f32", [0.0, 0.013, 0.04], 2, 5),
    ("snaptp_mhd2d_f64_late_start", 2, 32, True, "f64", [0.004, 0.02, 0.05], 2, 7),
]
FORCED_CASES = [
    # name, n, dtype, t_end, seed
    ("forced_mhd3d_f64", 12, "f64", 0.025, 12),
    ("forced_mhd3d_f32", 12, "f32", 0.025, 13),
]


def _params(sp, wd, t_end, **extra):
    base = sp.SimulationParams()
    pk = dict(extra)
    for f in ("C_cfl", "gamma", "dt_max", "gravitational_constant", "minimum_density", "minimum_pressure"):
        pk[f] = wd(getattr(base, f))  # jit canonicalises params leaves to the working dtype (Q20)
    pk["t_end"] = wd(t_end)
    return sp.SimulationParams(**pk)


def run_snapshots(name, dim, n, mhd, dt, t_end, ns, bcode, seed, timepoints=None):
    from astronomix_b200 import problems as pr
    x64 = dt == "f64"
    wd = np.float64 if x64 else np.float32
    load_reference(x64=x64)
    import jax.numpy as jnp
    sc = ref("option_classes.simulation_config")
    sp = ref("option_classes.simulation_params")
    rv = ref("variable_registry.registered_variables")
    ti = ref("time_stepping.time_integration")
    s0 = pr.sod(n, wd) if dim == 1 else pr.smooth_random(dim, n, mhd=mhd, seed=seed, dtype=wd)
    b1 = sc.BoundarySettings1D(left_boundary=bcode, right_boundary=bcode)
    st = sc.SnapshotSettings(return_states=True, return_final_state=True, return_total_mass=True,
                             return_total_energy=True, return_internal_energy=True, return_kinetic_energy=True,
                             return_radial_momentum=dim > 1, return_magnetic_divergence=mhd and dim == 3)
    config = sc.SimulationConfig(progress_bar=False, dimensionality=dim, mhd=mhd, box_size=1.0, num_cells=n,
                                 boundary_settings=b1 if dim == 1 else sc.BoundarySettings(b1, b1, b1),
                                 return_snapshots=True, num_snapshots=ns, snapshot_settings=st,
                                 use_specific_snapshot_timepoints=timepoints is not None)
    regs = rv.get_registered_variables(config)
    config = sc.finalize_config(config, s0.shape)
    extra = {} if timepoints is None else {"snapshot_timepoints": jnp.array(np.array(timepoints, dtype=wd))}
    out = ti.time_integration(jnp.array(s0), config, _params(sp, wd, t_end, **extra), regs)
    rec = {"initial_state": s0, "final_state": np.asarray(out.final_state).view(np.ndarray),
           "num_iterations": np.int64(out.num_iterations), "current_checkpoint": np.int64(out.current_checkpoint)}
    for f in ("time_points", "states", "total_mass", "total_energy", "internal_energy", "kinetic_energy",
              "radial_momentum", "magnetic_divergence"):
        v = getattr(out, f)
        if v is not None:
            rec[f] = np.asarray(v).view(np.ndarray)
    for k, v in rec.items():
        if isinstance(v, np.ndarray) and v.dtype.kind == "f":
            assert v.dtype == wd, (name, k, v.dtype)
    spec = dict(name=name, dim=dim, n=n, mhd=mhd, dtype=dt, t_end=t_end, num_snapshots=ns, boundary=bcode, box=1.0)
    if timepoints is not None:
        spec["snapshot_timepoints"] = [float(v) for v in timepoints]
    np.savez_compressed(os.path.join(OUT, name + ".npz"), spec=np.array(json.dumps(spec)), **rec)
    print(f"{name:32s} steps={int(out.num_iterations)} snapshots taken={int(out.current_checkpoint)}", flush=True)


def run_forced(name, n, dt, t_end, seed):
    from astronomix_b200 import problems as pr
    x64 = dt == "f64"
    wd = np.float64 if x64 else np.float32
    load_reference(x64=x64)
    import jax
    import jax.numpy as jnp
    sc = ref("option_classes.simulation_config")
    sp = ref("option_classes.simulation_params")
    rv = ref("variable_registry.registered_variables")
    ti = ref("time_stepping.time_integration")
    s0 = pr.smooth_random(3, n, mhd=True, seed=seed, dtype=wd)
    b1 = sc.BoundarySettings1D(left_boundary=2, right_boundary=2)
    config = sc.SimulationConfig(progress_bar=False, dimensionality=3, mhd=True, box_size=1.0, num_cells=n,
                                 boundary_settings=sc.BoundarySettings(b1, b1, b1), boundary_handling=sc.PERIODIC_ROLL,
                                 turbulent_forcing_config=sc.TurbulentForcingConfig(turbulent_forcing=True))
    regs = rv.get_registered_variables(config)
    config = sc.finalize_config(config, s0.shape)
    params = _params(sp, wd, t_end)
    edot = float(params.turbulent_forcing_params.energy_injection_rate)
    draws = []
    real_normal = jax.random.normal

    def recording_normal(k, shape=(), dtype=None):
        x = real_normal(k, shape, dtype)
        draws.append(np.asarray(x).view(np.ndarray).copy())
        return x

    jax.random.normal = recording_normal
    try:
        out = ti.time_integration(jnp.array(s0), config, params, regs)
    finally:
        jax.random.normal = real_normal
    out = np.asarray(out).view(np.ndarray)
    assert out.dtype == wd and len(draws) % 2 == 0 and np.isfinite(out).all()
    # _turbulent_forcing.py:50-51: real part from the first draw of a step, imaginary part from the second
    noise = np.stack([draws[2 * k] + 1j * draws[2 * k + 1] for k in range(len(draws) // 2)])
    assert noise.dtype == (np.complex128 if x64 else np.complex64), noise.dtype
    spec = dict(name=name, n=n, dtype=dt, t_end=t_end, box=1.0, energy_injection_rate=edot, steps=len(noise))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), spec=np.array(json.dumps(spec)), initial_state=s0,
                        final_state=out, noise=noise)
    print(f"{name:32s} steps={len(noise)}", flush=True)


def main():
    os.makedirs(OUT, exist_ok=True)
    for c in SNAP_CASES:
        run_snapshots(*c)
    for name, dim, n, mhd, dt, tps, bcode, seed in SNAPTP_CASES:
        run_snapshots(name, dim, n, mhd, dt, tps[-1], len(tps), bcode, seed, timepoints=tps)
    for c in FORCED_CASES:
        run_forced(*c)
    with open(os.path.join(OUT, "INDEX.json"), "w") as f:
        json.dump({"generator": "oracle/gen_reference_aux_vectors.py",
                   "reference": "leo1200/astronomix v0.5.1 sources executed over oracle/jaxshim (NumPy backend)",
                   "cases": [c[0] for c in SNAP_CASES + SNAPTP_CASES + FORCED_CASES]}, f, indent=1)


if __name__ == "__main__":
    main()
