"""The reference's option carriers as data -- TEST INFRASTRUCTURE ONLY.

Imports the reference's own ``SimulationConfig`` / ``SimulationParams`` / option constants /
``get_registered_variables`` / ``finalize_config`` / ``construct_primitive_state`` /
``get_helper_data`` (through ``oracle/ref_on_numpy.py``; the files under /root/reference are
executed where they lie) and writes what a drop-in has to match to ``tests/golden/reference_api.json``:
field names and defaults of every NamedTuple, the integer option codes, the variable registry of
every (dimensionality, mhd) pair, what ``finalize_config`` fills in, and which inputs it rejects.
``tests/test_api_parity.py`` holds ``astronomix_b200``'s mirrors to it.

Usage (in the build container):  python oracle/gen_reference_api.py
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.ref_on_numpy import load_reference, ref  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "reference_api.json")


def plain(v):
    """NamedTuples -> {"__fields__": [...], field: value}; arrays -> lists; the rest through repr-safe types."""
    if hasattr(v, "_fields"):
        d = {"__fields__": list(v._fields)}
        for f in v._fields:
            d[f] = plain(getattr(v, f))
        return d
    if isinstance(v, np.ndarray):
        return {"__array__": np.asarray(v).view(np.ndarray).tolist()}
    if isinstance(v, (np.floating, np.integer, np.bool_)):
        return v.item()
    if isinstance(v, (bool, int, float, str)) or v is None:
        return v
    if isinstance(v, (tuple, list)):
        return [plain(x) for x in v]
    return {"__repr__": type(v).__name__}


def main():
    load_reference(x64=True)
    sc = ref("option_classes.simulation_config")
    sp = ref("option_classes.simulation_params")
    rv = ref("variable_registry.registered_variables")
    api = {"generator": "oracle/gen_reference_api.py", "reference": "leo1200/astronomix v0.5.1"}
    api["constants"] = {n: getattr(sc, n) for n in sorted(dir(sc)) if n.isupper() and isinstance(getattr(sc, n), int)}
    api["SimulationConfig"] = plain(sc.SimulationConfig())
    api["SimulationParams"] = plain(sp.SimulationParams())
    api["BoundarySettings"] = plain(sc.BoundarySettings())
    api["BoundarySettings1D"] = plain(sc.BoundarySettings1D())
    api["SnapshotSettings"] = plain(sc.SnapshotSettings())
    # variable registry of every (dimensionality, mhd) pair the reference accepts
    reg = {}
    for dim in (1, 2, 3):
        for mhd in (False, True):
            cfg = sc.SimulationConfig(dimensionality=dim, mhd=mhd, progress_bar=False)
            try:
                reg[f"{dim}d_{'mhd' if mhd else 'hydro'}"] = plain(rv.get_registered_variables(cfg))
            except Exception as e:  # noqa: BLE001 -- the error type is the datum
                reg[f"{dim}d_{'mhd' if mhd else 'hydro'}"] = {"__raises__": type(e).__name__}
    api["registered_variables"] = reg
    # finalize_config: what it fills in, per state shape
    fin = {}
    for name, kw, shape in [
        ("1d_101", dict(dimensionality=1, box_size=1.0), (3, 101)),
        ("2d_64", dict(dimensionality=2, box_size=2.0), (4, 64, 64)),
        ("2d_mhd_32_roll", dict(dimensionality=2, mhd=True, boundary_handling=sc.PERIODIC_ROLL), (7, 32, 32)),
        ("3d_mhd_16", dict(dimensionality=3, mhd=True, box_size=0.5), (8, 16, 16, 16)),
        ("3d_split", dict(dimensionality=3, split=sc.SPLIT, time_integrator=sc.MUSCL), (5, 12, 12, 12)),
        ("2d_noncubic", dict(dimensionality=2), (4, 64, 32)),
        ("3d_noncubic", dict(dimensionality=3), (5, 16, 16, 8)),
    ]:
        cfg = sc.SimulationConfig(progress_bar=False, **kw)
        try:
            out = sc.finalize_config(cfg, shape)
            fin[name] = {"kwargs": {k: plain(v) for k, v in kw.items()}, "shape": list(shape), "config": plain(out)}
        except Exception as e:  # noqa: BLE001
            fin[name] = {"kwargs": {k: plain(v) for k, v in kw.items()}, "shape": list(shape),
                         "raises": type(e).__name__}
    api["finalize_config"] = fin
    # construct_primitive_state / get_helper_data on small grids (Cartesian)
    cps = ref("initial_condition_generation.construct_primitive_state")
    hd = ref("data_classes.simulation_helper_data")
    import jax.numpy as jnp
    rng = np.random.default_rng(3)
    built = {}
    for dim, mhd in ((1, False), (2, False), (2, True), (3, False), (3, True)):
        n = 4
        cfg = sc.SimulationConfig(dimensionality=dim, mhd=mhd, progress_bar=False, box_size=2.0, num_cells=n)
        regs = rv.get_registered_variables(cfg)
        names = ["density", "velocity_x", "velocity_y", "velocity_z", "magnetic_field_x", "magnetic_field_y",
                 "magnetic_field_z", "gas_pressure"]
        fields = {k: rng.random((n,) * dim) for k in names}
        kw = {"density": fields["density"], "gas_pressure": fields["gas_pressure"], "velocity_x": fields["velocity_x"]}
        if dim >= 2:
            kw["velocity_y"] = fields["velocity_y"]
        if dim == 3:
            kw["velocity_z"] = fields["velocity_z"]
        if mhd:
            kw.update({k: fields[k] for k in names if k.startswith("magnetic")})
        state = cps.construct_primitive_state(cfg, regs, **{k: jnp.array(v) for k, v in kw.items()})
        rec = {"fields": {k: v.tolist() for k, v in kw.items()}, "state": np.asarray(state).view(np.ndarray).tolist()}
        for padded in (False, True):
            h = hd.get_helper_data(cfg, padded=padded)
            rec["helper_padded" if padded else "helper"] = {
                f: np.asarray(getattr(h, f)).view(np.ndarray).tolist() for f in h._fields
                if getattr(h, f) is not None and np.asarray(getattr(h, f)).dtype.kind == "f"}
        built[f"{dim}d_{'mhd' if mhd else 'hydro'}"] = rec
    api["construct_and_helpers"] = built
    with open(OUT, "w") as f:
        json.dump(api, f, indent=1, sort_keys=True)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
