"""Golden vectors from the reference's OWN code -- TEST INFRASTRUCTURE ONLY.

Runs the unmodified sources under /root/reference/astronomix (``time_integration`` with the
reference's ``SimulationConfig`` / ``SimulationParams`` / ``get_registered_variables`` /
``finalize_config``) through ``oracle/ref_on_numpy.py`` -- i.e. with ``oracle/jaxshim``
standing in for jax, because jax cannot be installed here -- on seeded NumPy initial states,
and writes one ``.npz`` per case to ``tests/golden/reference_vectors/``:

    initial_state, final_state, spec (JSON: dimensionality, num_cells, mhd, box_size,
    boundaries, SimulationConfig integer options, SimulationParams numbers, dtype)

The restatement in ``oracle/fv_oracle.py`` and the CUDA path are then both checked against
these files (``tests/test_reference_vectors.py``).  What the vectors certify: the reference's
Python (control flow, index arithmetic, operation order, quirks) evaluated with IEEE
elementwise arithmetic in the working dtype.  What they cannot certify: XLA's own instruction
selection (e.g. FMA contraction on the CPU/GPU backend), which differs at round-off level.

Scalar typing.  Under ``jax.jit`` the leaves of ``SimulationParams`` are traced scalars of the
working dtype.  There is no tracer here, so the generator casts those leaves to
the working dtype before the call; ``SimulationConfig`` stays Python numbers (static under jit).

Usage (in the build container):  python oracle/gen_reference_vectors.py [--only SUBSTR]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.ref_on_numpy import load_reference, ref  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "reference_vectors")

# option codes as in option_classes/simulation_config.py:20-93 (asserted against the reference below)
MINMOD, OSHER, DOUBLE_MINMOD, SUPERBEE, VAN_ALBADA, VAN_ALBADA_PP = 0, 1, 2, 3, 4, 5
UNSPLIT, SPLIT = 0, 1
HLL, HLLC, HLLC_LM, LAX_FRIEDRICHS, HYBRID_HLLC, AM_HLLC = 0, 1, 2, 3, 4, 5
RK2_SSP, MUSCL = 0, 1
OPEN, REFLECTIVE, PERIODIC, MHD_JET = 0, 1, 2, 3
GHOST_CELLS, PERIODIC_ROLL = 0, 1
IMPLICIT_MIDPOINT, IMPLICIT_EULER = 0, 1
SINGLE, DOUBLE = 0, 1

RS_NAMES = {HLL: "hll", HLLC: "hllc", HLLC_LM: "hllclm", LAX_FRIEDRICHS: "lf", HYBRID_HLLC: "hybrid", AM_HLLC: "amhllc"}
LIM_NAMES = {MINMOD: "minmod", OSHER: "osher", DOUBLE_MINMOD: "dminmod", SUPERBEE: "superbee",
             VAN_ALBADA: "valbada", VAN_ALBADA_PP: "valbadapp"}
GRID = {1: 64, 2: 24, 3: 12}


def case(name, dim, mhd=False, n=None, init=("smooth_random", 0), box=1.0, bnd=PERIODIC, steps=5,
         dtype="f64", params=None, t_end=None, **cfg):
    if not isinstance(bnd, list):
        bnd = [[bnd, bnd]] * dim
    return dict(name=name, dim=dim, mhd=mhd, n=n or GRID[dim], init=list(init), box=box, boundaries=bnd,
                steps=steps, dtype=dtype, params=params or {}, t_end=t_end, cfg=cfg)


def all_cases():
    cs = []
    # BASELINE configurations, small
    for rs in (HLL, HLLC):
        for dt in ("f64", "f32"):
            cs.append(case(f"sod101_{RS_NAMES[rs]}_{dt}", 1, n=101, init=("sod",), bnd=OPEN, t_end=0.2,
                           dtype=dt, riemann_solver=rs))
    for dt in ("f64", "f32"):
        cs.append(case(f"kh32_hllc_{dt}", 2, n=32, init=("kelvin_helmholtz",), steps=10, dtype=dt, riemann_solver=HLLC))
        cs.append(case(f"ot32_mhd_{dt}", 2, mhd=True, n=32, init=("orszag_tang",), box=2 * np.pi, steps=10, dtype=dt))
        cs.append(case(f"blast16_mhd_{dt}", 3, mhd=True, n=16, init=("mhd_blast3d_arena",), steps=6, dtype=dt))
    cs.append(case("blast16_mhd_strong_f64", 3, mhd=True, n=16, init=("mhd_blast3d",), steps=5))
    # Riemann solvers x limiters (hydro, unsplit RK2, periodic)
    for dim in (1, 2, 3):
        for rs in RS_NAMES:
            cs.append(case(f"hydro{dim}d_{RS_NAMES[rs]}_minmod", dim, init=("smooth_random", 10 + rs), riemann_solver=rs))
        for lim in (OSHER, DOUBLE_MINMOD, SUPERBEE, VAN_ALBADA):
            cs.append(case(f"hydro{dim}d_hll_{LIM_NAMES[lim]}", dim, init=("smooth_random", 20 + lim), limiter=lim))
            cs.append(case(f"hydro{dim}d_hllc_{LIM_NAMES[lim]}", dim, init=("smooth_random", 30 + lim),
                           riemann_solver=HLLC, limiter=lim))
    # the positivity-preserving pair of the published fv_mhd benchmark (2-D / 3-D only, Q18)
    for dim in (2, 3):
        cs.append(case(f"hydro{dim}d_lf_valbadapp", dim, init=("smooth_random", 41), riemann_solver=LAX_FRIEDRICHS,
                       limiter=VAN_ALBADA_PP))
        cs.append(case(f"hydro{dim}d_hll_valbadapp", dim, init=("smooth_random", 42), limiter=VAN_ALBADA_PP))
        cs.append(case(f"mhd{dim}d_lf_valbadapp", dim, mhd=True, init=("smooth_random", 43),
                       riemann_solver=LAX_FRIEDRICHS, limiter=VAN_ALBADA_PP))
    # MHD: solvers, magnetic integrators, precision flag of the fixed point (Q6)
    for dim in (2, 3):
        for rs in (HLL, HLLC, HLLC_LM, LAX_FRIEDRICHS, HYBRID_HLLC):
            cs.append(case(f"mhd{dim}d_{RS_NAMES[rs]}", dim, mhd=True, init=("smooth_random", 50 + rs), riemann_solver=rs,
                           numerical_precision=DOUBLE))
        cs.append(case(f"mhd{dim}d_implicit_euler", dim, mhd=True, init=("smooth_random", 57),
                       fv_magnetic_integrator=IMPLICIT_EULER, numerical_precision=DOUBLE))
        cs.append(case(f"mhd{dim}d_single_precision_tol", dim, mhd=True, init=("smooth_random", 58)))
        cs.append(case(f"mhd{dim}d_f32", dim, mhd=True, init=("smooth_random", 59), dtype="f32"))
        cs.append(case(f"hydro{dim}d_hllc_f32", dim, init=("smooth_random", 60), dtype="f32", riemann_solver=HLLC))
    # boundaries
    for dim in (1, 2, 3):
        for b, bn in ((OPEN, "open"), (REFLECTIVE, "reflective")):
            cs.append(case(f"hydro{dim}d_{bn}", dim, init=("smooth_random", 70 + b), bnd=b, riemann_solver=HLLC))
            if dim > 1:
                cs.append(case(f"mhd{dim}d_{bn}", dim, mhd=True, init=("smooth_random", 73 + b), bnd=b,
                               numerical_precision=DOUBLE))
    # reflective sides in unsplit MHD: the ghost gas cells are not images of interior cells (Q3 + Q11 + Q12)
    # and take part in the CFL reduction -- longer runs, both precisions, the solvers that need wave bounds
    cs.append(case("mhd3d_reflective_long", 3, mhd=True, init=("smooth_random", 110), bnd=REFLECTIVE, steps=12,
                   numerical_precision=DOUBLE))
    cs.append(case("mhd2d_reflective_long", 2, mhd=True, init=("smooth_random", 111), bnd=REFLECTIVE, steps=12,
                   numerical_precision=DOUBLE))
    cs.append(case("mhd3d_reflective_f32", 3, mhd=True, init=("smooth_random", 112), bnd=REFLECTIVE, dtype="f32"))
    cs.append(case("mhd2d_reflective_f32", 2, mhd=True, init=("smooth_random", 113), bnd=REFLECTIVE, dtype="f32"))
    cs.append(case("mhd3d_reflective_lf", 3, mhd=True, init=("smooth_random", 114), bnd=REFLECTIVE,
                   riemann_solver=LAX_FRIEDRICHS, numerical_precision=DOUBLE))
    cs.append(case("mhd2d_reflective_lf_valbadapp", 2, mhd=True, init=("smooth_random", 115), bnd=REFLECTIVE,
                   riemann_solver=LAX_FRIEDRICHS, limiter=VAN_ALBADA_PP, numerical_precision=DOUBLE))
    cs.append(case("mhd3d_reflective_hllc_euler", 3, mhd=True, init=("smooth_random", 116), bnd=REFLECTIVE,
                   riemann_solver=HLLC, fv_magnetic_integrator=IMPLICIT_EULER, numerical_precision=DOUBLE))
    cs.append(case("mhd2d_mixed_bnd", 2, mhd=True, init=("smooth_random", 117),
                   bnd=[[REFLECTIVE, OPEN], [OPEN, REFLECTIVE]], steps=8, numerical_precision=DOUBLE))
    cs.append(case("mhd3d_reflective_split_muscl", 3, mhd=True, init=("smooth_random", 118), bnd=REFLECTIVE,
                   split=SPLIT, time_integrator=MUSCL, params={"dt_max": 1.0}, numerical_precision=DOUBLE))
    # MHD jet inflow (tests/mhd_tests/mhd_jet.py): left y side = jet, i.e. constants in a window of the x-left
    # ghost rows; the CFL reduction sees them (|v_x| + c = 801) from the first step on
    jet_bnd = [[OPEN, OPEN], [MHD_JET, OPEN]]
    cs.append(case("mhd2d_jet", 2, mhd=True, n=80, init=("jet",), bnd=jet_bnd, steps=12, numerical_precision=DOUBLE))
    cs.append(case("mhd2d_jet_f32", 2, mhd=True, n=80, init=("jet",), bnd=jet_bnd, steps=12, dtype="f32"))
    cs.append(case("mhd2d_jet_hllc_random", 2, mhd=True, n=48, init=("smooth_random", 120), bnd=jet_bnd, steps=8,
                   riemann_solver=HLLC, numerical_precision=DOUBLE))
    cs.append(case("mhd2d_jet_split_muscl", 2, mhd=True, n=48, init=("jet",), bnd=jet_bnd, steps=8, split=SPLIT,
                   time_integrator=MUSCL, numerical_precision=DOUBLE))
    cs.append(case("hydro2d_jet", 2, n=48, init=("smooth_random", 121), bnd=jet_bnd, steps=8, riemann_solver=HLLC))
    cs.append(case("mhd2d_jet_reflective_x", 2, mhd=True, n=48, init=("jet",), bnd=[[REFLECTIVE, OPEN], [MHD_JET, REFLECTIVE]],
                   steps=8, numerical_precision=DOUBLE))
    cs.append(case("hydro2d_mixed_bnd", 2, init=("smooth_random", 76), bnd=[[PERIODIC, PERIODIC], [REFLECTIVE, OPEN]]))
    cs.append(case("mhd3d_mixed_bnd", 3, mhd=True, init=("smooth_random", 77),
                   bnd=[[OPEN, REFLECTIVE], [PERIODIC, PERIODIC], [REFLECTIVE, OPEN]], numerical_precision=DOUBLE))
    # PERIODIC_ROLL (no ghost cells)
    for dim, mhd in ((1, False), (2, False), (2, True), (3, True)):
        cs.append(case(f"{'mhd' if mhd else 'hydro'}{dim}d_periodic_roll", dim, mhd=mhd, init=("smooth_random", 80 + dim),
                       boundary_handling=PERIODIC_ROLL, numerical_precision=DOUBLE))
    # Strang-split MUSCL-Hancock sweeps (Q1), incl. the first-order fallback (Q7) and non-periodic sides
    for dim, mhd in ((1, False), (2, False), (2, True), (3, False), (3, True)):
        cs.append(case(f"{'mhd' if mhd else 'hydro'}{dim}d_split_muscl", dim, mhd=mhd, init=("smooth_random", 90 + dim),
                       split=SPLIT, time_integrator=MUSCL, params={"dt_max": 1.0}, numerical_precision=DOUBLE))
    cs.append(case("hydro2d_split_first_order", 2, init=("smooth_random", 95), split=SPLIT, time_integrator=MUSCL,
                   first_order_fallback=True))
    cs.append(case("hydro2d_split_hllc_open", 2, init=("smooth_random", 96), split=SPLIT, time_integrator=MUSCL,
                   riemann_solver=HLLC, bnd=OPEN, params={"dt_max": 1.0}))
    cs.append(case("hydro3d_split_dtmax_clipped", 3, init=("smooth_random", 97), split=SPLIT, time_integrator=MUSCL))
    # HYBRID_HLLC / AM_HLLC inside the split sweeps (the compression switch of the sweep's input state)
    for dim, mhd, rs in ((1, False, HYBRID_HLLC), (2, False, HYBRID_HLLC), (2, False, AM_HLLC), (3, False, AM_HLLC),
                         (3, True, HYBRID_HLLC), (2, True, AM_HLLC)):
        cs.append(case(f"{'mhd' if mhd else 'hydro'}{dim}d_split_{RS_NAMES[rs]}", dim, mhd=mhd,
                       init=("smooth_random", 130 + dim + rs), split=SPLIT, time_integrator=MUSCL, riemann_solver=rs,
                       params={"dt_max": 1.0}, numerical_precision=DOUBLE))
    cs.append(case("hydro2d_split_hybrid_reflective", 2, init=("smooth_random", 139), split=SPLIT, time_integrator=MUSCL,
                   riemann_solver=HYBRID_HLLC, bnd=REFLECTIVE, params={"dt_max": 1.0}))
    # HYBRID_HLLC / AM_HLLC fluxes on the positivity-preserving reconstruction
    for dim, mhd, rs in ((2, False, HYBRID_HLLC), (2, True, AM_HLLC), (3, False, AM_HLLC), (3, True, HYBRID_HLLC)):
        cs.append(case(f"{'mhd' if mhd else 'hydro'}{dim}d_{RS_NAMES[rs]}_valbadapp", dim, mhd=mhd,
                       init=("smooth_random", 140 + dim + rs), riemann_solver=rs, limiter=VAN_ALBADA_PP,
                       numerical_precision=DOUBLE))
    # parameters off their defaults
    cs.append(case("hydro2d_gamma14_cfl08", 2, init=("smooth_random", 98), riemann_solver=HLLC,
                   params={"gamma": 1.4, "C_cfl": 0.8}))
    cs.append(case("mhd2d_gamma14", 2, mhd=True, init=("smooth_random", 99), params={"gamma": 1.4},
                   numerical_precision=DOUBLE))
    cs.append(case("hydro2d_no_exact_end_time", 2, init=("smooth_random", 100), exact_end_time=False))
    return cs


def initial_state(c):
    from astronomix_b200 import problems as pr  # NumPy only
    dtype = np.float64 if c["dtype"] == "f64" else np.float32
    kind, *args = c["init"]
    n, dim = c["n"], c["dim"]
    if kind == "smooth_random":
        return pr.smooth_random(dim, n, mhd=c["mhd"], seed=args[0], dtype=dtype)
    if kind == "sod":
        return pr.sod(n, dtype)
    if kind == "kelvin_helmholtz":
        return pr.kelvin_helmholtz(n, dtype)
    if kind == "orszag_tang":
        return pr.orszag_tang(n, dtype)
    if kind == "mhd_blast3d":
        return pr.mhd_blast3d(n, dtype)
    if kind == "jet":  # tests/mhd_tests/mhd_jet.py:68-84
        s = np.zeros((7 if c["mhd"] else 4, n, n), dtype=dtype)
        s[0], s[3] = dtype(5 / 3) * dtype(0.1), 1.0
        if c["mhd"]:
            s[4] = 200 ** 0.5
        return s
    if kind == "mhd_blast3d_arena":
        return pr.mhd_blast3d(n, dtype, p_in=100.0, p_out=1.0, B0=10.0, r_inj=0.125)
    raise ValueError(kind)


def _pick_t_end(c, s0):
    """Any t_end is a valid input; this one is chosen so that the run takes about ``steps``
    steps and ends on a clipped step (time_integration.py:677-678)."""
    if c["t_end"] is not None:
        return float(c["t_end"])
    import astronomix_b200 as ax
    from astronomix_b200 import problems as pr
    from oracle import fv_oracle as fo
    config, params, _ = _mirror_config(c, 1.0)
    o = fo.Opts.from_config(config, params)
    _, info = fo.time_integration(s0, o, max_steps=c["steps"], return_info=True)
    return float(np.float32(0.97 * info["time"]))  # exactly representable in both dtypes


def _mirror_config(c, t_end):
    import astronomix_b200 as ax
    from astronomix_b200 import option_classes as oc
    from astronomix_b200 import problems as pr
    b1 = [oc.BoundarySettings1D(l, r) for l, r in c["boundaries"]]
    bs = b1[0] if c["dim"] == 1 else oc.BoundarySettings(*(b1 + [b1[-1]] * (3 - len(b1))))
    cfg = dict(c["cfg"])
    cfg.setdefault("numerical_precision", oc.SINGLE_PRECISION)  # the reference's default; pr.make's follows the dtype
    params = ax.SimulationParams(t_end=t_end, **c["params"])
    return pr.make(c["dim"], c["n"], mhd=c["mhd"], box=c["box"], boundaries=bs,
                   dtype=np.float64 if c["dtype"] == "f64" else np.float32, params=params, **cfg)


def run_reference(c, s0, t_end):
    """The reference's public entry point on the reference's own option classes."""
    x64 = c["dtype"] == "f64"
    load_reference(x64=x64)
    import jax.numpy as jnp
    sc = ref("option_classes.simulation_config")
    sp = ref("option_classes.simulation_params")
    rv = ref("variable_registry.registered_variables")
    ti = ref("time_stepping.time_integration")
    for name, val in (("MINMOD", MINMOD), ("OSHER", OSHER), ("DOUBLE_MINMOD", DOUBLE_MINMOD), ("SUPERBEE", SUPERBEE),
                      ("VAN_ALBADA", VAN_ALBADA), ("VAN_ALBADA_PP", VAN_ALBADA_PP), ("UNSPLIT", UNSPLIT), ("SPLIT", SPLIT),
                      ("HLL", HLL), ("HLLC", HLLC), ("HLLC_LM", HLLC_LM), ("LAX_FRIEDRICHS", LAX_FRIEDRICHS),
                      ("HYBRID_HLLC", HYBRID_HLLC), ("AM_HLLC", AM_HLLC), ("RK2_SSP", RK2_SSP), ("MUSCL", MUSCL),
                      ("OPEN_BOUNDARY", OPEN), ("REFLECTIVE_BOUNDARY", REFLECTIVE), ("PERIODIC_BOUNDARY", PERIODIC),
                      ("MHD_JET_BOUNDARY", MHD_JET),
                      ("GHOST_CELLS", GHOST_CELLS), ("PERIODIC_ROLL", PERIODIC_ROLL),
                      ("IMPLICIT_MIDPOINT", IMPLICIT_MIDPOINT), ("IMPLICIT_EULER", IMPLICIT_EULER),
                      ("SINGLE_PRECISION", SINGLE), ("DOUBLE_PRECISION", DOUBLE)):
        assert getattr(sc, name) == val, name
    b1 = [sc.BoundarySettings1D(left_boundary=l, right_boundary=r) for l, r in c["boundaries"]]
    bs = b1[0] if c["dim"] == 1 else sc.BoundarySettings(*(b1 + [b1[-1]] * (3 - len(b1))))
    config = sc.SimulationConfig(progress_bar=False, dimensionality=c["dim"], mhd=c["mhd"], box_size=c["box"],
                                 num_cells=c["n"], boundary_settings=bs, **c["cfg"])
    regs = rv.get_registered_variables(config)
    config = sc.finalize_config(config, s0.shape)
    wd = np.float64 if x64 else np.float32
    pk = dict(c["params"], t_end=t_end)
    base = sp.SimulationParams()
    for f in ("C_cfl", "gamma", "dt_max", "t_end", "gravitational_constant", "minimum_density", "minimum_pressure"):
        pk[f] = wd(pk.get(f, getattr(base, f)))  # jit canonicalises params leaves to the working dtype (Q20)
    params = sp.SimulationParams(**pk)
    out = ti.time_integration(jnp.array(s0), config, params, regs)
    out = np.asarray(out).view(np.ndarray)
    assert out.dtype == s0.dtype, f"{c['name']}: dtype drifted to {out.dtype}"
    assert out.shape == s0.shape
    return np.ascontiguousarray(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    args = ap.parse_args()
    os.makedirs(OUT, exist_ok=True)
    index = []
    for c in all_cases():
        if args.only and args.only not in c["name"]:
            continue
        s0 = initial_state(c)
        t_end = _pick_t_end(c, s0)
        t0 = time.time()
        out = run_reference(c, s0, t_end)
        spec = dict(c, t_end=t_end)
        spec.pop("steps")
        assert np.isfinite(out).all(), c["name"]
        np.savez_compressed(os.path.join(OUT, c["name"] + ".npz"), initial_state=s0, final_state=out,
                            spec=np.array(json.dumps(spec)))
        index.append(c["name"])
        print(f"{c['name']:40s} t_end={t_end:.6g} shape={out.shape} {time.time() - t0:.1f}s", flush=True)
    path = os.path.join(OUT, "INDEX.json")
    if args.only and os.path.exists(path):  # partial run: keep the recorded cases, add the new ones
        with open(path) as f:
            known = json.load(f)["cases"]
        index = known + [n for n in index if n not in known]
    with open(path, "w") as f:
        json.dump({"generator": "oracle/gen_reference_vectors.py", "reference": "leo1200/astronomix v0.5.1 sources "
                   "executed over oracle/jaxshim (NumPy backend)", "cases": index}, f, indent=1)


if __name__ == "__main__":
    main()
