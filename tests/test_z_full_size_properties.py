"""Size-independent properties at BASELINE.json's full single-GPU sizes.

The oracle cannot follow the CUDA path to 256^3 / 1024^2 / 2048^2 (4.8e5 cell-updates/s), so at those
sizes the kernels are held to properties that need no second implementation:

* finite, rho > 0, p > 0;
* mass conserved to round-off (flux-difference form; hydro: momentum and energy too);
* translation invariance on the periodic box: shifting the initial state by a few cells along every
  axis and shifting the result back must give the same state.  Every cell is then computed by a
  different thread, tile, halo path and block than before, while the arithmetic per cell and the
  global reductions (max) are the same -- an indexing, staging or halo mistake anywhere in the grid
  shows up as an O(1) difference.  STRICT arithmetic follows one operation order on every path
  (bound 1e-6, in practice identical bits); FAST lets the compiler contract each path's expressions
  its own way (bound: the north-star 1e-5);
* for Orszag-Tang: B_z stays exactly zero (the reference's fixture has the same property).

The CPU leg runs the same checks on the oracle at small sizes: it pins the test logic and is an
invariant test of the oracle in its own right.

Further down: the layout seams (_pad / _unpad / _boundary_handler on the C-ABI, identical bits against the
oracle) and the seam loop with the MHD jet inflow.  This file sorts last on purpose: its GPU legs were
written after the round's GPU minutes were spent and have not been on hardware yet, so under ``-x`` they
cannot hide a test that has."""
import numpy as np
import pytest

from astronomix_b200 import option_classes as oc
from astronomix_b200 import problems as pr
from oracle import fv_oracle as fo

# BASELINE.json configs 2-4 (config 1 runs at its real size everywhere; config 5 is the 8-GPU run)
CASES = {
    "C2_kelvin_helmholtz": dict(dim=2, mhd=False, rs=oc.HLLC, init=pr.kelvin_helmholtz, box=1.0, full=1024, small=32),
    "C3_orszag_tang": dict(dim=2, mhd=True, rs=oc.HLL, init=pr.orszag_tang, box=2 * np.pi, full=2048, small=32),
    "C4_mhd_blast": dict(dim=3, mhd=True, rs=oc.HLL, init=pr.mhd_blast3d, box=1.0, full=256, small=16),
}
SHIFT = (3, 5, 7)


def check_properties(run, case, n, dtype, steps, tol_shift, tol_cons):
    """run(state, config, params, rv, steps) -> state after `steps` steps."""
    c = CASES[case]
    dim, mhd = c["dim"], c["mhd"]
    s = c["init"](n, dtype)
    config, params, rv = pr.make(dim, n, mhd=mhd, box=c["box"], dtype=dtype, riemann_solver=c["rs"])
    params = params._replace(t_end=1.0e9)
    out = run(s, config, params, rv, steps)
    assert out.shape == s.shape and out.dtype == s.dtype and np.isfinite(out).all()
    assert out[0].min() > 0 and out[dim + 1].min() > 0
    m0, p0, e0 = pr.totals(s, dim, mhd)
    m1, p1, e1 = pr.totals(out, dim, mhd)
    assert abs(m1 - m0) <= tol_cons * m0
    if not mhd:
        assert abs(e1 - e0) <= tol_cons * e0
        for a, b in zip(p0, p1):
            assert abs(a - b) <= tol_cons * m0
    if case == "C3_orszag_tang":
        assert not out[-1].any(), "B_z must stay exactly zero"
    axes = tuple(range(1, dim + 1))
    shift = SHIFT[:dim]
    back = tuple(-k for k in shift)
    out2 = run(np.ascontiguousarray(np.roll(s, shift, axis=axes)), config, params, rv, steps)
    err = pr.rel_l1(np.roll(out2, back, axis=axes), out, dim, mhd)
    assert err <= tol_shift, f"{case} n={n}: shifted run differs by rel-L1 {err:.3e}"
    return err


def oracle_run(s, config, params, rv, steps):
    return fo.time_integration(s, fo.Opts.from_config(config, params), max_steps=steps)


@pytest.mark.parametrize("case", sorted(CASES))
def test_oracle_properties_small(case):
    err = check_properties(oracle_run, case, CASES[case]["small"], np.float64, 3, 1e-12, 1e-12)
    assert err <= 1e-13


@pytest.mark.gpu
@pytest.mark.parametrize("arith,tol_shift", [("strict", 1e-6), ("fast", 1e-5)])
@pytest.mark.parametrize("case", sorted(CASES))
def test_cuda_properties_full_size(cuda, case, arith, tol_shift):
    from astronomix_b200 import _lib
    from astronomix_b200 import stepping as ti
    code = _lib.AMX_ARITH_STRICT if arith == "strict" else _lib.AMX_ARITH_FAST

    def run(s, config, params, rv, steps):
        out = ti.integrate_steps(s, config, params, rv, steps, arithmetic=code)
        assert ti.last_info.steps == steps
        return out

    check_properties(run, case, CASES[case]["full"], np.float32, 4, tol_shift, 2e-6)


# ---- the layout seams: _pad / _unpad / _boundary_handler on the C-ABI ----
def _boundaries(dim, kind):
    b = oc.BoundarySettings1D
    if kind == "mixed":
        sides = [b(oc.OPEN_BOUNDARY, oc.REFLECTIVE_BOUNDARY), b(oc.PERIODIC_BOUNDARY, oc.PERIODIC_BOUNDARY),
                 b(oc.REFLECTIVE_BOUNDARY, oc.OPEN_BOUNDARY)]
    elif kind == "jet":
        sides = [b(oc.REFLECTIVE_BOUNDARY, oc.OPEN_BOUNDARY), b(oc.MHD_JET_BOUNDARY, oc.OPEN_BOUNDARY),
                 b(oc.OPEN_BOUNDARY, oc.OPEN_BOUNDARY)]
    else:
        return kind
    return sides[0] if dim == 1 else oc.BoundarySettings(*sides)


@pytest.mark.gpu
@pytest.mark.parametrize("dim,n,mhd,kind", [(1, 33, False, "open"), (1, 33, False, "reflective"), (1, 32, False, "periodic"),
                                            (2, 20, False, "mixed"), (2, 20, True, "reflective"), (2, 48, True, "jet"),
                                            (2, 48, False, "jet"), (3, 10, True, "mixed"), (3, 10, True, "reflective"),
                                            (3, 10, False, "periodic"), (3, 10, True, "open")])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_layout_seams_vs_oracle(cuda, dim, n, mhd, kind, dtype):
    """Pure data movement (copies, sign flips, constants): identical bits required."""
    from astronomix_b200 import stepping as ti
    s = pr.smooth_random(dim, n, mhd=mhd, seed=3, dtype=dtype)
    config, params, rv = pr.make(dim, n, mhd=mhd, boundaries=_boundaries(dim, kind), dtype=dtype)
    o = fo.Opts.from_config(config, params)
    ng = dim + 2
    padded = ti._pad(s, config)
    assert np.array_equal(padded, fo.pad(s, o))                       # _utils.py:65-75: edge mode, every row
    assert np.array_equal(ti._unpad(padded, config), s)               # _utils.py:19-61
    gas = ti._boundary_handler(np.ascontiguousarray(padded[:ng]), config, 0)
    assert np.array_equal(gas, fo.boundary_handler(padded[:ng], o, fo.GAS_STATE))
    assert np.array_equal(fo.prepare_state(s, o)[:ng], gas)           # time_integration.py:274-286
    if mhd:
        rng = np.random.default_rng(4)
        vec = rng.standard_normal((3,) + padded.shape[1:]).astype(dtype)
        for kind_, code in ((fo.VELOCITY_ONLY, 1), (fo.MAGNETIC_FIELD_ONLY, 2)):
            assert np.array_equal(ti._boundary_handler(vec, config, code), fo.boundary_handler(vec, o, kind_))


def _jet_dt_bound(gamma):
    """Relative bound on |dt - dt_ref| in the FAST flavour.  dt = C / sum_a(alpha_a / dx) is attained in the
    inflow ghost cells (boundaries.py:399-441): rho = 5/3, v_x = 800, p = 1.  Their pressure is recovered as
    p = (gamma-1) * (E - rho v^2 / 2) with E / p ~ 5e5 (_equations.py:43-94), so each rounding of E or of the
    kinetic term (relative eps) moves p by (E/p) * eps; c = sqrt(gamma p / rho) moves by half of that
    relatively, and alpha = |v_x| + c by c / (2 alpha) of that.  The FAST flavour contracts multiply-adds and
    uses reciprocal multiplies, i.e. its roundings differ from NumPy's in every one of the ~10 operations of
    cons -> average -> prim: 16 roundings is a safe count.  STRICT follows NumPy's operation order."""
    rho, v, p = 5.0 / 3.0, 800.0, 1.0
    E = p / (gamma - 1.0) + 0.5 * rho * v * v
    c = np.sqrt(gamma * p / rho)
    return 16 * np.finfo(np.float64).eps * (E / p) * c / (2.0 * (v + c))


@pytest.mark.gpu
@pytest.mark.parametrize("arith", ["strict", "fast"])
@pytest.mark.parametrize("mhd", [True, False])
def test_seam_loop_with_jet_inflow(cuda, mhd, arith, monkeypatch):
    """_cfl_time_step -> _evolve_state_fv on padded arrays with the MHD jet boundary, each step fed with
    the previous output: dt (bounded by the inflow cells, |v_x| + c = 801) and the whole padded state
    -- ghost cells of the gas AND of the magnetic rows -- must follow the oracle."""
    from astronomix_b200 import stepping as ti
    monkeypatch.setenv("AMX_ARITHMETIC", arith)
    dim, n = 2, 48
    s = pr.smooth_random(dim, n, mhd=mhd, seed=120, dtype=np.float64)
    config, params, rv = pr.make(dim, n, mhd=mhd, boundaries=_boundaries(dim, "jet"), dtype=np.float64,
                                 numerical_precision=oc.DOUBLE_PRECISION)
    o = fo.Opts.from_config(config, params)
    g, C, dm = np.float64(params.gamma), np.float64(params.C_cfl), np.float64(params.dt_max)
    dt_tol = 1e-13 if arith == "strict" else max(1e-13, _jet_dt_bound(float(params.gamma)))  # ~1.2e-12
    sp_ref = fo.prepare_state(s, o)
    sp = sp_ref.copy()
    for _ in range(3):
        dt_ref = fo.cfl_time_step(sp_ref, o, g, C, dm)
        dt = ti._cfl_time_step(sp, config.grid_spacing, params.dt_max, params.gamma, config, rv, params.C_cfl)
        assert abs(float(dt) - float(dt_ref)) <= dt_tol * float(dt_ref), (float(dt), float(dt_ref))
        assert float(dt_ref) < 0.4 * config.grid_spacing / 800.0  # the inflow cells bound it
        sp_ref = fo.evolve_state_fv(sp_ref, dt_ref, g, o)
        sp = ti._evolve_state_fv(sp, float(dt_ref), params.gamma, 1.0, config, params, None, rv)
        assert np.abs(sp - sp_ref).max() <= 1e-12 * np.abs(sp_ref).max()  # every row, ghost cells included
        assert pr.rel_l1(fo.unpad(sp, o), fo.unpad(sp_ref, o), dim, mhd) < 1e-12


# ---- reference vectors whose CUDA leg is new (see tests/test_reference_vectors.py) ----
import test_reference_vectors as _rv  # noqa: E402


@pytest.mark.gpu
@pytest.mark.parametrize("name", _rv.NOT_YET_ON_HARDWARE)
def test_cuda_reproduces_new_reference_vectors(cuda, name):
    _rv.check_cuda_vector(name)
