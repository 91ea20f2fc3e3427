"""Oracle properties that do not depend on trusting the oracle (SURVEY.md §8c i-vii):
they guard the parts no reference artefact pins (unsplit RK2 driver, HLLC family, 3-D)."""
import numpy as np
import pytest

from oracle import fv_oracle as fo
from astronomix_b200 import problems as pr


def opts(dim, dx, **kw):
    kw.setdefault("boundaries", [(fo.PERIODIC_BOUNDARY,) * 2] * dim)
    kw.setdefault("numerical_precision", fo.DOUBLE_PRECISION)
    return fo.Opts(dim=dim, grid_spacing=dx, **kw)


@pytest.mark.parametrize("dim,n", [(1, 64), (2, 24), (3, 12)])
@pytest.mark.parametrize("rs", [fo.HLL, fo.HLLC, fo.HLLC_LM, fo.LAX_FRIEDRICHS, fo.HYBRID_HLLC])
def test_periodic_hydro_conserves_mass_momentum_energy(dim, n, rs):
    s = pr.smooth_random(dim, n, seed=dim, dtype=np.float64)
    o = opts(dim, 1.0 / n, riemann_solver=rs, t_end=1.0)
    out, info = fo.time_integration(s, o, max_steps=5, return_info=True)
    m0, p0, e0 = pr.totals(s, dim, False)
    m1, p1, e1 = pr.totals(out, dim, False)
    assert info["steps"] == 5
    assert abs(m1 - m0) / m0 < 1e-13
    assert abs(e1 - e0) / e0 < 1e-13
    for a, b in zip(p0, p1):
        assert abs(a - b) / m0 < 1e-13


@pytest.mark.parametrize("dim,n", [(2, 24), (3, 12)])
def test_magnetic_update_keeps_divB_and_pressure(dim, n):
    s = pr.smooth_random(dim, n, mhd=True, seed=7, dtype=np.float64)
    dx = 1.0 / n
    o = opts(dim, dx, mhd=True)
    sp = fo.prepare_state(s, o)
    gas, B = sp[:-3], sp[-3:]
    B = fo.boundary_handler(B, o)
    div0 = fo.divergence(B, dx, dim)
    B1, gas1, it = fo.magnetic_update(B, gas, np.float64(2e-3), o, return_iterations=True)
    assert 1 <= it < 50
    div1 = fo.divergence(B1, dx, dim)
    inner = (slice(2, -2),) * dim
    assert np.abs(div1 - div0)[inner].max() < 1e-11 * max(1.0, np.abs(div0).max())
    # Q4: pressure changes by round-off only
    assert np.abs(gas1[dim + 1] - gas[dim + 1]).max() < 1e-13


def test_mhd_energy_conserved_by_implicit_midpoint():
    n = 32
    s = pr.orszag_tang(n)
    o = opts(2, 2 * np.pi / n, mhd=True, t_end=1.0)
    out = fo.time_integration(s, o, max_steps=10)
    _, _, e0 = pr.totals(s, 2, True)
    _, _, e1 = pr.totals(out, 2, True)
    assert abs(e1 - e0) / e0 < 1e-11


def test_ghost_cells_equal_periodic_roll_bitwise():
    for dim, n, mhd in [(1, 40, False), (2, 20, True), (3, 10, True)]:
        s = pr.smooth_random(dim, n, mhd=mhd, seed=3, dtype=np.float32)
        kw = dict(mhd=mhd, riemann_solver=fo.HLLC, numerical_precision=fo.SINGLE_PRECISION, t_end=1.0)
        a = fo.time_integration(s, opts(dim, 1.0 / n, boundary_handling=fo.GHOST_CELLS, **kw), max_steps=3)
        b = fo.time_integration(s, opts(dim, 1.0 / n, boundary_handling=fo.PERIODIC_ROLL, **kw), max_steps=3)
        assert a.dtype == np.float32
        assert np.array_equal(a, b)


def test_axis_permutation_symmetry_3d():
    """Transposing x<->z of the input (and swapping vx<->vz) transposes the output."""
    n = 10
    s = pr.smooth_random(3, n, seed=11, dtype=np.float64)
    o = opts(3, 1.0 / n, riemann_solver=fo.HLLC, t_end=1.0)
    a = fo.time_integration(s, o, max_steps=2)
    st = s.transpose(0, 3, 2, 1)[[0, 3, 2, 1, 4]]
    b = fo.time_integration(np.ascontiguousarray(st), o, max_steps=2)
    bt = b.transpose(0, 3, 2, 1)[[0, 3, 2, 1, 4]]
    assert np.abs(a - bt).max() < 1e-13


def test_second_order_convergence_gaussian_pulse():
    """tests/hydro_tests/convergence.py:47-86: advected density pulse, periodic, HLLC."""
    errs = []
    for n in (32, 64, 128):
        x = pr.centers(n)
        s = np.zeros((3, n))
        s[0] = 1.0 + 0.5 * np.exp(-((x - 0.5) ** 2) / 0.01)
        s[1] = 1.0
        s[2] = 1.0
        o = opts(1, 1.0 / n, riemann_solver=fo.HLLC, C_cfl=0.1, t_end=0.1, dt_max=1.0)
        out = fo.time_integration(s, o)
        xe = (x - 0.1) % 1.0
        exact = 1.0 + 0.5 * np.exp(-((xe - 0.5) ** 2) / 0.01)
        errs.append(np.abs(out[0] - exact).mean())
    order = np.log2(errs[0] / errs[1]), np.log2(errs[1] / errs[2])
    assert order[1] > 1.6, (errs, order)


def test_sod_against_exact_riemann_solution():
    """<= 2% L1 density error at 101 cells, HLL and HLLC, open boundaries."""
    n = 101
    s = pr.sod(n, np.float64)
    x = pr.centers(n)
    # exact Sod solution at t = 0.2, gamma = 5/3 (star values from the standard iteration)
    g = 5 / 3
    from scipy.optimize import brentq
    rl, pl, rr, prr = 1.0, 1.0, 0.125, 0.1
    cl, cr = np.sqrt(g * pl / rl), np.sqrt(g * prr / rr)

    def f(p):
        fl = 2 * cl / (g - 1) * ((p / pl) ** ((g - 1) / (2 * g)) - 1)
        A, Bc = 2 / ((g + 1) * rr), (g - 1) / (g + 1) * prr
        fr = (p - prr) * np.sqrt(A / (p + Bc))
        return fl + fr

    ps = brentq(f, 1e-6, 1.0)
    us = -2 * cl / (g - 1) * ((ps / pl) ** ((g - 1) / (2 * g)) - 1)
    rsl = rl * (ps / pl) ** (1 / g)
    rsr = rr * ((ps / prr + (g - 1) / (g + 1)) / ((g - 1) / (g + 1) * ps / prr + 1))
    S = cr * np.sqrt((g + 1) / (2 * g) * ps / prr + (g - 1) / (2 * g))
    t = 0.2
    xi = (x - 0.5) / t
    csl = cl * (ps / pl) ** ((g - 1) / (2 * g))
    rho = np.where(xi < -cl, rl, np.where(xi < us - csl,
          rl * (2 / (g + 1) - (g - 1) / ((g + 1) * cl) * xi) ** (2 / (g - 1)),
          np.where(xi < us, rsl, np.where(xi < S, rsr, rr))))
    for rs in (fo.HLL, fo.HLLC):
        o = fo.Opts(dim=1, riemann_solver=rs, grid_spacing=1.0 / n, t_end=0.2,
                    numerical_precision=fo.DOUBLE_PRECISION)
        out = fo.time_integration(s, o)
        assert np.abs(out[0] - rho).mean() / rho.mean() < 0.02


def test_van_albada_pp_keeps_interface_states_positive():
    """reconstruction.py:134-394: with the PP scaling every interface density / pressure stays
    >= 0 where plain van Albada undershoots, and on smooth data the scaling is the identity."""
    import numpy as np
    from astronomix_b200 import problems as pr, option_classes as oc
    from oracle import fv_oracle as fo
    rng = np.random.default_rng(5)
    for dim, n in [(2, 24), (3, 12)]:
        s = pr.smooth_random(dim, n, seed=3, dtype=np.float64)
        s[0] = 10 ** rng.uniform(-3, 1, s[0].shape)
        s[dim + 1] = 10 ** rng.uniform(-4, 1, s[0].shape)
        s[1:1 + dim] = rng.normal(0, 2, s[1:1 + dim].shape)
        c, p, _ = pr.make(dim, n, dtype=np.float64, riemann_solver=oc.LAX_FRIEDRICHS, limiter=oc.VAN_ALBADA_PP)
        o = fo.Opts.from_config(c, p)
        padded = fo._bc(fo.pad(s, o), o)
        Ls, Rs = fo.reconstruct_unsplit_pp(padded, np.float64(p.gamma), o)
        o_va = fo.Opts.from_config(*pr.make(dim, n, dtype=np.float64, limiter=oc.VAN_ALBADA)[:2])
        L_va, _ = fo.reconstruct_unsplit_single(padded, o_va, 1)
        assert L_va[0].min() < 0 and L_va[dim + 1].min() < 0          # the problem is harsh enough
        for X in Ls + Rs:
            assert X[0].min() >= 0 and X[dim + 1].min() >= 0
        out = fo.time_integration(s, o, max_steps=3)
        assert np.isfinite(out).all() and out[0].min() > 0 and out[dim + 1].min() > 0
        smooth = pr.smooth_random(dim, n, seed=4, dtype=np.float64)
        a = fo.time_integration(smooth, o, max_steps=2)
        b = fo.time_integration(smooth, fo.Opts.from_config(
            *pr.make(dim, n, dtype=np.float64, riemann_solver=oc.LAX_FRIEDRICHS, limiter=oc.VAN_ALBADA)[:2]), max_steps=2)
        assert np.array_equal(a, b)


def test_snapshot_schedule_of_the_oracle():
    """time_integration.py:397-570: one snapshot per step at most, taken at the start of the first step
    whose time has reached k * t_end / N; checkpoints the loop never reaches stay zero."""
    import numpy as np
    from astronomix_b200 import problems as pr
    from oracle import fv_oracle as fo
    c, p, _ = pr.make(2, 16, mhd=True, dtype=np.float64)
    p = p._replace(t_end=0.02)
    o = fo.Opts.from_config(c, p)
    s = pr.smooth_random(2, 16, mhd=True, seed=1, dtype=np.float64)
    out, rec, steps = fo.time_integration_snapshots(s, o, 4)
    assert np.array_equal(out, fo.time_integration(s, o))           # snapshots do not change the run
    assert steps == 3 and rec["current_checkpoint"] == 3            # 3 steps -> at most 3 snapshots
    tp = rec["time_points"]
    assert tp[0] == 0 and tp[3] == 0 and np.all(np.diff(tp[:3]) > 0)
    for k in range(1, 3):
        assert tp[k] >= k * 0.02 / 4
    assert np.allclose(rec["total_mass"][:3], rec["total_mass"][0], rtol=1e-13)  # periodic: mass conserved
    out2, rec2, steps2 = fo.time_integration_snapshots(s, o, 40)     # more snapshots than steps
    assert steps2 == 3 and rec2["current_checkpoint"] == 3 and np.count_nonzero(rec2["time_points"]) == 2
