"""Pin the oracle on the reference's only numerical fixture for this path.

Runs ``oracle/fv_oracle.py`` in the configuration that produced
``/root/reference/tests/mhd_tests/final_state_orszag_tang.npz`` (2-D
Orszag-Tang 200^2, box 2*pi, fp64, SPLIT + MUSCL, HLL, minmod, periodic ghost cells,
C_cfl 0.4, dt_max 1e-3, implicit midpoint with tolerance 1e-10, t_end 3 -> 3001 steps;
initial condition from tests/mhd_tests/orszag_tang.py:39-100) and writes

  tests/golden/orszag_tang_fixture.npz      the reference's file, re-saved compressed (data, not source)
  tests/golden/orszag_tang_ckpt.npz         the oracle's padded state + time at step CKPT
  tests/golden/pin_orszag_tang.json         per-field relative L1 / max-abs errors of the full run

Takes ~10 minutes of NumPy; run by hand in the build container (needs /root/reference).
tests/test_oracle_pin.py re-runs the last 3001-CKPT steps from the checkpoint on every
CPU test run and checks the fixture again.
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import fv_oracle as fo  # noqa: E402

REF = "/root/reference/tests/mhd_tests/final_state_orszag_tang.npz"
GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")
CKPT = 2900


def orszag_tang_state(n=200, dtype=np.float64):
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
    return s, float(dx)


def fixture_opts(dx):
    return fo.Opts(dim=2, mhd=True, riemann_solver=fo.HLL, limiter=fo.MINMOD, split=fo.SPLIT,
                   time_integrator=fo.MUSCL, boundary_handling=fo.GHOST_CELLS,
                   boundaries=[(fo.PERIODIC_BOUNDARY,) * 2] * 2, grid_spacing=dx, gamma=5 / 3,
                   C_cfl=0.4, dt_max=1e-3, t_end=3.0, numerical_precision=fo.DOUBLE_PRECISION)


def errors(a, ref):
    out = {}
    names = ["rho", "vx", "vy", "p", "Bx", "By", "Bz"]
    for k, nme in enumerate(names):
        denom = np.sum(np.abs(ref[k]))
        l1 = float(np.sum(np.abs(a[k] - ref[k])) / denom) if denom > 0 else float(np.sum(np.abs(a[k])))
        out[nme] = {"rel_l1": l1, "max_abs": float(np.max(np.abs(a[k] - ref[k])))}
    return out


if __name__ == "__main__":
    ref = np.load(REF)["final_state"]
    s0, dx = orszag_tang_state()
    o = fixture_opts(dx)
    t0 = time.time()
    out, info = fo.time_integration(s0, o, return_info=True, checkpoint_at=CKPT)
    wall = time.time() - t0
    err = errors(out, ref)
    rec = {"steps": info["steps"], "final_time": info["time"], "wall_s": wall,
           "mean_fixed_point_iterations": float(np.mean(info["iterations"])),
           "checkpoint_step": CKPT, "errors": err,
           "worst_rel_l1": max(v["rel_l1"] for v in err.values())}
    print(json.dumps(rec, indent=1))
    os.makedirs(GOLD, exist_ok=True)
    np.savez_compressed(os.path.join(GOLD, "orszag_tang_fixture.npz"), final_state=ref)
    ck, tck = info["checkpoint"]
    np.savez_compressed(os.path.join(GOLD, "orszag_tang_ckpt.npz"), padded_state=ck, time=np.float64(tck),
                        step=np.int64(CKPT))
    with open(os.path.join(GOLD, "pin_orszag_tang.json"), "w") as f:
        json.dump(rec, f, indent=1)
