"""The oracle against the reference's only numerical fixture for this path.

The full 3001-step run (7 minutes of NumPy) is done by oracle/pin_orszag_tang.py, whose
record is committed as tests/golden/pin_orszag_tang.json.  Here the last 101 steps are
re-run from the committed step-2900 checkpoint on every CPU test run."""
import json
import os

import numpy as np

from oracle import fv_oracle as fo
from oracle.pin_orszag_tang import CKPT, errors, fixture_opts, orszag_tang_state

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_full_run_record_is_within_1e11():
    rec = json.load(open(os.path.join(GOLD, "pin_orszag_tang.json")))
    assert rec["steps"] == 3001
    assert rec["worst_rel_l1"] < 1e-11
    assert rec["errors"]["Bz"]["max_abs"] == 0.0


def test_fixture_from_checkpoint():
    ref = np.load(os.path.join(GOLD, "orszag_tang_fixture.npz"))["final_state"]
    ck = np.load(os.path.join(GOLD, "orszag_tang_ckpt.npz"))
    assert int(ck["step"]) == CKPT
    _, dx = orszag_tang_state()
    o = fixture_opts(dx)
    out, steps = fo.resume_padded(ck["padded_state"], float(ck["time"]), o)
    assert steps == 3001 - CKPT
    err = errors(out, ref)
    assert max(v["rel_l1"] for v in err.values()) < 1e-11, err


def test_fixture_invariants():
    """Properties of the reference's file itself: mass, energy, div B, B_z."""
    ref = np.load(os.path.join(GOLD, "orszag_tang_fixture.npz"))["final_state"]
    s0, dx = orszag_tang_state()
    m0 = s0[0].sum()
    assert abs(ref[0].sum() - m0) / m0 < 1e-13
    assert np.abs(fo.divergence(ref[4:7], dx, 2)).max() < 1e-11
    assert np.all(ref[6] == 0.0)
