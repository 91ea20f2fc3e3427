"""CUDA path against the oracle at sizes and step counts beyond the unit tests: the 1000:0.1
3-D MHD blast at 128^3 (shocks from the first step), Kelvin-Helmholtz 256^2 (HLLC) and Orszag-Tang 256^2, after
N = 1, 10, 50 steps; relative L1 per field group, FAST and STRICT arithmetic.

  python tools/error_growth.py oracle     # CPU, minutes: writes oracle/_cache/eg_*.npy (git-ignored, travels to the GPU box)
  python tools/error_growth.py gpu        # on the GPU box: prints one JSON line per case and step count
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from astronomix_b200 import option_classes as oc, problems as pr  # noqa: E402

CACHE = os.path.join(ROOT, "oracle", "_cache")
STEPS = (1, 10, 50)


def cases():
    c = {}
    config, params, rv = pr.make(3, 128, mhd=True)
    c["blast128"] = (pr.mhd_blast3d(128, np.float32), config, params._replace(t_end=1e9), rv, 3, True)
    config, params, rv = pr.make(2, 256, riemann_solver=oc.HLLC)
    c["kh256"] = (pr.kelvin_helmholtz(256, np.float32), config, params._replace(t_end=1e9), rv, 2, False)
    config, params, rv = pr.make(2, 256, mhd=True, box=2 * np.pi)
    c["ot256"] = (pr.orszag_tang(256, np.float32), config, params._replace(t_end=1e9), rv, 2, True)
    return c


def groups(dim, mhd):
    g = {"density": slice(0, 1), "velocity": slice(1, 1 + dim), "pressure": slice(1 + dim, 2 + dim)}
    if mhd:
        g["magnetic"] = slice(2 + dim, 5 + dim)
    return g


def main():
    mode = sys.argv[1]
    os.makedirs(CACHE, exist_ok=True)
    if mode == "oracle":
        from oracle import fv_oracle as fo
        for name, (s, config, params, rv, dim, mhd) in cases().items():
            o = fo.Opts.from_config(config, params)
            cur, done = s, 0
            for n in STEPS:  # continue from the previous checkpoint: the loop has no state besides the array
                cur = fo.time_integration(cur, o, max_steps=n - done)
                done = n
                np.save(os.path.join(CACHE, f"eg_{name}_{n}.npy"), cur)
                print(name, n, "done", flush=True)
        return
    from astronomix_b200 import _lib, stepping as ti
    for name, (s, config, params, rv, dim, mhd) in cases().items():
        for arith, code in (("fast", _lib.AMX_ARITH_FAST), ("strict", _lib.AMX_ARITH_STRICT)):
            for n in STEPS:
                ref = np.load(os.path.join(CACHE, f"eg_{name}_{n}.npy"))
                out = ti.integrate_steps(s, config, params, rv, n, arithmetic=code)
                err = {}
                for g, sl in groups(dim, mhd).items():
                    den = np.abs(ref[sl]).sum()
                    err[g] = float(np.abs(out[sl].astype(np.float64) - ref[sl]).sum() / max(den, 1e-300))
                print(json.dumps({"case": name, "arithmetic": arith, "steps": n, "rel_l1": err,
                                  "worst": max(err.values())}), flush=True)


if __name__ == "__main__":
    main()
