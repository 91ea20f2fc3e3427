"""Kernel-in-isolation timing through the stage-level C-ABI (amx_hydro_stage, amx_magnetic_iter,
amx_magnetic_pair_guarded) on the 256^3 blast state: CUDA events around N launches of ONE kernel on
fixed inputs, so that kernel variants (AMX_LIB_VARIANT=...) can be compared without running the step --
also variants whose results are deliberately wrong (timing experiments).  Prints one JSON line.

    python tools/kernel_bench.py [--n 256] [--reps 20] [--what stage,stage2,iter,pair]
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from astronomix_b200 import _lib, problems as pr  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=256)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--what", default="stage,stage2,iter,pair")
    a = ap.parse_args()
    _lib.require_device()
    lib = _lib.load()
    n = a.n
    dev = torch.device("cuda:0")
    config, params, rv = pr.make(3, n, mhd=True)
    desc = _lib.make_desc(config, params, (n, n, n), _lib.AMX_F32)
    nbytes = C.c_size_t()
    _lib.check(lib.amx_workspace_bytes(C.byref(desc), C.byref(nbytes)))
    ws = torch.zeros(min(nbytes.value, 1 << 20) + 256, dtype=torch.uint8, device=dev)
    ws_ptr = C.c_void_p((ws.data_ptr() + 255) & ~255)
    s = torch.from_numpy(pr.mhd_blast3d(n, np.float32)).to(dev)
    gas, B = s[:5].contiguous(), s[5:].contiguous()
    gas1, gas2 = torch.empty_like(gas), torch.empty_like(gas)
    X0 = torch.cat([B, gas[1:4]]).contiguous()
    X1, X2 = torch.empty_like(X0), torch.empty_like(X0)
    dt = torch.tensor(1.0e-4, dtype=torch.float32, device=dev)
    alpha = torch.zeros(3, dtype=torch.float32, device=dev)
    change = torch.zeros(4, dtype=torch.float32, device=dev)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None  # noqa: E731
    cells = n ** 3

    def stage():
        _lib.check(lib.amx_hydro_stage(C.byref(desc), p(gas), None, p(gas1), p(dt), 0.5, p(alpha), None, ws_ptr, st))

    def stage2():
        _lib.check(lib.amx_hydro_stage(C.byref(desc), p(gas1), p(gas), p(gas2), p(dt), 0.5, p(alpha), None, ws_ptr, st))

    def it():
        _lib.check(lib.amx_magnetic_iter(C.byref(desc), p(B), p(gas), p(X0), p(X1), p(dt), p(change), ws_ptr, st))

    def pair():
        _lib.check(lib.amx_magnetic_pair_guarded(C.byref(desc), p(B), p(gas), p(X0), p(X2), p(dt), p(change), 2,
                                                 None, None, ws_ptr, st))

    # algorithmic bytes per cell per launch
    table = {"stage": (stage, 40), "stage2": (stage2, 60), "iter": (it, 76), "pair": (pair, 76)}
    out = {"n": n, "variant": os.environ.get("AMX_LIB_VARIANT", "")}
    stage()  # gas1 must hold a valid state for stage2
    # the pair entry point is guarded by the device-side loop state: switch the loop on (norms above the tolerance)
    change.fill_(1.0)
    _lib.check(lib.amx_magnetic_check(C.byref(desc), p(change), 0, ws_ptr, st))
    for name in a.what.split(","):
        fn, bpc = table[name]
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.reps  # includes the small finalize / control launches of the entry point
        out[name] = {"ms": round(ms, 4), "GBs": round(cells * bpc / ms / 1e6, 1)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
