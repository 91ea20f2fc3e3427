"""Device-resident throughput of the other BASELINE.json configurations (parity-test cases, not the
bench line): C1 Sod 101 (launch-bound), C2 Kelvin-Helmholtz 1024^2 HLLC hydro, C3 Orszag-Tang
2048^2 MHD, plus the split MUSCL-Hancock flavour of C4.  One JSON line per configuration.
  python tools/bench_configs.py > gpurun_out/configs.jsonl"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from astronomix_b200 import _lib, option_classes as oc, problems as pr
from bench import KERNEL_CLASSES

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
_lib.require_device()
lib = _lib.load()
peak = 6550.4
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                       "MEASURED_PEAKS.json")))["hbm_gbs"]
except (OSError, KeyError):
    pass


def run(name, state, config, params, steps, warm, alg_bytes):
    params = params._replace(t_end=1.0e9)
    s0 = torch.from_numpy(state).to(dev)
    desc = _lib.make_desc(config, params, state.shape[1:], _lib.AMX_F32)
    nbytes = C.c_size_t()
    _lib.check(lib.amx_workspace_bytes(C.byref(desc), C.byref(nbytes)))
    ws = torch.empty(nbytes.value + 256, dtype=torch.uint8, device=dev)
    ws_ptr = C.c_void_p((ws.data_ptr() + 255) & ~255)
    s1, s2 = torch.empty_like(s0), torch.empty_like(s0)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    stats = _lib.AmxStats()

    def integ(a, b, k):
        _lib.check(lib.amx_integrate(C.byref(desc), C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), 1.0e9, k,
                                     ws_ptr, stream, C.byref(stats)))
    integ(s0, s1, warm)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    integ(s1, s2, steps)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    lib.amx_profile_enable(1)
    integ(s1, s2, steps)
    torch.cuda.synchronize()
    kernels = {}
    for k, cname in enumerate(KERNEL_CLASSES):
        t, c = C.c_double(), C.c_int64()
        lib.amx_profile_read(k, C.byref(t), C.byref(c))
        if c.value:
            kernels[cname] = {"launches": c.value, "ms_total": round(t.value, 4)}
    lib.amx_profile_enable(0)
    cells = int(np.prod(state.shape[1:]))
    it = stats.magnetic_iterations / steps
    value = cells * steps / (ms * 1e-3)
    b = alg_bytes(it)
    print(json.dumps({"config": name, "cells": cells, "steps": steps, "ms_per_step": ms / steps,
                      "cell_updates_per_s": value, "magnetic_iterations_per_step": it,
                      "algorithmic_bytes_per_cell_update": b, "frac_of_hbm": value * b / 1e9 / peak,
                      "launches": int(stats.kernel_launches), "kernels": kernels}), flush=True)


w = 4
c, p, rv = pr.make(1, 101, boundaries="open", riemann_solver=oc.HLLC)
run("C1 Sod 101, HLLC, open (launch-bound)", pr.sod(101, np.float32), c, p, 200, 20, lambda it: 5 * 3 * w)
c, p, rv = pr.make(2, 1024, riemann_solver=oc.HLLC)
run("C2 Kelvin-Helmholtz 1024^2 hydro, HLLC, RK2", pr.kelvin_helmholtz(1024, np.float32), c, p, 50, 5,
    lambda it: 5 * 4 * w)
c, p, rv = pr.make(2, 2048, mhd=True, box=2 * np.pi)
run("C3 Orszag-Tang 2048^2 MHD (Pang-Wu), HLL, RK2", pr.orszag_tang(2048, np.float32), c, p, 20, 3,
    lambda it: (10 * 4 + 13 + 19 * it) * w)
c, p, rv = pr.make(3, 256, mhd=True, split=oc.SPLIT, time_integrator=oc.MUSCL)
run("C4 variant: 3-D MHD blast 256^3, split MUSCL-Hancock sweeps (generic kernels)",
    pr.mhd_blast3d(256, np.float32), c, p, 5, 1, lambda it: (2 * 5 * 2 * 5 + 13 + 19 * it) * w)
