#!/usr/bin/env python
"""bench.py -- cell-updates/s of the fused finite-volume MHD step on B200.

N = 1 (BASELINE.json config 4, the configuration the north star's 60 % target is quoted on): 3-D MHD blast
wave, 256^3 cells, fp32, HLL, minmod, unsplit RK2-SSP, Pang-Wu magnetic sub-step (implicit midpoint),
periodic, C_cfl 0.4, synthetic initial condition of tests/mhd_tests/mhd_blast3D.py.  One "step" is one
full time step: CFL dt + hydro(dt/2) . magnetic(dt) . hydro(dt/2).  The timed region is blocks of exactly
K steps (one amx_integrate call each, always from the same warmed-up state), repeated until it is at
least one second long.  `secondary` carries the other BASELINE configurations (C1 Sod, C2
Kelvin-Helmholtz 1024^2, C3 Orszag-Tang 2048^2), the published benchmark's flux (Lax-Friedrichs,
C_cfl 0.8, PERIODIC_ROLL; arena/fv_mhd.py:36-49) and the 512^3 single-GPU point of the weak series.

N > 1 (config 5 geometry): the north-star weak series, 512^3 cells per GPU in slabs along x
(1024x512^2 on 2, 1024^2x512 on 4, 1024^3 = 128x1024^2 slabs on 8), halo planes stored into the
neighbours' ghost planes by the producing kernels, dt / fixed-point MAX all-reduces through peer memory;
`config.secondary` keeps round 1's 256^3-per-rank line.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]

prints ONE JSON line (rank 0).  --impl reference times the CPU oracle (the NumPy restatement of the
reference; JAX is not installable here) on the host cores, on a bounded sample that its
`config.workload` names.
"""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "cell-updates/sec (3D MHD)"
UNIT = "cell-updates/s"
N_CELLS = 256  # per axis, per GPU
KERNEL_CLASSES = ["hydro_stage", "split_sweep", "magnetic_iter", "magnetic_final", "wave_bound", "control",
                  "magnetic_pair"]


def workload_config(n=N_CELLS):
    return {"workload": f"3D MHD blast {n}^3 per GPU, fp32, HLL, minmod, unsplit RK2-SSP, "
                        "Pang-Wu implicit-midpoint magnetic step, periodic, C_cfl=0.4",
            "cells_per_gpu": n ** 3, "l2_policy": "inputs larger than L2 (state 0.54 GB per GPU)"}


def make_problem(n, dtype=np.float32):
    from astronomix_b200 import problems as pr
    config, params, rv = pr.make(3, n, mhd=True, dtype=dtype)
    params = params._replace(t_end=1.0e9)  # never reached: the step count is fixed
    state = pr.mhd_blast3d(n, dtype)
    return config, params, rv, state


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.lines = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for ts, line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9 or not (t0 - 0.05 <= ts <= t1 + 0.15):
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU arms
def _oracle_worker(args):
    n, steps, warmup = args
    from astronomix_b200 import problems as pr
    from oracle import fv_oracle as fo
    config, params, rv = pr.make(3, n, mhd=True)
    params = params._replace(t_end=1.0e9)
    s = pr.mhd_blast3d(n, np.float32)
    o = fo.Opts.from_config(config, params)
    fo.time_integration(s, o, max_steps=max(1, warmup))  # untimed: caches, allocator, page faults
    t0 = time.perf_counter()
    fo.time_integration(s, o, max_steps=steps)
    return time.perf_counter() - t0


def cpu_oracle_rate(n, steps, procs, warmup=1):
    """cell-updates/s of the NumPy oracle: `procs` independent replicas of the same workload at
    n^3 (throughput of all replicas over the slowest replica's wall time); `warmup` untimed steps first."""
    if procs == 1:
        dt = _oracle_worker((n, steps, warmup))
        return n ** 3 * steps / dt
    with mp.get_context("spawn").Pool(procs) as pool:
        times = pool.map(_oracle_worker, [(n, steps, warmup)] * procs)
    return procs * n ** 3 * steps / max(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n = 48
    # each "step" of this arm = one oracle time step of every replica (a bounded sample of the 256^3
    # workload: cores x 48^3 cells); W untimed warm-up steps, then exactly K timed steps (K <= 40 keeps
    # the run within a minute or two at ~0.5 s per step)
    steps, warmup = max(1, min(args.steps, 40)), max(1, args.warmup)
    t0 = time.perf_counter()
    rate = cpu_oracle_rate(n, steps, cores, warmup)
    wall = time.perf_counter() - t0
    sample = (f"NumPy oracle (restatement of the reference; JAX not installable here), {cores} independent "
              f"replicas of the 3D MHD blast at {n}^3 fp32, {steps} time steps each, one process per core")
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * cores * n ** 3 / rate, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"3D MHD blast, bounded sample of the native arm's workload: {cores} replicas of "
                                   f"{n}^3 cells (one per host core), fp32, HLL, minmod, unsplit RK2-SSP, Pang-Wu "
                                   "implicit-midpoint magnetic step, periodic, C_cfl=0.4",
                       "cells_per_step": cores * n ** 3, "native_workload": workload_config()["workload"]},
            "wall_s": wall,
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ----------------------------------------------------------------------------- native arm
def run_native(args):
    import ctypes as C
    import torch
    import torch.distributed as dist
    from astronomix_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch N>1 with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    _lib.require_device()
    lib = _lib.load()
    K, W = args.steps, args.warmup
    n = args.cells
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak, peak_src = (peaks["hbm_gbs"], "measured") if "hbm_gbs" in peaks else (6650.0, "fallback")

    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        from astronomix_b200 import slab
        result = slab.bench_weak(n, K, W, dev)  # times are already the max over ranks
        ms, e2e_ms = result["ms"], result["e2e_ms"]
        if rank == 0:
            cells = result["cells"]
            cfg = workload_config(n)
            if result["workload"]:
                cfg = {"workload": result["workload"], "cells_per_gpu": cells // world,
                       "l2_policy": "inputs larger than L2 (state 4.3 GB per GPU)"}
            else:
                cfg["decomposition"] = f"slab along x, {n}^3 per rank, global {world * n}x{n}x{n}"
            line = {"metric": METRIC, "value": cells * K / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K,
                    "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
                    "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                    "config": dict(cfg, **result["config"]),
                    "e2e": {"value": cells * K / (e2e_ms * 1e-3), "unit": UNIT,
                            "h2d_bytes_per_step": result["h2d_bytes"] * world / K,
                            "d2h_bytes_per_step": result["d2h_bytes"] * world / K,
                            "call": f"per rank: pinned host slab -> device, {K} steps of SlabDriver.step, device -> host"},
                    "gpu_launches": result["launches"], "clocks": result["clocks"],
                    "roofline": result["roofline"], "kernels": result["kernels"],
                    "magnetic_iterations_per_step": result["mag_iters"] / K}
            print(json.dumps(line))
        dist.barrier()
        dist.destroy_process_group()
        return

    config, params, rv, state = make_problem(n)
    host = torch.from_numpy(state).pin_memory()
    s0 = host.to(dev, non_blocking=True)
    desc = _lib.make_desc(config, params, state.shape[1:], _lib.AMX_F32)
    nbytes = C.c_size_t()
    _lib.check(lib.amx_workspace_bytes(C.byref(desc), C.byref(nbytes)))
    ws = torch.empty(nbytes.value + 256, dtype=torch.uint8, device=dev)
    ws_ptr = C.c_void_p((ws.data_ptr() + 255) & ~255)
    s1, s2 = torch.empty_like(s0), torch.empty_like(s0)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    stats = _lib.AmxStats()

    def integrate(src, dst, steps):
        _lib.check(lib.amx_integrate(C.byref(desc), C.c_void_p(src.data_ptr()), C.c_void_p(dst.data_ptr()),
                                     1.0e9, steps, ws_ptr, stream, C.byref(stats)))
        return stats.steps, stats.magnetic_iterations, stats.kernel_launches

    # ---- device-resident throughput: W warm-up steps, then blocks of exactly K timed steps (each block one
    # amx_integrate call from the same warmed-up state s1) until the timed region is >= 1 s
    integrate(s0, s1, W)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    integrate(s1, s2, K)  # untimed probe block: how many blocks fill a second
    ev1.record()
    torch.cuda.synchronize()
    blocks = max(1, min(200, int(np.ceil(1000.0 / max(ev0.elapsed_time(ev1), 1e-3)))))
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    torch.cuda.synchronize()
    t0 = time.time()
    ev0.record()
    launches = 0
    for _ in range(blocks):
        steps_done, mag_iters, nl = integrate(s1, s2, K)
        launches += nl
        assert steps_done == K, steps_done
    ev1.record()
    torch.cuda.synchronize()
    t1 = time.time()
    clocks = sampler.stop(t0, t1)
    ms = ev0.elapsed_time(ev1) / blocks  # per block of K steps
    launches //= blocks
    # ---- the same K steps again with per-kernel-class CUDA events on the launch stream (the events
    # add a little launch overhead, so this pass explains `value` but is not the one that is reported)
    lib.amx_profile_enable(1)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    steps_p, mag_iters_p, _ = integrate(s1, s2, K)
    p1.record()
    torch.cuda.synchronize()
    ms_prof = p0.elapsed_time(p1)
    assert (steps_p, mag_iters_p) == (steps_done, mag_iters)
    kernels = {}
    for k, name in enumerate(KERNEL_CLASSES):
        t, c = C.c_double(), C.c_int64()
        lib.amx_profile_read(k, C.byref(t), C.byref(c))
        if c.value:
            kernels[name] = {"launches": c.value, "ms_total": t.value, "share_of_step": t.value / ms_prof}
    lib.amx_profile_enable(0)
    cells = n ** 3
    value = cells * K / (ms * 1e-3)

    # ---- roofline of the dominant kernel class (algorithmic bytes: DESIGN.md / SURVEY.md §8d)
    w, ng = 4, 5
    alg = {"hydro_stage": 0.5 * (2 * ng + 3 * ng) * w * cells}  # stage 1: r ng w ng; stage 2: r 2ng w ng
    # launches that found the fixed point already converged exit at once (speculative slack of the
    # device-side loop): only launches that did work carry bytes and time (amx_stats counts them)
    pairs, singles = int(stats.pair_launches), int(stats.single_launches)
    if "magnetic_pair" in kernels and pairs:
        # two evaluations per launch, R_{k+1} kept on chip: the first pair of a step reads B,v,rho (7)
        # and writes 6 words per cell, later pairs read 13 and write 6
        kernels["magnetic_pair"]["evaluations"] = 2 * pairs
        kernels["magnetic_pair"]["launches_with_work"] = pairs
        alg["magnetic_pair"] = (K * 13 + (pairs - K) * 19) / pairs * w * cells
    if "magnetic_iter" in kernels and singles:
        kernels["magnetic_iter"]["evaluations"] = singles
        kernels["magnetic_iter"]["launches_with_work"] = singles
        first = 0 if pairs else K
        alg["magnetic_iter"] = (first * 13 + (singles - first) * 19) / singles * w * cells
    if "hydro_stage" in kernels:
        # the speculative pipeline enqueues a step or two past the end of the run; those launches exit at once
        kernels["hydro_stage"]["launches_with_work"] = 4 * K
    alg["magnetic_final"] = (6 + ng + ng + 3) * w * cells  # reads iterate + gas, writes gas (rho, v, p) + B
    if "magnetic_final" in kernels:
        # write-back launches pre-empted by an emitting pair launch exit at once
        kernels["magnetic_final"]["launches_with_work"] = max(1, int(stats.final_launches))
    for k in kernels:
        if k in alg:
            work = kernels[k].get("launches_with_work", kernels[k]["launches"])
            kernels[k]["avg_ms"] = kernels[k]["ms_total"] / work
            kernels[k]["algorithmic_gbs"] = alg[k] / (kernels[k]["avg_ms"] * 1e-3) / 1e9
            kernels[k]["frac_of_hbm"] = kernels[k]["algorithmic_gbs"] / hbm_peak
    dom = max((k for k in kernels if k in alg), key=lambda k: kernels[k]["ms_total"])
    avg_ms = kernels[dom]["avg_ms"]
    achieved = alg[dom] / (avg_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "peak_source": f"{peak_src} (MEASURED_PEAKS.json copy bandwidth)",
                "traffic": None, "algorithmic_bytes_per_launch": alg[dom], "avg_launch_ms": avg_ms,
                "step_algorithmic_bytes_per_cell": (10 * ng + 13 + 19 * (mag_iters / K)) * w,
                "step_frac_of_hbm": value * (10 * ng + 13 + 19 * (mag_iters / K)) * w / 1e9 / hbm_peak}
    if pairs:
        # temporal blocking: fused pairs never write / re-read the intermediate iterate,
        # so the step moves fewer bytes than the one-barrier-per-evaluation figure above
        moved = (10 * ng * K + 13 * K + 19 * (pairs - K) + 19 * singles) * w / K
        roofline["step_bytes_per_cell_with_fused_pairs"] = moved
        roofline["step_frac_of_hbm_moved_bytes"] = value * moved / 1e9 / hbm_peak
        roofline["note"] = ("step_frac_of_hbm counts SURVEY 8d's algorithmic bytes (one global barrier per Psi "
                            "evaluation); the fixed point runs two evaluations per launch, so the bytes "
                            "actually required are step_bytes_per_cell_with_fused_pairs")
    try:
        # dram__bytes_read.sum + dram__bytes_write.sum per launch from this round's `ncu --set full` capture
        # (tools/traffic_json.py writes the file and names the capture); a bench run is never under a profiler
        prof = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        roofline["traffic"] = prof.get(dom, {}).get("dram_bytes_per_launch")
        roofline["traffic_source"] = prof.get("source")
    except (OSError, ValueError):
        pass

    # ---- end to end at the C-ABI (what a jax.ffi custom call sees when the caller's buffers are host
    # memory): pinned host state -> H2D -> one amx_integrate call of K steps -> D2H
    out_host = torch.empty(state.shape, dtype=torch.float32, pin_memory=True)

    def e2e_once():
        d_in = host.to(dev, non_blocking=True)
        integrate(d_in, s2, K)
        out_host.copy_(s2, non_blocking=True)

    e2e_once()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(1, min(blocks, 10))
    e0.record()
    for _ in range(reps):
        e2e_once()
    e1.record()
    torch.cuda.synchronize()
    e2e_ms = e0.elapsed_time(e1) / reps
    state_bytes = state.nbytes
    del ws, s1, s2, out_host
    torch.cuda.empty_cache()

    # ---- end to end through the public Python call with a plain (pageable) NumPy state: staging through
    # pinned memory, workspace look-up, H2D, K steps, D2H and the copy into the NumPy result all inside
    from astronomix_b200 import stepping
    stepping.integrate_steps(state, config, params, rv, K)  # first call allocates the workspace / staging buffers
    tp0 = time.perf_counter()
    for _ in range(3):
        res = stepping.integrate_steps(state, config, params, rv, K)
    api_s = (time.perf_counter() - tp0) / 3
    assert isinstance(res, np.ndarray) and res.shape == state.shape
    e2e_api = {"value": cells * K / api_s, "unit": UNIT, "ms_per_call": api_s * 1e3,
               "call": f"astronomix_b200.stepping.integrate_steps(numpy_state, ..., {K}): time_integration's path "
                       "with the step count fixed; host wall clock around the call, pageable NumPy in and out"}
    stepping._ws_cache.clear()
    stepping._staging.clear()
    del res
    torch.cuda.empty_cache()

    # ---- the other configurations as secondary keys (device-resident, same method, short runs)
    secondary = None
    if not args.no_secondary:
        try:
            secondary = secondary_configs(lib, dev, hbm_peak)
        except Exception as e:  # a secondary line must never take the headline down
            secondary = {"error": repr(e)}

    # ---- CPU baseline beside it (bounded sample of the same workload)
    cpu = None
    if not args.no_cpu:
        nc, sc = 64, 16  # about 10 s of single-core work
        t0c = time.perf_counter()
        rate = cpu_oracle_rate(nc, sc, 1)
        cpu = {"value": rate, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"NumPy oracle, same 3D MHD blast at {nc}^3 fp32, {sc} steps, single process "
                         f"({time.perf_counter() - t0c:.1f} s)"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(n),
            "timed_blocks": blocks, "timed_region_ms": ms * blocks,
            "e2e": {"value": cells * K / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": state_bytes / K,
                    "d2h_bytes_per_step": state_bytes / K,
                    "call": f"C-ABI: pinned host state -> device, one amx_integrate call of {K} steps (workspace "
                            "allocated once outside), device -> pinned host; copies amortised over the call's "
                            f"{K} steps, as in the API's natural use; mean of {reps} calls"},
            "e2e_public_api": e2e_api,
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "kernels": kernels, "ms_per_step_profiled_pass": ms_prof / K,
            "magnetic_iterations_per_step": mag_iters / K,
            "arithmetic": "fast" if desc.arithmetic == 0 else "strict", "secondary": secondary}
    print(json.dumps(line))


def secondary_configs(lib, dev, hbm_peak):
    """Device-resident throughput of the other BASELINE.json configurations and of the published benchmark's
    flux, same timing method as the headline (CUDA events around one amx_integrate call after a warm-up call);
    frac = cell-updates/s x SURVEY 8d's algorithmic bytes per cell-update / measured HBM bandwidth."""
    import ctypes as C
    import torch
    from astronomix_b200 import _lib, option_classes as oc, problems as pr

    out = {}

    def run(key, name, state, config, params, steps, warm, alg_bytes, published=None):
        params = params._replace(t_end=1.0e9)
        s0 = torch.from_numpy(state).to(dev) if isinstance(state, np.ndarray) else state
        desc = _lib.make_desc(config, params, tuple(s0.shape[1:]), _lib.AMX_F32)
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
        integ(s1, s2, steps)  # untimed block: the evaluation count the speculative pipeline enqueues has settled
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 2
        e0.record()
        for _ in range(reps):
            integ(s1, s2, steps)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        cells = int(np.prod(s0.shape[1:]))
        it = stats.magnetic_iterations / steps
        v = cells * steps / (ms * 1e-3)
        b = alg_bytes(it)
        out[key] = {"workload": name, "value": v, "unit": UNIT, "steps": steps, "ms_per_step": ms / steps,
                    "magnetic_iterations_per_step": it, "algorithmic_bytes_per_cell_update": b,
                    "frac_of_hbm": v * b / 1e9 / hbm_peak}
        if published:
            out[key]["published"] = published
        del ws, s0, s1, s2
        torch.cuda.empty_cache()

    w = 4
    c, p, rv = pr.make(1, 101, boundaries="open", riemann_solver=oc.HLLC)
    run("C1_sod_101", "C1 Sod 101 cells, HLLC, open (launch-bound)", pr.sod(101, np.float32), c, p, 200, 20,
        lambda it: 5 * 3 * w)
    c, p, rv = pr.make(2, 1024, riemann_solver=oc.HLLC)
    run("C2_kelvin_helmholtz_1024", "C2 Kelvin-Helmholtz 1024^2 hydro, HLLC, RK2 (state 16 MB: L2-resident)",
        pr.kelvin_helmholtz(1024, np.float32), c, p, 100, 10, lambda it: 5 * 4 * w)
    c, p, rv = pr.make(2, 2048, mhd=True, box=2 * np.pi)
    run("C3_orszag_tang_2048", "C3 Orszag-Tang 2048^2 MHD (Pang-Wu), HLL, RK2 (state 117 MB: about L2 size)",
        pr.orszag_tang(2048, np.float32), c, p, 40, 5, lambda it: (10 * 4 + 13 + 19 * it) * w)
    # the published benchmark, like for like (2.97e8 cell-updates/s at 256^3, 3.11e8 at 512^3 on H200,
    # jax): arena/fv_mhd.py:36-49 -- LAX_FRIEDRICHS, PERIODIC_ROLL, C_cfl 0.8, the default limiter (minmod), single
    # precision; arena_tests/scaling/scaling.py:108-130 -- blast with p 100 : 1 inside r 0.125, B0 = 10
    for nn, pub, steps in ((256, 2.97e8, 10), (512, 3.11e8, 4)):
        c, p, rv = pr.make(3, nn, mhd=True, riemann_solver=oc.LAX_FRIEDRICHS, boundary_handling=oc.PERIODIC_ROLL)
        p = p._replace(C_cfl=0.8)
        s0 = torch.empty((8, nn, nn, nn), dtype=torch.float32, device=dev)
        blast_on_device(s0, nn, p_in=100.0, p_out=1.0, r_inj=0.125, B0=10.0)
        run(f"published_config_lf_{nn}", f"3-D MHD blast {nn}^3 in the configuration of the reference's published fv_mhd "
            "benchmark: LAX_FRIEDRICHS, minmod, PERIODIC_ROLL, C_cfl 0.8, fp32", s0, c, p, steps, 2,
            lambda it: (10 * 5 + 13 + 19 * it) * w,
            published={"value": pub, "unit": UNIT, "hardware": "H200, jax (BASELINE.md)", "note": "other hardware"})
        del s0
    c, p, rv = pr.make(3, 256, mhd=True, split=oc.SPLIT, time_integrator=oc.MUSCL)
    run("C4_split_muscl_256", "C4 variant: 3-D MHD blast 256^3, Strang-split MUSCL-Hancock sweeps",
        pr.mhd_blast3d(256, np.float32), c, p, 5, 2, lambda it: (2 * 5 * 2 * 5 + 13 + 19 * it) * w)
    # single-GPU point of the north-star weak series (N > 1 runs 512^3 cells per GPU)
    c, p, rv = pr.make(3, 512, mhd=True)
    s512 = torch.empty((8, 512, 512, 512), dtype=torch.float32, device=dev)
    blast_on_device(s512, 512)
    run("weak_series_512_per_gpu", "3-D MHD blast 512^3 on one GPU (N = 1 point of the 512^3-per-GPU weak series)",
        s512, c, p, 5, 2, lambda it: (10 * 5 + 13 + 19 * it) * w)
    return out


def blast_on_device(s, n, p_in=1000.0, p_out=0.1, r_inj=0.1, B0=None):
    """problems.mhd_blast3d(n) built on the device (512^3 is 4.3 GB: not worth a host array)."""
    import torch
    B0 = 100 / np.sqrt(4 * np.pi) if B0 is None else B0
    c = (torch.arange(n, device=s.device, dtype=torch.float64) + 0.5) / n - 0.5
    c2 = c * c
    r2 = c2[:, None, None] + c2[None, :, None] + c2[None, None, :]
    s[0].fill_(1.0)
    s[1:4].zero_()
    s[4].copy_(torch.where(r2 < r_inj ** 2, p_in, p_out))
    del r2
    theta, phi = np.pi / 2, np.pi / 4
    s[5].fill_(float(B0 * np.sin(theta) * np.cos(phi)))
    s[6].fill_(float(B0 * np.sin(theta) * np.sin(phi)))
    s[7].fill_(float(B0 * np.cos(theta)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--cells", type=int, default=N_CELLS, help="cells per axis per GPU")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary configurations")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
