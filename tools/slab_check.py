"""torchrun --nproc-per-node N tools/slab_check.py : N-GPU slab run == 1-GPU amx_integrate, bitwise
(STRICT arithmetic: one operation order in every kernel family)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from astronomix_b200 import _lib
from astronomix_b200 import option_classes as oc
from astronomix_b200 import problems as pr
from astronomix_b200 import stepping as ti
from astronomix_b200.slab import SlabDriver

rank = int(os.environ["RANK"])
local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ok = True
for dim, n, mhd, rs, steps in [(3, 64, True, oc.HLL, 6), (2, 512, True, oc.HLL, 6), (3, 64, False, oc.HLLC, 6),
                               (3, 64, True, oc.LAX_FRIEDRICHS, 4), (3, 64, True, oc.HLLC_LM, 3),
                               (3, 32, True, (oc.HLL, "reflective"), 5), (2, 64, True, (oc.HLL, "reflective"), 5),
                               (3, 32, True, (oc.LAX_FRIEDRICHS, "reflective"), 4), (3, 32, True, (oc.HLL, "open"), 4),
                               (2, 64, True, (oc.HLL, "jet"), 4), (2, 64, False, (oc.HLLC, "jet"), 4),
                               (2, 64, True, (oc.HLL, "jet_split"), 3)]:
    s = pr.smooth_random(dim, n, mhd=mhd, seed=3, dtype=np.float32)
    bnd = "periodic"
    if isinstance(rs, tuple):  # (Riemann solver, boundaries): reflective sides in unsplit MHD = ghost-shell mode
        rs, bnd = rs
    kw = {}
    if bnd in ("jet", "jet_split"):  # MHD jet inflow: constants in a window of the x-left ghost rows (rank 0)
        b1 = oc.BoundarySettings1D
        if bnd == "jet_split":
            kw = dict(split=oc.SPLIT, time_integrator=oc.MUSCL)
        bnd = oc.BoundarySettings(b1(oc.REFLECTIVE_BOUNDARY, oc.OPEN_BOUNDARY), b1(oc.MHD_JET_BOUNDARY, oc.REFLECTIVE_BOUNDARY),
                                  b1(oc.OPEN_BOUNDARY, oc.OPEN_BOUNDARY))
    config, params, rv = pr.make(dim, n, mhd=mhd, riemann_solver=rs, boundaries=bnd, **kw)
    bnd = bnd if isinstance(bnd, str) else "jet"
    drv = SlabDriver(config, params, rv, (n,) * dim, torch.float32, dev, arithmetic=_lib.AMX_ARITH_STRICT)
    drv.scatter(torch.from_numpy(s))
    info = drv.run(max_steps=steps)
    out = drv.gather().cpu().numpy()
    if rank == 0:
        ref = ti.integrate_steps(s, config, params, rv, steps, arithmetic=_lib.AMX_ARITH_STRICT)
        same = np.array_equal(out, ref)
        ok &= same and info.magnetic_iterations == ti.last_info.magnetic_iterations
        print(f"dim={dim} n={n} mhd={mhd} rs={rs} {bnd}: {dist.get_world_size()}-GPU slab == 1-GPU bitwise: {same}; "
              f"fixed-point iterations {info.magnetic_iterations} vs {ti.last_info.magnetic_iterations}; "
              f"max|diff| = {np.abs(out - ref).max():.3e}", flush=True)
    dist.barrier()
# FAST arithmetic (the bench's), the 3-D MHD blast, through the public entry point with sharding=
import astronomix_b200 as ax  # noqa: E402
n = int(os.environ.get("AMX_CHECK_N", "128"))
s = pr.mhd_blast3d(n, np.float32)
config, params, rv = pr.make(3, n, mhd=True)
params = params._replace(t_end=2.0e-3)
out = ax.time_integration(s, config, params, rv, sharding=ax.SlabSharding())
steps_n, iters_n = ti.last_info.steps, ti.last_info.magnetic_iterations
if rank == 0:
    ref = ax.time_integration(s, config, params, rv)
    same = np.array_equal(out, ref) and (steps_n, iters_n) == (ti.last_info.steps, ti.last_info.magnetic_iterations)
    ok &= same
    print(f"time_integration(sharding=SlabSharding()), FAST, blast {n}^3 to t=2e-3: {dist.get_world_size()}-GPU == "
          f"1-GPU bitwise: {same}; steps {steps_n} vs {ti.last_info.steps}, fixed-point iterations {iters_n} vs "
          f"{ti.last_info.magnetic_iterations}", flush=True)
dist.barrier()
# driven turbulence (row f1) through the public entry point: slab ranks draw their part of the Threefry noise,
# distributed inverse FFT, all-reduced sums -- equal to one GPU to round-off (another FFT factorisation)
n = 64
tfc = oc.TurbulentForcingConfig(turbulent_forcing=True)
config, params, rv = pr.make(3, n, mhd=True, boundary_handling=oc.PERIODIC_ROLL, turbulent_forcing_config=tfc)
params = params._replace(t_end=0.01)
s = np.zeros((8, n, n, n), dtype=np.float32)
s[0], s[4], s[5], s[6] = 0.5, 0.1, 0.05, 0.05
out = ax.time_integration(s, config, params, rv, sharding="x")
steps_n = ti.last_info.steps
if rank == 0:
    ref = ax.time_integration(s, config, params, rv)
    err = pr.rel_l1(out, ref, 3, True)
    good = err <= 1e-5 and steps_n == ti.last_info.steps and np.isfinite(out).all()
    ok &= bool(good)
    print(f"driven turbulence {n}^3 to t=0.01, FAST: {dist.get_world_size()}-GPU vs 1-GPU rel-L1 = {err:.3e} "
          f"(bar 1e-5), steps {steps_n} vs {ti.last_info.steps}: {good}", flush=True)
dist.barrier()
if rank == 0:
    print("SLAB_CHECK", "PASS" if ok else "FAIL")
dist.destroy_process_group()
