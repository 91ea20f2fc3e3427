"""torchrun --nproc-per-node 2 tools/slab_shell_debug.py : 3-D reflective unsplit MHD (ghost-shell mode): the slab ranks'
shells after one step against the ghost cells of the single-GPU seam step (amx_step carries the shell in seam_out)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import astronomix_b200 as ax
from astronomix_b200 import _lib, option_classes as oc, problems as pr, stepping as ti
from astronomix_b200.slab import SlabDriver, NG

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
world = dist.get_world_size()
dim, n = 3, 32
s = pr.smooth_random(dim, n, mhd=True, seed=3, dtype=np.float32)
config, params, rv = pr.make(dim, n, mhd=True, riemann_solver=oc.HLL, boundaries="reflective")
helper = ax.get_helper_data(config)
sd = torch.from_numpy(s).to(dev)
pad0 = ti._boundary_handler(ti._pad(sd, config), config)
dt = ti._cfl_time_step(pad0, config.grid_spacing, params.dt_max, params.gamma, config, rv, params.C_cfl)
pad1 = ti._evolve_state_fv(pad0, dt, params.gamma, 0.0, config, params, helper, rv)
drv = SlabDriver(config, params, rv, (n,) * dim, torch.float32, dev)
drv.scatter(sd)
nxl = drv.nxl
x0 = rank * nxl
def cmp(tag, ref_pad):
    # this rank's part of the padded reference: padded x index = x0 + local padded index
    ref = ref_pad[:5, x0:x0 + nxl + 2 * NG]
    sh = drv.shell
    mask = torch.ones_like(sh[0], dtype=torch.bool)
    mask[NG:NG + nxl, NG:-NG, NG:-NG] = False          # interior: not shell
    if rank > 0: mask[:NG] = False                      # halo planes belong to the neighbour
    if rank < world - 1: mask[NG + nxl:] = False
    d = (sh - ref).abs() * mask
    w = np.unravel_index(int(torch.argmax(d)), d.shape)
    print(f"rank {rank} {tag}: max shell diff {float(d.max()):.3e} at (row, x, y, z) = {w} of padded local shape {tuple(sh.shape)}; "
          f"per row {[('%.1e' % float(d[r].max())) for r in range(5)]}", flush=True)
cmp("initial", pad0)
drv.step()
print(f"rank {rank}: dt slab {float(drv.dt):.9e} single {float(dt):.9e}", flush=True)
cmp("after step 1", pad1)
act = drv.S[drv.cur][:, NG:NG + nxl]
print(f"rank {rank}: interior diff {float((act - pad1[:, NG + x0:NG + x0 + nxl, NG:-NG, NG:-NG]).abs().max()):.3e}", flush=True)
dist.barrier()
dist.destroy_process_group()
