"""Time the slab driver on ONE GPU (world = 1: ghost planes wrapped locally, no NCCL) against
amx_integrate on the same problem: isolates the cost of the per-stage driver from communication."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from astronomix_b200 import problems as pr
from astronomix_b200.slab import SlabDriver, NG

n, K, W = 256, 20, 3
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
config, params, rv = pr.make(3, n, mhd=True)
params = params._replace(t_end=1.0e9)
local = np.ascontiguousarray(pr.mhd_blast3d(n, np.float32))
drv = SlabDriver(config, params, rv, (n, n, n), torch.float32, dev)
drv.S[0][:, NG:NG + n].copy_(torch.from_numpy(local).to(dev))
drv.cur = 0
drv._alpha_of(drv.S[0][: drv.ng])
for _ in range(W):
    drv.step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
it = 0
for _ in range(K):
    it += drv.step()
e1.record()
t_cpu = time.perf_counter() - t0
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"slab driver world=1: {ms / K:.3f} ms/step (host loop returned after {t_cpu * 1e3 / K:.3f} ms/step), "
      f"{n ** 3 * K / ms / 1e6:.3f}e9 cell-updates/s, {it / K} iterations/step, {drv.collectives / (K + W):.1f} exchanges/step")
