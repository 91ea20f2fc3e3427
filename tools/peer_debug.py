"""torchrun --nproc-per-node N tools/peer_debug.py : the peer-memory plumbing in isolation -- the MAX all-reduce
kernel on known values, then a few slab steps with AMX_SLAB_PUSH_VERIFY=1 (every pushed halo compared with the
neighbour's edge planes)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("AMX_SLAB_PUSH_VERIFY", "1")
import numpy as np
import torch
import torch.distributed as dist

from astronomix_b200 import _lib, option_classes as oc, problems as pr
from astronomix_b200.slab import SlabDriver

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
world = dist.get_world_size()
n = 64
s = pr.smooth_random(3, n, mhd=True, seed=3, dtype=np.float32)
config, params, rv = pr.make(3, n, mhd=True)
drv = SlabDriver(config, params, rv, (n,) * 3, torch.float32, dev, arithmetic=_lib.AMX_ARITH_STRICT)
print(f"rank {rank}: p2p {drv.p2p} push bits {drv.push} comm {drv._comm is not None}", flush=True)
if drv._comm is not None:
    for k in range(5):
        v = torch.tensor([rank + 1.0 + k, 10.0 - rank, 3.0, float(k)], device=dev)
        drv._allreduce_max(v)
        exp = [world + k, 10.0, 3.0, float(k)]
        print(f"rank {rank} allreduce {k}: {v.tolist()} expected {exp}", flush=True)
drv.scatter(torch.from_numpy(s))
try:
    info = drv.run(max_steps=2)
    print(f"rank {rank}: {info}", flush=True)
except Exception as e:
    print(f"rank {rank}: run failed: {e!r}", flush=True)
dist.barrier()
dist.destroy_process_group()
