"""torchrun --nproc-per-node 2 tools/comm_micro.py : device-side cost of the slab driver's collectives."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from astronomix_b200 import problems as pr
from astronomix_b200.slab import SlabDriver, NG

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 256
config, params, rv = pr.make(3, n, mhd=True)

def timeit(name, fn, reps=200):
    for _ in range(10):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    if rank == 0:
        print(f"{name:50s} {e0.elapsed_time(e1) / reps * 1e3:8.1f} us per call", flush=True)

for p2p in ("1", "0"):
    os.environ["AMX_SLAB_P2P"] = p2p
    drv = SlabDriver(config, params, rv, (n * dist.get_world_size(), n, n), torch.float32, dev)
    tag = "p2p stores + barrier" if drv.p2p else "NCCL send/recv (packed)"
    timeit(f"exchange 5 rows (gas)          [{tag}]", lambda: drv.exchange(drv.G))
    timeit(f"exchange 6 rows (iterate)      [{tag}]", lambda: drv.exchange(drv.X[0]))
    timeit(f"exchange 3 rows (B)            [{tag}]", lambda: drv.exchange(drv.S[0][5:]))
    if drv.p2p:
        h = drv._symm[0][1]
        timeit("symmetric-memory barrier alone", lambda: h.barrier(channel=0))
    del drv
ch = torch.zeros(2, device=dev)
timeit("all_reduce MAX of 2 floats (NCCL)", lambda: dist.all_reduce(ch, op=dist.ReduceOp.MAX))
x = torch.zeros(1, device=dev)
timeit("empty-ish kernel (x.add_(1))", lambda: x.add_(1))
dist.destroy_process_group()
