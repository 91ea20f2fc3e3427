"""N-GPU == 1-GPU on real devices (self-skips below 2 GPUs): tools/slab_check.py under torch.distributed.run --
STRICT family cases and the FAST 3-D MHD blast through time_integration(sharding=SlabSharding()), with the
halo planes stored by the kernels themselves and the MAX all-reduces through peer memory."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("push", ["1", "0"])
def test_slab_equals_single_gpu_bitwise(push):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    n = min(torch.cuda.device_count(), 4)
    env = dict(os.environ, AMX_SLAB_PUSH=push, AMX_CHECK_N="128")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.join(ROOT, "tools", "slab_check.py")],
                       capture_output=True, text=True, timeout=600, env=env)
    assert "SLAB_CHECK PASS" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
