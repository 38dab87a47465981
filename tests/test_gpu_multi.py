"""GPU (>= 2 devices): the multi-GPU path (block-column Cholesky with NCCL panel broadcasts, split
triangular inverse, row-cyclic W'W + trace) against the CPU oracle.  Skipped on 1-GPU boxes."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_gpu_parity(world):
    if _ngpu() < world:
        pytest.skip("needs %d GPUs" % world)
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29611 + world), os.path.join(here, "mgpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    sys.stdout.write(r.stdout[-4000:]); sys.stderr.write(r.stderr[-2000:])
    assert r.returncode == 0 and "MGPU_RESULT PASS" in r.stdout
