"""Multi-GPU parity proper: one process per GPU over NCCL (tools/multigpu_check.py under torchrun), every rank's block
against the single-rank run of the undecomposed mesh (the reference's criterion, tests/test_parallel.py:63-81),
fp64 1e-10. Skipped on boxes with a single GPU; the host-side logic is covered on CPU by tests/test_multirank.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_decomposition_invariance_nccl(world, cudalib):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    port = 29600 + world + (os.getpid() % 1000)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "multigpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("maxerr") == world
