"""Multi-GPU parity on a box with at least two GPUs (skipped otherwise): z-sharded phase retrieval through the
peer-window reduction and through NCCL, z-sharded HanserPSF, z-sharded OTFi (tools/multigpu_check.py under torchrun,
one rank per GPU).  The host-side sharding logic is covered on CPU by tests/test_gloo_sharding.py."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.timeout(900)
def test_two_rank_sharded_paths_match_single_gpu():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least two GPUs")
    port = 29600 + os.getpid() % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "multigpu_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=850)
    assert r.returncode == 0 and "MULTIGPU CHECK PASSED" in r.stdout, (r.stdout[-3000:], r.stderr[-3000:])
