"""GPU (>= 2 devices): the frame-sharded NCCL solve agrees with the single-GPU solve."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", ["peer", "nccl", "peer_unavailable_on_rank1"])
def test_sharded_solve_matches_single_gpu(mode):
    """The reduced system is summed by the peer-memory kernel (default), by NCCL (MCBA_NO_PEER=1), and
    by NCCL on every rank when one rank cannot set the peer buffers up (collective fallback)."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29613",
           os.path.join(ROOT, "scripts", "multigpu_check.py"), "997"]
    env = dict(os.environ)
    if mode == "nccl":
        env["MCBA_NO_PEER"] = "1"
    elif mode == "peer_unavailable_on_rank1":
        env["MCBA_TEST_PEER_FAIL"] = "1"
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "MULTIGPU_OK" in out.stdout
    assert ("peer_memory=True" in out.stdout) == (mode == "peer")
