"""Multi-GPU parity (needs >= 2 B200s on the box; skipped otherwise): runs tools/gpu_multi_check.py under torchrun,
one rank per GPU -- slice-sharded (tne_tree.seg_shard, NCCL inside the library) and term-sharded energy + gradient
of a 4x4 D=3 PEPS against the oracle at 1e-10."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_ranks_match_the_oracle():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(29700 + os.getpid() % 200), os.path.join(ROOT, "tools", "gpu_multi_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "MULTI-GPU CHECK world 2" in res.stdout
