"""Row-sharded multi-GPU parity (NCCL).  Needs >= 2 GPUs on the box; skipped otherwise."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_fit_matches_oracle(salg):
    n = salg.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tests", "dist_gpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    print(r.stdout[-6000:])          # (pytest -s: the per-check lines of both ranks go to the committed log)
    assert r.returncode == 0 and "DIST_GPU_CHECK_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
