"""configs[4] on hardware: 4 Lines x 256 ch, one per GPU, through the 4-stage chain, then the fan-in sum on rank 0 over NVLink
(NCCL reduce and the peer-memory mixer kernel), both against the oracle's sum.  Needs 4 devices: skipped below that."""
import os
import subprocess
import sys

import pytest

from pipe_b200 import abi

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fan_in_sum_over_nvlink_matches_the_oracle():
    if abi.device_count() < 4:
        pytest.skip("configs[4] needs 4 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "4", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tools", "fanin_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "fan-in of 4 Lines" in res.stdout
