"""ProcessGroupB200 across real processes on however many GPUs the box has (3 ranks share ONE GPU on the
single-GPU tier): exact values for every collective + kill-a-rank resiliency. See tests/_pg_b200_worker.py."""

import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [3])
def test_process_group_b200_multi_process_collectives_and_resiliency(world):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "tests", "_pg_b200_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-4000:]
    assert '"failures": 0' in r.stdout, r.stdout[-2000:]
