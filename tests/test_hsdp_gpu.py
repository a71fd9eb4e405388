"""HSDP = FSDP2 shards inside a replica group x fault-tolerant replication across groups.

The reference wires this up with ``FSDPModule.set_all_reduce_hook`` feeding a ``ManagedProcessGroup``
and only checks the plumbing with mocks (/root/reference/torchft/fsdp_test.py:24-101). Here 4 GPUs
run 2 replica groups x 2 shards for real: the cross-replica all-reduce of every reduce-scattered
gradient shard goes through Manager -> ProcessGroupB200 (fused NVLink kernel), and the final weights
must equal those of a single-process run on the concatenated batch.
"""

import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 4, reason="needs >= 4 GPUs")
def test_hsdp_two_groups_of_two_shards(tmp_path):
    out = tmp_path / "hsdp.json"
    cmd = [sys.executable, os.path.join(ROOT, "examples", "hsdp_fsdp2.py"), "--groups", "2", "--shards", "2", "--steps", "4",
           "--out", str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    res = json.loads(out.read_text())
    assert res["steps_committed"] == [4, 4, 4, 4]
    assert res["max_param_diff_vs_reference"] < 2e-3, res
    assert res["max_param_diff_between_groups"] == 0.0, res
