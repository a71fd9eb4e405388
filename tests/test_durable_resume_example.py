"""Whole-job restart: both replica groups of train_ddp.py stop, a new job with the same CKPT_DIR resumes from the
last durable checkpoint (not from step 0) and finishes with identical weights on both groups."""

import os
import re
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run_job(addr, ckpt_dir, steps, out_dir):
    procs = []
    for g in range(2):
        env = dict(os.environ, TORCHFT_LIGHTHOUSE=addr, REPLICA_GROUP_ID=str(g), NUM_REPLICA_GROUPS="2", MIN_REPLICAS="2", USE_CPU="1",
                   CUDA_VISIBLE_DEVICES="", TRAIN_STEPS=str(steps), TRAIN_OUT=os.path.join(out_dir, f"final_{g}.pt"),
                   CKPT_DIR=ckpt_dir, CKPT_EVERY="40", LOGLEVEL="WARNING", OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "train_ddp.py")], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(o[-1500:] for o in outs)
    return outs


def test_whole_job_restart_resumes_from_durable_checkpoint(tmp_path):
    from torchft_b200.bench_utils import loopback
    from torchft_b200.coordination import LighthouseServer

    ckpt_dir = str(tmp_path / "ckpt")
    for phase, steps in enumerate((90, 130)):
        # a NEW Lighthouse per phase, as after a real outage; clock-based ids keep store prefixes unique
        lh = LighthouseServer(bind="[::]:0", min_replicas=2, join_timeout_ms=2000, quorum_id_base=-1)
        try:
            outs = _run_job(loopback(lh.address()), ckpt_dir, steps, str(tmp_path))
        finally:
            lh.shutdown()
        resumed = [re.search(r"resumed_from_step=(\w+)", o).group(1) for o in outs]
        if phase == 0:
            assert resumed == ["None", "None"]
            assert sorted(f for f in os.listdir(ckpt_dir) if f.endswith(".pt")) == ["step_80.rank_0.pt", "step_90.rank_0.pt"]
        else:
            assert resumed == ["90", "90"], resumed  # both groups restored the forced end-of-job checkpoint
            assert all(f'"final_step": {steps}' in o for o in outs)
            assert all(int(re.findall(r"step=(\d+)", o)[0]) >= 90 for o in outs)  # nobody replayed from zero
    a, b = (torch.load(tmp_path / f"final_{g}.pt", weights_only=False) for g in range(2))
    assert a["step"] == b["step"] == 130
    assert all(torch.equal(a["model"][k], b["model"][k]) for k in a["model"])
