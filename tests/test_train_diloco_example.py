"""BASELINE config 3 in miniature: train_diloco.py (Streaming DiLoCo, 2 fragments) with 2 replica groups on
CPU/gloo against a real Lighthouse; both must commit the same number of outer steps."""

import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_train_diloco_streaming_two_replicas(tmp_path):
    from torchft_b200.bench_utils import loopback
    from torchft_b200.coordination import LighthouseServer, wait_for_lighthouse

    lh = LighthouseServer(bind="[::]:0", min_replicas=2, join_timeout_ms=2000)
    addr = loopback(lh.address())
    wait_for_lighthouse(addr)
    procs = []
    try:
        for g in range(2):
            env = dict(os.environ, TORCHFT_LIGHTHOUSE=addr, REPLICA_GROUP_ID=str(g), NUM_REPLICA_GROUPS="2", MIN_REPLICAS="2",
                       USE_CPU="1", CUDA_VISIBLE_DEVICES="", TRAIN_STEPS="40", SYNC_EVERY="10", USE_STREAMING="1",
                       N_FRAGMENTS="2", LOGLEVEL="WARNING", OMP_NUM_THREADS="1")
            procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "train_diloco.py")], env=env,
                                          stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
        outs = [p.communicate(timeout=240)[0] for p in procs]
        assert all(p.returncode == 0 for p in procs), "\n".join(o[-1500:] for o in outs)
        finals = [json.loads(o.strip().splitlines()[-1]) for o in outs]
        # 2 fragments, each synced every 10 inner steps (staggered): 40 inner steps -> 8 committed outer steps
        assert finals[0]["outer_steps"] == finals[1]["outer_steps"] == 8, finals
        assert all("participants=2" in o for o in outs)
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
        lh.shutdown()
