"""BASELINE config 1: train_ddp.py toy CNN, 2 replica groups x world_size 1 on CPU/gloo with a
Lighthouse; one group is SIGKILLed mid-run and restarted; both must finish with identical weights."""

import os
import signal
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _spawn(group, lh_addr, out, steps):
    env = dict(os.environ, TORCHFT_LIGHTHOUSE=lh_addr, REPLICA_GROUP_ID=str(group), NUM_REPLICA_GROUPS="2", USE_CPU="1",
               TRAIN_STEPS=str(steps), TRAIN_OUT=out, LOGLEVEL="WARNING", OMP_NUM_THREADS="1")
    return subprocess.Popen([sys.executable, os.path.join(ROOT, "train_ddp.py")], env=env, stdout=subprocess.PIPE,
                            stderr=subprocess.STDOUT, text=True)


def test_train_ddp_kill_and_rejoin(tmp_path):
    from torchft_b200.bench_utils import loopback
    from torchft_b200.coordination import LighthouseServer

    lh = LighthouseServer(bind="[::]:0", min_replicas=1, join_timeout_ms=1500, heartbeat_timeout_ms=1500)
    addr = loopback(lh.address())
    outs = [str(tmp_path / f"g{i}.pt") for i in range(2)]
    steps = 400
    procs = [_spawn(i, addr, outs[i], steps) for i in range(2)]
    try:
        # let them train together for a bit, then kill group 1 hard (well before group 0 can finish)
        time.sleep(10)
        assert procs[1].poll() is None and procs[0].poll() is None, "training finished before the kill"
        procs[1].send_signal(signal.SIGKILL)
        procs[1].wait()
        time.sleep(3)
        procs[1] = _spawn(1, addr, outs[1], steps)
        for p in procs:
            p.wait(timeout=300)
        logs = [p.stdout.read() for p in procs]
        assert all(p.returncode == 0 for p in procs), logs
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
        lh.shutdown()
    a, b = torch.load(outs[0]), torch.load(outs[1])
    assert a["step"] == b["step"] == steps
    for k in a["model"]:
        torch.testing.assert_close(a["model"][k], b["model"][k])
