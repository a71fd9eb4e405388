"""BASELINE config 1: train_ddp.py toy CNN, 2 replica groups x world_size 1 on CPU/gloo with a
Lighthouse; one group is SIGKILLed mid-run and restarted; it must heal live from the survivor and
both must finish with identical weights."""

import os
import re
import signal
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STEPS = 4000


def _spawn(group, lh_addr, out, log):
    env = dict(os.environ, TORCHFT_LIGHTHOUSE=lh_addr, REPLICA_GROUP_ID=str(group), NUM_REPLICA_GROUPS="2", USE_CPU="1",
               TRAIN_STEPS=str(STEPS), TRAIN_OUT=out, LOGLEVEL="WARNING", OMP_NUM_THREADS="1")
    return subprocess.Popen([sys.executable, os.path.join(ROOT, "train_ddp.py")], env=env, stdout=open(log, "a"),
                            stderr=subprocess.STDOUT)


def _last_step(log):
    try:
        steps = re.findall(r"step=(\d+)", open(log).read())
        return int(steps[-1]) if steps else 0
    except FileNotFoundError:
        return 0


def _wait_for(pred, timeout, what):
    t0 = time.time()
    while time.time() - t0 < timeout:
        if pred():
            return
        time.sleep(0.1)
    raise AssertionError(f"timed out waiting for {what}")


def test_train_ddp_kill_and_rejoin(tmp_path):
    from torchft_b200.bench_utils import loopback
    from torchft_b200.coordination import LighthouseServer

    lh = LighthouseServer(bind="[::]:0", min_replicas=1, join_timeout_ms=1500, heartbeat_timeout_ms=1500)
    addr = loopback(lh.address())
    outs = [str(tmp_path / f"g{i}.pt") for i in range(2)]
    logs = [str(tmp_path / f"g{i}.log") for i in range(2)]
    procs = [_spawn(i, addr, outs[i], logs[i]) for i in range(2)]
    try:
        _wait_for(lambda: _last_step(logs[0]) >= 100 and _last_step(logs[1]) >= 100, 120, "both groups training")
        procs[1].send_signal(signal.SIGKILL)  # hard kill: no goodbye to anybody
        procs[1].wait()
        killed_at = _last_step(logs[0])
        # the survivor keeps going alone (min_replicas=1) once the lighthouse drops the dead group
        _wait_for(lambda: _last_step(logs[0]) >= killed_at + 100, 120, "survivor progress after the kill")
        assert procs[0].poll() is None, "survivor finished before the rejoin; raise STEPS"
        procs[1] = _spawn(1, addr, outs[1], logs[1])
        for p in procs:
            p.wait(timeout=400)
        assert all(p.returncode == 0 for p in procs), [open(l).read()[-3000:] for l in logs]
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
        lh.shutdown()
    a, b = torch.load(outs[0]), torch.load(outs[1])
    assert a["step"] == b["step"] == STEPS
    for k in a["model"]:
        torch.testing.assert_close(a["model"][k], b["model"][k])
    # the restarted group healed (its second incarnation logged steps far beyond where it was killed)
    second_life = re.findall(r"step=(\d+)", open(logs[1]).read())
    assert int(second_life[-1]) == STEPS
