"""Chaos soak on real GPUs: the orchestrator keeps 2 replica groups of train_ddp.py (ProcessGroupB200
data plane, in-place NVLink heal) training while failures are injected, then checks the outcome.

    python bench/chaos_soak.py --steps 4000 --mtbf-secs 10 --out gpurun_out/chaos_soak.json

Pass criteria: every group finishes at exactly --steps committed steps, and the final weights of all
groups are bit-identical (what the reference's integration tests assert with `assert_equal_global_state`).
"""

from __future__ import annotations

import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--replicas", type=int, default=2)
    ap.add_argument("--steps", type=int, default=4000)
    ap.add_argument("--mtbf-secs", type=float, default=10.0)
    ap.add_argument("--failures", default="kill_proc,segfault,comms,kill_group")
    ap.add_argument("--min-replicas", type=int, default=1,
                    help="1: survivors keep training alone while a victim restarts (needs enough steps after the last failure for it "
                         "to rejoin); = --replicas: survivors stall until the victim is back (deterministic end state)")
    ap.add_argument("--max-failures", type=int, default=0)
    ap.add_argument("--inject-until", type=float, default=0.5, help="fraction of --steps after which no more failures are injected")
    ap.add_argument("--timeout", type=int, default=420)
    ap.add_argument("--out", default="gpurun_out/chaos_soak.json")
    a = ap.parse_args()
    work = tempfile.mkdtemp(prefix="tft_soak_")
    env = dict(os.environ, TRAIN_STEPS=str(a.steps), TRAIN_OUT=os.path.join(work, "final_{group}.pt"), LOGLEVEL="INFO")
    cmd = [sys.executable, os.path.join(ROOT, "examples/orchestrator/train_orchestrated.py"), "--replicas", str(a.replicas),
           "--gpus-per-node", str(max(torch.cuda.device_count(), 1)), "--min-replicas", str(a.min_replicas), "--inject-below-min", "--max-failures", str(a.max_failures), "--join-timeout-ms", "2000",
           "--mtbf-secs", str(a.mtbf_secs), "--failures", a.failures, "--relaunch-delay", "1", "--log-dir", work,
           "--stop-injecting-at-step", str(int(a.steps * a.inject_until)),
           os.path.join(ROOT, "train_ddp.py")]
    t0 = time.time()
    p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=a.timeout)
    wall = time.time() - t0
    injected = re.findall(r"injecting (\w+) into (\w+)", p.stdout)
    launches = re.findall(r"launches: (\{.*\})", p.stdout)
    finals = {}
    for g in range(a.replicas):
        f = os.path.join(work, f"final_{g}.pt")
        if os.path.exists(f):
            finals[g] = torch.load(f, weights_only=False)
    same = len(finals) == a.replicas and all(
        all(torch.equal(finals[0]["model"][k], finals[g]["model"][k]) for k in finals[0]["model"]) for g in finals)
    steps = {g: int(v["step"]) for g, v in finals.items()}
    heals = 0
    for fn in os.listdir(work):
        if fn.endswith(".log"):
            heals += open(os.path.join(work, fn), errors="replace").read().count("healing required")
    res = {"replicas": a.replicas, "target_steps": a.steps, "wall_s": round(wall, 1), "orchestrator_rc": p.returncode,
           "injected": [f"{k}->{g}" for k, g in injected], "n_injected": len(injected), "launches": launches[-1] if launches else None,
           "final_steps": steps, "final_weights_identical": bool(same), "heals_logged": heals,
           "pass": p.returncode == 0 and bool(same) and all(s == a.steps for s in steps.values()) and len(steps) == a.replicas}
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    with open(a.out, "w") as f:
        json.dump(res, f, indent=1)
    print("CHAOS_SOAK " + json.dumps(res), flush=True)
    if not res["pass"]:
        print(p.stdout[-3000:], p.stderr[-2000:], flush=True)
        for fn in sorted(os.listdir(work)):
            if fn.endswith(".log"):
                print("----", fn)
                print(open(os.path.join(work, fn), errors="replace").read()[-1500:])
        sys.exit(1)


if __name__ == "__main__":
    main()
