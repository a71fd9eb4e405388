"""Single-node orchestrator with failure injection (the role of the reference's monarch example,
``examples/monarch/train_distributed.py`` + ``utils/failure.py``, without an actor framework).

Starts a Lighthouse, launches ``--replicas`` replica groups (one torchrun each, pinned to disjoint
GPU subsets), keeps them alive (a dead group is relaunched after ``--relaunch-delay`` seconds),
and — with ``--mtbf-secs`` — injects a random failure into a random group: SEGFAULT / KILL_PROC /
COMMS / DEADLOCK / STALL_PEER go through the in-process :class:`torchft_b200.failure.FailureInjector`
of the victim (the training script must start one; ``train_ddp.py`` does when
``TORCHFT_FAILURE_PORT_FILE`` is set), KILL_GROUP kills the whole torchrun process group from
outside (the analogue of the reference's KILL_SLURM).

    python examples/orchestrator/train_orchestrated.py --replicas 2 --mtbf-secs 30 --duration 300 train_ddp.py
"""

from __future__ import annotations

import argparse
import os
import random
import re
import subprocess
import sys
import tempfile
import time
from typing import Dict, List, Optional

import psutil

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

from torchft_b200.coordination import LighthouseServer  # noqa: E402
from torchft_b200.failure import Failure, send_failure  # noqa: E402
from torchft_b200.launcher import Role, hsdp  # noqa: E402

OUTSIDE_KILL = "kill_group"


class Group:
    def __init__(self, role: Role, port_file: str, log_dir: str) -> None:
        self.role, self.port_file, self.log_dir = role, port_file, log_dir
        self.proc: Optional[subprocess.Popen] = None
        self._stragglers: List[psutil.Process] = []
        self.launches = 0
        self.died_at: Optional[float] = None

    def start(self) -> None:
        if os.path.exists(self.port_file):
            os.unlink(self.port_file)
        log = open(os.path.join(self.log_dir, f"{self.role.name}.{self.launches}.log"), "w")
        env = dict(os.environ, **self.role.env)
        env["TORCHFT_FAILURE_PORT_FILE"] = self.port_file
        # own process group so an outside kill takes torchrun and its workers together
        self.proc = subprocess.Popen([self.role.entrypoint, *self.role.args], env=env, stdout=log, stderr=subprocess.STDOUT,
                                     start_new_session=True)
        self.launches += 1
        self.died_at = None

    def alive(self) -> bool:
        return self.proc is not None and self.proc.poll() is None

    def kill(self) -> None:
        """SIGKILL the launcher AND its workers. torchrun puts every worker in its own session, so killing
        the launcher's process group alone would orphan them; walk the exact process tree we started instead."""
        if self.proc is None:
            return
        victims = []
        try:
            root = psutil.Process(self.proc.pid)
            victims = root.children(recursive=True) + [root]
        except psutil.NoSuchProcess:
            pass
        victims += self._stragglers
        for v in victims:
            try:
                v.kill()
            except psutil.NoSuchProcess:
                pass
        self._stragglers = []

    def remember_workers(self) -> None:
        """Record the live worker processes so that they can still be reaped after the launcher died
        (a worker that os._exit()s takes the launcher down; a hung sibling would otherwise survive)."""
        if self.proc is None:
            return
        try:
            self._stragglers = psutil.Process(self.proc.pid).children(recursive=True)
        except psutil.NoSuchProcess:
            pass

    def logged_step(self) -> int:
        """Highest `step=N` in the current launch's log (the training scripts print one every few steps)."""
        try:
            with open(os.path.join(self.log_dir, f"{self.role.name}.{self.launches - 1}.log"), errors="replace") as f:
                f.seek(max(0, os.path.getsize(f.name) - 4096))
                steps = re.findall(r"step=(\d+)", f.read())
            return int(steps[-1]) if steps else 0
        except (OSError, ValueError):
            return 0

    def injector_port(self) -> Optional[int]:
        try:
            with open(self.port_file) as f:
                return int(f.read().strip())
        except (OSError, ValueError):
            return None


def main() -> None:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--replicas", type=int, default=2)
    ap.add_argument("--workers-per-replica", type=int, default=1)
    ap.add_argument("--gpus-per-node", type=int, default=None)
    ap.add_argument("--min-replicas", type=int, default=1)
    ap.add_argument("--join-timeout-ms", type=int, default=5000)
    ap.add_argument("--mtbf-secs", type=float, default=0.0, help="mean time between injected failures (0 = none)")
    ap.add_argument("--failures", default="kill_proc,segfault,comms,kill_group",
                    help="comma list from: " + ",".join([f.value for f in Failure] + [OUTSIDE_KILL]))
    ap.add_argument("--max-failures", type=int, default=0, help="stop injecting after this many failures (0 = unlimited)")
    ap.add_argument("--inject-below-min", action="store_true",
                    help="also inject when that takes the job below --min-replicas (survivors then stall until the victim is back)")
    ap.add_argument("--stop-injecting-at-step", type=int, default=0,
                    help="no more failures once any group logged `step=N` with N >= this (0 = never stop); lets a soak "
                         "end with all groups in one quorum so their final weights can be compared")
    ap.add_argument("--relaunch-delay", type=float, default=2.0)
    ap.add_argument("--duration", type=float, default=0.0, help="stop after this many seconds (0 = until all groups exit 0)")
    ap.add_argument("--log-dir", default=None)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("script")
    ap.add_argument("script_args", nargs="*")
    a = ap.parse_args()

    rng = random.Random(a.seed)
    log_dir = a.log_dir or tempfile.mkdtemp(prefix="tft_orch_")
    os.makedirs(log_dir, exist_ok=True)
    lighthouse = LighthouseServer(bind="127.0.0.1:0", min_replicas=a.min_replicas, join_timeout_ms=a.join_timeout_ms,
                                  quorum_id_base=-1)  # clock-based ids: a restarted orchestrator never reuses one
    roles = hsdp(*a.script_args, replicas=a.replicas, workers_per_replica=a.workers_per_replica, max_restarts=0,
                 script=a.script, lighthouse=lighthouse.address(), gpus_per_node=a.gpus_per_node)
    groups: List[Group] = [Group(r, os.path.join(log_dir, f"{r.name}.port"), log_dir) for r in roles]
    for g in groups:
        g.start()
    print(f"lighthouse {lighthouse.address()}  logs {log_dir}", flush=True)

    kinds = [k.strip() for k in a.failures.split(",") if k.strip()]
    t0 = time.monotonic()
    next_failure = t0 + rng.expovariate(1.0 / a.mtbf_secs) if a.mtbf_secs > 0 else float("inf")
    stats: Dict[str, int] = {}
    done: Dict[str, int] = {}
    try:
        while True:
            time.sleep(0.5)
            now = time.monotonic()
            for g in groups:
                if g.role.name not in done and g.alive():
                    g.remember_workers()
                if g.role.name in done or g.alive():
                    continue
                assert g.proc is not None
                if g.proc.returncode == 0:
                    done[g.role.name] = 0
                    print(f"{g.role.name} finished", flush=True)
                elif g.died_at is None:
                    g.died_at = now
                    print(f"{g.role.name} died rc={g.proc.returncode}; relaunch in {a.relaunch_delay}s", flush=True)
                elif now - g.died_at >= a.relaunch_delay:
                    g.kill()  # reap any worker that outlived its launcher
                    g.start()
                    print(f"{g.role.name} relaunched (launch #{g.launches})", flush=True)
            if len(done) == len(groups) or (a.duration and now - t0 >= a.duration):
                break
            if now >= next_failure:
                next_failure = now + rng.expovariate(1.0 / a.mtbf_secs)
                if a.max_failures and sum(stats.values()) >= a.max_failures:
                    next_failure = float("inf")
                    continue
                if a.stop_injecting_at_step and max(g.logged_step() for g in groups) >= a.stop_injecting_at_step:
                    next_failure = float("inf")
                    print(f"[{now - t0:7.1f}s] step {a.stop_injecting_at_step} reached: no more failures", flush=True)
                    continue
                victims = [g for g in groups if g.alive() and g.role.name not in done]
                if not victims or (len(victims) <= a.min_replicas and not a.inject_below_min):
                    continue  # by default never take the job below its quorum floor
                g, kind = rng.choice(victims), rng.choice(kinds)
                stats[kind] = stats.get(kind, 0) + 1
                print(f"[{now - t0:7.1f}s] injecting {kind} into {g.role.name}", flush=True)
                if kind == OUTSIDE_KILL:
                    g.kill()
                else:
                    port = g.injector_port()
                    if port is None:
                        g.kill()
                    else:
                        try:
                            send_failure(port, Failure(kind))
                        except OSError as e:
                            print(f"  injector unreachable ({e}); killing from outside", flush=True)
                            g.kill()
    except KeyboardInterrupt:
        pass
    finally:
        for g in groups:
            g.kill()
        lighthouse.shutdown()
    print(f"injected: {stats}; launches: { {g.role.name: g.launches for g in groups} }", flush=True)


if __name__ == "__main__":
    main()
