"""Control-plane latency benchmark (CPU only): what the per-step protocol costs.

Measures, over loopback TCP with the real C++ servers and clients:
  * should_commit round trip (1 rank per group: pure RPC cost; this is on EVERY training step)
  * ManagerClient.quorum round trip in steady state (unchanged membership -> Lighthouse fast path)
  * time for a Lighthouse to form a quorum of N replica groups arriving together (N = 2..64)
  * time for the survivors to get a new quorum after one member stops (heartbeat expiry + shrink)

    python bench/control_plane_bench.py --out profiles/control_plane_bench.json
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time
from datetime import timedelta

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from torchft_b200.coordination import LighthouseClient, LighthouseServer, ManagerClient, ManagerServer  # noqa: E402

T = timedelta(seconds=30)


def pct(xs, p):
    xs = sorted(xs)
    return xs[min(len(xs) - 1, int(len(xs) * p))]


def summarize(us):
    return {"p50_us": round(statistics.median(us), 1), "p99_us": round(pct(us, 0.99), 1), "mean_us": round(statistics.fmean(us), 1), "n": len(us)}


def bench_manager_rpcs(iters: int) -> dict:
    lh = LighthouseServer(bind="127.0.0.1:0", min_replicas=1, join_timeout_ms=100, quorum_tick_ms=10)
    ms = ManagerServer(replica_id="g0", lighthouse_addr=lh.address(), hostname="127.0.0.1", bind="127.0.0.1:0",
                       store_addr="s:1", world_size=1, heartbeat_interval=timedelta(milliseconds=100),
                       connect_timeout=T, quorum_retries=0)
    c = ManagerClient(ms.address(), T)
    c._quorum(0, 0, "", False, T, 0, True)
    commit, quorum = [], []
    for i in range(iters):
        t0 = time.perf_counter()
        c.should_commit(0, i, True, T)
        commit.append((time.perf_counter() - t0) * 1e6)
    for i in range(max(20, iters // 10)):
        t0 = time.perf_counter()
        c._quorum(0, i + 1, "", False, T, 0, True)
        quorum.append((time.perf_counter() - t0) * 1e6)
    ms.shutdown()
    lh.shutdown()
    return {"should_commit_rtt": summarize(commit), "manager_quorum_rtt_steady_state": summarize(quorum)}


class Heartbeats:
    """What every ManagerServer does on a background thread: keep its replica alive at the Lighthouse."""

    def __init__(self, addr: str, ids: list, period_s: float = 0.05) -> None:
        self.live = set(ids)
        self._client = LighthouseClient(addr, T)
        self._stop = threading.Event()
        self._period = period_s
        self._beat()
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def _beat(self) -> None:
        for rid in list(self.live):
            self._client.heartbeat(rid, T)

    def _run(self) -> None:
        while not self._stop.wait(self._period):
            self._beat()

    def stop(self) -> None:
        self._stop.set()
        self._t.join()


def ask_all(clients, ids, step):
    ids = list(ids)
    """All replicas in ``ids`` request a quorum at once; returns (ms until the last one was answered, sizes)."""
    barrier = threading.Barrier(len(ids) + 1)
    done = {}

    def run(i: int) -> None:
        barrier.wait()
        q = clients[i].quorum(replica_id=f"r{i}", timeout=T, address=f"http://a{i}", store_address=f"s{i}:1", step=step, world_size=1)
        # stamp first; touch the result (a pybind conversion of all N members under the GIL: with N client threads in ONE
        # process that alone cost 100+ ms at N = 64 and was mistaken for server time) only for one replica
        done[i] = (time.perf_counter(), len(q.participants) if i == ids[0] else None)

    ts = [threading.Thread(target=run, args=(i,)) for i in ids]
    [t.start() for t in ts]
    barrier.wait()
    t0 = time.perf_counter()
    [t.join() for t in ts]
    return (max(v[0] for v in done.values()) - t0) * 1e3, sorted({v[1] for v in done.values() if v[1] is not None})


def bench_quorum_formation(n: int, rounds: int) -> dict:
    lh = LighthouseServer(bind="127.0.0.1:0", min_replicas=n, join_timeout_ms=200, quorum_tick_ms=10, heartbeat_timeout_ms=2000)
    clients = [LighthouseClient(lh.address(), T) for _ in range(n)]
    hb = Heartbeats(lh.address(), [f"r{i}" for i in range(n)])
    times = []
    for r in range(rounds):
        ms, sizes = ask_all(clients, range(n), r)
        assert sizes == [n], sizes
        times.append(ms)
    hb.stop()
    lh.shutdown()
    return {"replicas": n, "all_arrive_to_all_notified_ms": {"p50": round(statistics.median(times), 2), "max": round(max(times), 2), "n": rounds}}


def bench_shrink(n: int, heartbeat_timeout_ms: int = 500) -> dict:
    """n replicas, min_replicas=1: after one dies, how long until the survivors hold a quorum without it.
    Lower bound = heartbeat_timeout (the Lighthouse must first declare the member dead)."""
    lh = LighthouseServer(bind="127.0.0.1:0", min_replicas=1, join_timeout_ms=100, quorum_tick_ms=10,
                          heartbeat_timeout_ms=heartbeat_timeout_ms)
    clients = [LighthouseClient(lh.address(), T) for _ in range(n)]
    hb = Heartbeats(lh.address(), [f"r{i}" for i in range(n)])
    _, sizes = ask_all(clients, range(n), 0)
    assert sizes == [n], sizes
    hb.live.discard(f"r{n - 1}")  # replica n-1 dies now: no heartbeats, no requests
    ms, sizes = ask_all(clients, range(n - 1), 1)
    hb.stop()
    lh.shutdown()
    return {"replicas": n, "heartbeat_timeout_ms": heartbeat_timeout_ms, "join_timeout_ms": 100,
            "dead_member_to_new_quorum_ms": round(ms, 1), "new_quorum_sizes": sizes}


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=2000)
    ap.add_argument("--out", default="gpurun_out/control_plane_bench.json")
    a = ap.parse_args()
    res = {"host_cpus": os.cpu_count()}
    res.update(bench_manager_rpcs(a.iters))
    res["quorum_formation"] = [bench_quorum_formation(n, 5) for n in (2, 8, 32, 64, 128, 256, 512)]
    res["shrink_after_failure"] = [bench_shrink(n) for n in (2, 8)]
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    with open(a.out, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
