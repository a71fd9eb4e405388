"""End-to-end fault-injection tests on CPU (real Lighthouse + ManagerServer over loopback, real Gloo).

Harness modelled on the reference's manager_integ_test.py: an ``EventInjector`` schedules
failures (``fail_at``: the training loop raises; ``fail_allreduce_at``: the next collective's
future errors via ``FakeProcessGroupWrapper``), a ``Runner`` restarts a replica group up to
``attempts`` times (emulating torchelastic), and the oracle is that every replica finishes
with IDENTICAL state_dicts. Replica groups are threads; ranks within a group are threads too.
"""

import logging
import threading
import time
from concurrent.futures import ThreadPoolExecutor
from contextlib import ExitStack
from dataclasses import dataclass, field
from datetime import timedelta
from enum import Enum, auto
from typing import Any, Dict, List, Optional, Tuple

import pytest
import torch
from torch import nn, optim
from torch.distributed import TCPStore

from torchft_b200.checkpointing import PGTransport
from torchft_b200.coordination import LighthouseServer
from torchft_b200.ddp import DistributedDataParallel
from torchft_b200.manager import Manager
from torchft_b200.optim import OptimizerWrapper
from torchft_b200.process_group import FakeProcessGroupWrapper, ProcessGroupGloo

logger = logging.getLogger(__name__)


_INIT_LOCK = threading.Lock()


class InjectedFailure(Exception):
    pass


class EventType(Enum):
    FAILURE = auto()
    ALLREDUCE_FAILURE = auto()
    BARRIER = auto()


@dataclass
class Event:
    kind: EventType
    data: Any = None


class EventInjector:
    def __init__(self) -> None:
        self._lock = threading.Lock()
        self._events: Dict[Tuple[int, int], Event] = {}
        self.count: Dict[EventType, int] = {k: 0 for k in EventType}

    def fail_at(self, rank: int, step: int) -> "EventInjector":
        with self._lock:
            self._events[(rank, step)] = Event(EventType.FAILURE)
        return self

    def fail_allreduce_at(self, rank: int, step: int, times: int = 1) -> "EventInjector":
        """``times`` > 1: that many replica threads reaching (rank, step) each get the failure."""
        with self._lock:
            self._events[(rank, step)] = Event(EventType.ALLREDUCE_FAILURE, times)
        return self

    def barrier_at(self, rank: int, step: int, barrier: threading.Barrier) -> "EventInjector":
        with self._lock:
            self._events[(rank, step)] = Event(EventType.BARRIER, barrier)
        return self

    def check(self, rank: int, step: int, pg: Optional[FakeProcessGroupWrapper] = None) -> None:
        with self._lock:
            ev = self._events.get((rank, step))
            if ev is None:
                return
            if ev.kind == EventType.ALLREDUCE_FAILURE and isinstance(ev.data, int) and ev.data > 1:
                ev.data -= 1  # more replicas still have to see this one
            else:
                del self._events[(rank, step)]
            self.count[ev.kind] += 1
        if ev.kind == EventType.FAILURE:
            raise InjectedFailure(f"injected failure {rank=} {step=}")
        if ev.kind == EventType.ALLREDUCE_FAILURE:
            assert pg is not None
            pg.report_future_error(RuntimeError("injected allreduce failure"))
        if ev.kind == EventType.BARRIER:
            ev.data.wait()


class MyModel(nn.Module):
    def __init__(self, in_dim: int = 3, out_dim: int = 4) -> None:
        super().__init__()
        self.net = nn.Sequential(nn.Linear(in_dim, out_dim), nn.Sigmoid())
        self.in_dim = in_dim

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.net(x)

    def get_rand_inputs(self, bs: int) -> torch.Tensor:
        return torch.rand(bs, self.in_dim)


@dataclass
class Runner:
    replica_id: int
    lighthouse_address: str
    injector: EventInjector
    world_size: int = 1
    attempts: int = 3
    total_steps: int = 5
    use_async_quorum: bool = True
    init_sync: bool = True
    transport: str = "http"
    same_init: bool = False
    manager_kwargs: Dict[str, Any] = field(default_factory=dict)
    start_after: Optional[threading.Event] = None          # late joiner: wait for this before starting
    signal_at: Optional[Tuple[int, threading.Event]] = None  # (step, event): set the event once that step is reached

    def run(self) -> List[Dict[str, Any]]:
        if self.start_after is not None:
            assert self.start_after.wait(60), "late joiner was never released"
        for attempt in range(self.attempts):
            try:
                return self._run_group()
            except InjectedFailure as e:
                logger.info("replica %s attempt %s: %s", self.replica_id, attempt, e)
        raise RuntimeError("ran out of attempts")

    def _run_group(self) -> List[Dict[str, Any]]:
        store = TCPStore("127.0.0.1", 0, is_master=True, wait_for_workers=False)
        with ThreadPoolExecutor(max_workers=self.world_size, thread_name_prefix=f"rep{self.replica_id}") as ex:
            futs = [ex.submit(self._train, rank, store.port) for rank in range(self.world_size)]
            out = []
            err: Optional[BaseException] = None
            for f in futs:
                try:
                    out.append(f.result())
                except BaseException as e:  # noqa: BLE001
                    err = err or e
            if err is not None:
                raise err
            return out

    def _train(self, rank: int, store_port: int) -> Dict[str, Any]:
        with _INIT_LOCK:  # the global RNG is shared by all replica threads
            # different init per replica unless same_init: init_sync must make them equal
            torch.manual_seed(0 if self.same_init else 1000 * self.replica_id + rank)
            m = MyModel()
        opt_inner = optim.Adam(m.parameters(), lr=0.05)
        pg = FakeProcessGroupWrapper(ProcessGroupGloo(timeout=timedelta(seconds=10)))

        def load_state(sd: Dict[str, Any]) -> None:
            m.load_state_dict(sd["model"])
            opt_inner.load_state_dict(sd["optim"])

        def state() -> Dict[str, Any]:
            return {"model": m.state_dict(), "optim": opt_inner.state_dict()}

        transport = None
        if self.transport == "pg":
            transport = PGTransport(pg, timedelta(seconds=10), torch.device("cpu"))
        kwargs: Dict[str, Any] = dict(
            pg=pg, min_replica_size=2, load_state_dict=load_state, state_dict=state, replica_id=str(self.replica_id),
            store_addr="127.0.0.1", store_port=store_port, rank=rank, world_size=self.world_size,
            lighthouse_addr=self.lighthouse_address, use_async_quorum=self.use_async_quorum, init_sync=self.init_sync,
            timeout=timedelta(seconds=10), quorum_timeout=timedelta(seconds=20), checkpoint_transport=transport)
        kwargs.update(self.manager_kwargs)
        manager = Manager(**kwargs)
        stack = ExitStack()
        stack.callback(lambda: manager.shutdown(wait=False))
        stack.callback(pg.shutdown)
        with stack:
            ddp = DistributedDataParallel(manager, m)
            opt = OptimizerWrapper(manager, opt_inner)
            crit = nn.MSELoss()
            gen = torch.Generator().manual_seed(7)
            while manager.current_step() < self.total_steps:
                inputs = torch.rand(4, 3, generator=gen)
                labels = torch.rand(4, 4, generator=gen)
                opt.zero_grad()
                loss = crit(ddp(inputs), labels)
                loss.backward()
                self.injector.check(rank, manager.current_step(), pg)
                opt.step()
                if self.signal_at is not None and manager.current_step() >= self.signal_at[0]:
                    self.signal_at[1].set()
            return {"state": {k: v.clone() for k, v in m.state_dict().items()}, "step": manager.current_step(),
                    "batches": manager.batches_committed()}


def _run(runners: List[Runner]) -> List[List[Dict[str, Any]]]:
    with ThreadPoolExecutor(max_workers=len(runners)) as ex:
        futs = [ex.submit(r.run) for r in runners]
        return [f.result(timeout=120) for f in futs]


def _assert_equal_state(results: List[List[Dict[str, Any]]]) -> None:
    ref = results[0][0]["state"]
    for group in results:
        for r in group:
            for k, v in ref.items():
                torch.testing.assert_close(r["state"][k], v, msg=lambda m: f"{k}: {m}")


@pytest.fixture
def lighthouse():
    lh = LighthouseServer(bind="[::]:0", min_replicas=2, join_timeout_ms=200, heartbeat_timeout_ms=2000)
    yield lh
    lh.shutdown()


@pytest.mark.parametrize("use_async_quorum", [True, False])
def test_ddp_healthy(lighthouse, use_async_quorum):
    inj = EventInjector()
    res = _run([Runner(i, lighthouse.address(), inj, use_async_quorum=use_async_quorum) for i in range(2)])
    _assert_equal_state(res)
    assert all(r[0]["step"] == 5 for r in res)
    assert res[0][0]["batches"] == 10


@pytest.mark.parametrize("use_async_quorum", [True, False])
def test_ddp_recovery_after_replica_crash(lighthouse, use_async_quorum):
    """Replica 1 crashes at step 2, restarts from scratch, heals live from replica 0, and catches up."""
    inj = EventInjector().fail_at(0, 2)  # (group rank 0, step 2) -- first replica thread to reach it crashes
    res = _run([Runner(i, lighthouse.address(), inj, use_async_quorum=use_async_quorum) for i in range(2)])
    _assert_equal_state(res)
    assert inj.count[EventType.FAILURE] == 1


def test_ddp_skip_init_sync(lighthouse):
    """init_sync=False: replicas that start identical stay identical without a step-0 transfer."""
    inj = EventInjector()
    res = _run([Runner(i, lighthouse.address(), inj, init_sync=False, same_init=True) for i in range(2)])
    _assert_equal_state(res)


def test_ddp_allreduce_failure_is_discarded(lighthouse):
    """A failed collective on one replica discards that step on it; commit failure bumps the quorum id
    so both reconfigure, and replicas converge again."""
    inj = EventInjector().fail_allreduce_at(0, 1)
    res = _run([Runner(i, lighthouse.address(), inj, total_steps=4) for i in range(2)])
    assert inj.count[EventType.ALLREDUCE_FAILURE] == 1
    _assert_equal_state(res)


def test_ddp_commit_failure_on_every_replica(lighthouse):
    """The same step fails on BOTH replicas (reference: local_sgd_integ_test 'commit failure'): nobody commits it,
    nobody diverges, the quorum id is bumped and training resumes."""
    inj = EventInjector().fail_allreduce_at(0, 1, times=2)
    res = _run([Runner(i, lighthouse.address(), inj, total_steps=4) for i in range(2)])
    assert inj.count[EventType.ALLREDUCE_FAILURE] == 2
    _assert_equal_state(res)
    assert all(r[0]["step"] == 4 for r in res)
    # the failed step consumed a batch on each replica without being committed
    assert res[0][0]["batches"] == 8


def test_ddp_upscale_third_replica_joins_late(lighthouse):
    """Two replicas train; a third starts once they reached step 2 (reference: 'upscale'). It must be admitted,
    heal live from an up-to-date peer, and all three must finish with identical weights."""
    inj = EventInjector()
    go = threading.Event()
    runners = [Runner(0, lighthouse.address(), inj, total_steps=8, signal_at=(2, go)),
               Runner(1, lighthouse.address(), inj, total_steps=8),
               Runner(2, lighthouse.address(), inj, total_steps=8, start_after=go)]
    res = _run(runners)
    _assert_equal_state(res)
    assert all(r[0]["step"] == 8 for r in res)
    # the late joiner did not replay the steps it missed: it committed fewer batches than the founders' total
    assert res[2][0]["batches"] == res[0][0]["batches"]  # batches_committed is global state carried by the heal


def test_multi_rank_replica_groups(lighthouse):
    """world_size=2 inside each replica group: quorum and should_commit are group barriers (HSDP shape)."""
    inj = EventInjector()
    res = _run([Runner(i, lighthouse.address(), inj, world_size=2, total_steps=3) for i in range(2)])
    assert all(len(g) == 2 for g in res)
    # each group rank is its own "shard": rank r must match rank r of every other replica group
    for r in range(2):
        _assert_equal_state([[g[r]] for g in res])


def test_multi_rank_group_recovers_after_one_rank_crashes(lighthouse):
    """world_size=2 inside each group; rank 0 of one group crashes at step 2. Its sibling rank is stuck in the group barrier
    until its quorum deadline, the whole group restarts (torchelastic semantics), heals rank-for-rank from the healthy group
    and both groups finish with identical per-rank state (reference: manager_integ_test multi-rank recovery)."""
    inj = EventInjector().fail_at(0, 2)
    short = {"quorum_timeout": timedelta(seconds=4), "timeout": timedelta(seconds=4)}
    res = _run([Runner(i, lighthouse.address(), inj, world_size=2, total_steps=5, manager_kwargs=short, attempts=4) for i in range(2)])
    assert inj.count[EventType.FAILURE] == 1
    assert all(len(g) == 2 and all(r["step"] == 5 for r in g) for g in res)
    for r in range(2):
        _assert_equal_state([[g[r]] for g in res])


def test_recovery_with_pg_transport(lighthouse):
    inj = EventInjector().fail_at(0, 2)
    res = _run([Runner(i, lighthouse.address(), inj, transport="pg", use_async_quorum=False) for i in range(2)])
    _assert_equal_state(res)


def test_quorum_timeout_surfaces_quickly():
    lh = LighthouseServer(bind="[::]:0", min_replicas=2, join_timeout_ms=100)
    store = TCPStore("127.0.0.1", 0, is_master=True, wait_for_workers=False)
    pg = ProcessGroupGloo(timeout=timedelta(seconds=5))
    manager = Manager(pg=pg, min_replica_size=2, load_state_dict=lambda x: None, state_dict=lambda: {}, replica_id="solo",
                      store_addr="127.0.0.1", store_port=store.port, rank=0, world_size=1, lighthouse_addr=lh.address(),
                      timeout=timedelta(seconds=5))
    try:
        t0 = time.time()
        manager.start_quorum(timeout=timedelta(milliseconds=200))
        with pytest.raises(TimeoutError):
            manager.wait_quorum()
        assert time.time() - t0 < 1.5
    finally:
        manager.shutdown(wait=False)
        lh.shutdown()
