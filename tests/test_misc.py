"""Small-component tests: futures/timeouts, sampler, optimizer wrapper, DDP wrappers, parameter
server, structured event sink, coordination API docs, monitored pipe, quantization helpers."""

import json
import logging
import multiprocessing as mp
import threading
import time
from datetime import timedelta
from unittest.mock import MagicMock, create_autospec, patch

import pytest
import torch
from torch import nn
from torch.futures import Future

from torchft_b200 import coordination
from torchft_b200.data import DistributedSampler
from torchft_b200.ddp import DistributedDataParallel, PureDistributedDataParallel
from torchft_b200.futures import context_timeout, future_timeout, future_wait
from torchft_b200.manager import Manager
from torchft_b200.multiprocessing import _MonitoredPipe
from torchft_b200.optim import OptimizerWrapper
from torchft_b200.parameter_server import ParameterServer
from torchft_b200.process_group import ProcessGroup, ProcessGroupGloo
from torchft_b200.work import DummyWork


# ----------------------------------------------------------------------- futures
def test_future_timeout_fires_and_passes_through():
    f = Future()
    t = future_timeout(f, timedelta(milliseconds=50))
    with pytest.raises(TimeoutError):
        t.wait()
    f2 = Future()
    t2 = future_timeout(f2, timedelta(seconds=10))
    f2.set_result(7)
    assert t2.wait() == 7
    f3 = Future()
    t3 = future_timeout(f3, timedelta(seconds=10))
    f3.set_exception(RuntimeError("boom"))
    with pytest.raises(RuntimeError, match="boom"):
        t3.wait()


def test_future_wait():
    f = Future()
    with pytest.raises(TimeoutError):
        future_wait(f, timedelta(milliseconds=30))
    f.set_result("ok")
    assert future_wait(f, timedelta(seconds=1)) == "ok"


def test_context_timeout_calls_back_only_when_late():
    cb = MagicMock()
    with context_timeout(cb, timedelta(seconds=5)):
        pass
    time.sleep(0.05)
    cb.assert_not_called()
    with context_timeout(cb, timedelta(milliseconds=20)):
        time.sleep(0.2)
    cb.assert_called_once()


def test_timeout_watchdog_exits_when_timer_thread_is_stuck(monkeypatch):
    from torchft_b200 import futures as F

    monkeypatch.setenv(F.WATCHDOG_TIMEOUT_SEC_ENV, "0.1")
    mgr = F._TimeoutManager()
    exited = threading.Event()
    with patch("sys.exit", side_effect=lambda code: exited.set()):
        mgr.call_later(timedelta(milliseconds=1), lambda: time.sleep(1.0))  # wedge the timer thread
        assert exited.wait(3.0)
    mgr.shutdown()


# ------------------------------------------------------------------------- data
def test_distributed_sampler_grid():
    ds = list(range(1000))
    s = DistributedSampler(ds, replica_rank=1, num_replica_groups=2, group_rank=3, num_replicas=4, shuffle=False)
    assert s.global_rank == 3 + 4 * 1 and s.global_world_size == 8
    idx = list(iter(s))
    assert idx[0] == 7 and len(idx) == 125 and idx[1] - idx[0] == 8


# ------------------------------------------------------------------------ optim
def test_optimizer_wrapper_gates_step_on_commit():
    manager = create_autospec(Manager)
    m = nn.Linear(3, 4)
    inner = torch.optim.SGD(m.parameters(), lr=1.0)
    opt = OptimizerWrapper(manager, inner)
    opt.add_param_group({"params": [nn.Parameter(torch.zeros(1))], "lr": 1})
    assert len(opt.param_groups) == 2 and opt.state == inner.state
    opt.zero_grad()
    manager.start_quorum.assert_called_once()
    m(torch.ones(2, 3)).sum().backward()
    before = m.weight.detach().clone()
    manager.should_commit.return_value = False
    opt.step()
    assert torch.equal(m.weight, before)
    manager.should_commit.return_value = True
    opt.step()
    assert not torch.equal(m.weight, before)
    sd = opt.state_dict()
    opt.load_state_dict(sd)


# -------------------------------------------------------------------------- ddp
def test_optimizer_wrapper_works_with_lr_schedulers_and_forwards_state():
    from unittest.mock import MagicMock

    from torchft_b200.optim import OptimizerWrapper

    manager = MagicMock()
    manager.should_commit.return_value = True
    p = torch.nn.Parameter(torch.ones(3))
    inner = torch.optim.SGD([p], lr=0.1, momentum=0.9)
    opt = OptimizerWrapper(manager, inner)
    assert isinstance(opt, torch.optim.Optimizer)
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=1, gamma=0.5)  # patches opt.step on the facade only
    p.grad = torch.ones(3)
    opt.step()
    sched.step()
    assert opt.last_step_committed and torch.allclose(p.data, torch.full((3,), 0.9))
    assert opt.param_groups is inner.param_groups and inner.param_groups[0]["lr"] == pytest.approx(0.05)
    sd = opt.state_dict()
    assert set(sd) == {"state", "param_groups"} and len(sd["state"]) == 1  # momentum buffer of the inner optimizer
    opt.load_state_dict(sd)
    manager.should_commit.return_value = False
    p.grad = torch.ones(3)
    before = p.data.clone()
    opt.step()
    assert not opt.last_step_committed and torch.equal(p.data, before)
    with pytest.raises(NotImplementedError):
        opt.step(closure=lambda: 0.0)


def test_ddp_wrapper_routes_buckets_through_manager():
    manager = create_autospec(Manager)
    manager.allreduce.side_effect = lambda t, **kw: DummyWork(t)
    m = nn.Linear(3, 4)
    ddp = DistributedDataParallel(manager, m)
    ddp(torch.ones(2, 3)).sum().backward()
    assert manager.allreduce.call_count >= 1
    assert all(p.grad is not None for p in m.parameters())


def test_pure_ddp_one_allreduce_per_param():
    manager = create_autospec(Manager)
    manager.allreduce.side_effect = lambda t, **kw: DummyWork(t)
    m = nn.Linear(3, 4)
    ddp = PureDistributedDataParallel(manager, m)
    ddp(torch.ones(2, 3)).sum().backward()
    assert manager.allreduce.call_count == len(list(m.parameters()))


# ------------------------------------------------------------- parameter server
class _PS(ParameterServer):
    @classmethod
    def new_process_group(cls) -> ProcessGroup:
        return ProcessGroupGloo(timeout=timedelta(seconds=10))

    def forward(self, session_id: str, pg: ProcessGroup) -> None:
        t = torch.zeros(3)
        pg.recv([t], 1, 11).wait()
        pg.send([t * 2], 1, 12).wait()


def test_parameter_server_session_roundtrip():
    ps = _PS(port=0)
    try:
        pg = _PS.new_session(ps.address().replace(ps.address().split("//")[1].split(":")[0], "127.0.0.1"))
        x = torch.tensor([1.0, 2.0, 3.0])
        pg.send([x], 0, 11).wait()
        out = torch.zeros(3)
        pg.recv([out], 0, 12).wait()
        assert torch.equal(out, x * 2)
        pg.shutdown()
    finally:
        ps.shutdown()


# ------------------------------------------------------------------ event sink
def test_jsonl_event_sink(tmp_path, monkeypatch):
    from torchft_b200 import otel

    path = tmp_path / "events.jsonl"
    monkeypatch.setenv(otel.EVENTS_JSONL_ENV, str(path))
    otel.setup_logger("torchft_test_events")
    logging.getLogger("torchft_test_events").info("", extra={"job_id": "j", "replica_id": "r0", "quorum_id": 3, "step": 9})
    otel.shutdown()
    rec = json.loads(path.read_text().strip().splitlines()[-1])
    assert rec["logger"] == "torchft_test_events" and rec["quorum_id"] == 3 and rec["step"] == 9 and rec["replica_id"] == "r0"


def test_prometheus_sink_counts_quorums_commits_errors(monkeypatch):
    import socket
    import urllib.request

    from torchft_b200 import otel

    pytest.importorskip("prometheus_client")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    monkeypatch.setenv(otel.PROMETHEUS_PORT_ENV, str(port))
    for name in ("torchft_quorums", "torchft_commits", "torchft_errors"):
        otel.setup_logger(name)
    base = {"job_id": "j", "replica_id": "r7", "rank": 0}
    logging.getLogger("torchft_quorums").info("", extra=dict(base, quorum_id=5, step=11))
    for ok in (True, True, False):
        logging.getLogger("torchft_commits").info("", extra=dict(base, quorum_id=5, step=12, commit_result=ok))
    logging.getLogger("torchft_errors").info("", extra=dict(base, quorum_id=5, error="boom"))
    text = urllib.request.urlopen(f"http://127.0.0.1:{port}/metrics", timeout=5).read().decode()
    otel.shutdown()
    assert 'torchft_quorum_changes_total{rank="0",replica_id="r7"} 1.0' in text
    assert 'torchft_commits_total{rank="0",replica_id="r7",result="committed"} 2.0' in text
    assert 'torchft_commits_total{rank="0",replica_id="r7",result="failed"} 1.0' in text
    assert 'torchft_errors_total{rank="0",replica_id="r7"} 1.0' in text
    assert 'torchft_step{rank="0",replica_id="r7"} 12.0' in text and 'torchft_quorum_id{rank="0",replica_id="r7"} 5.0' in text


def test_otel_requested_but_missing_raises(monkeypatch):
    from torchft_b200 import otel

    monkeypatch.setenv(otel.TORCHFT_USE_OTEL, "true")
    try:
        import opentelemetry  # noqa: F401

        pytest.skip("opentelemetry installed")
    except ImportError:
        with pytest.raises(RuntimeError, match="opentelemetry"):
            otel.setup_logger("torchft_test_otel")


# ----------------------------------------------------------------- coordination
def test_coordination_api_is_documented():
    for name in coordination.__all__:
        obj = getattr(coordination, name)
        assert obj is not None and (obj.__doc__ or "").strip(), f"{name} is undocumented"
    assert "quorum" in (coordination.__doc__ or "").lower()
    q = coordination.QuorumResult()
    assert q.quorum_id == 0 and q.recover_src_replica_rank is None and q.replica_ids == []


# ------------------------------------------------------------------------ pipes
def test_monitored_pipe_timeout_and_exception_transport():
    a, b = mp.Pipe()
    pa, pb = _MonitoredPipe(a), _MonitoredPipe(b)
    with pytest.raises(TimeoutError):
        pa.recv(0.05)
    pb.send("hello")
    assert pa.recv(timedelta(seconds=1)) == "hello"
    pb.send(ValueError("remote"))
    with pytest.raises(ValueError, match="remote"):
        pa.recv(1.0)
    # failures sent as an envelope keep the remote traceback as __cause__
    from torchft_b200.multiprocessing import failure_of

    try:
        raise KeyError("deep in the child")
    except KeyError as e:
        pb.send(failure_of(e))
    with pytest.raises(KeyError) as ei:
        pa.recv(1.0)
    assert "deep in the child" in str(ei.value.__cause__) and "Traceback" in str(ei.value.__cause__)
    # a dead peer is noticed long before the deadline
    import time

    watched = _MonitoredPipe(a, alive=lambda: False)
    t0 = time.monotonic()
    with pytest.raises(RuntimeError, match="exited"):
        watched.recv(30.0)
    assert time.monotonic() - t0 < 2.0


def test_get_padded_sizes_and_reduce_scatter_output():
    from torchft_b200.collectives import allocate_reduce_scatter_output, get_padded_sizes

    ts = [torch.zeros(5, 3), torch.zeros(4), torch.zeros(8, 2)]
    assert get_padded_sizes(ts, 4) == [torch.Size([8, 3]), torch.Size([4]), torch.Size([8, 2])]
    out, padded = allocate_reduce_scatter_output(ts, 4)
    assert out.numel() == 6 + 1 + 4 and padded[0] == torch.Size([8, 3])


def test_reference_import_paths_resolve():
    """Names the reference exposes from torchft.process_group / torchft import the same way here."""
    import torchft_b200
    from torchft_b200.process_group import (  # noqa: F401
        ErrorSwallowingProcessGroupWrapper, FakeProcessGroupWrapper, ManagedProcessGroup, ProcessGroup, ProcessGroupBabyGloo,
        ProcessGroupBabyNCCL, ProcessGroupDummy, ProcessGroupGloo, ProcessGroupNCCL, ProcessGroupWrapper)

    for name in ("Manager", "Optimizer", "DistributedDataParallel", "DistributedSampler", "ProcessGroupGloo", "ProcessGroupNCCL",
                 "ProcessGroupBabyGloo", "ProcessGroupBabyNCCL", "ManagedProcessGroup", "WorldSizeMode"):
        assert hasattr(torchft_b200, name), name
    assert ProcessGroupBabyGloo is torchft_b200.ProcessGroupBabyGloo


def test_doctor_reports_and_exit_code():
    import subprocess
    import sys

    from torchft_b200 import doctor
    from torchft_b200.coordination import LighthouseServer

    lh = LighthouseServer(bind="127.0.0.1:0", min_replicas=3)
    try:
        checks = doctor.run(lighthouse=lh.address())
    finally:
        lh.shutdown()
    by_name = {n: (lvl, d) for lvl, n, d in checks}
    assert by_name["extension _C"][0] == "OK" and by_name["extension _K"][0] == "OK"
    assert by_name["lighthouse"][0] == "OK" and "min_replicas=3" in by_name["lighthouse"][1]
    if not torch.cuda.is_available():
        assert by_name["cuda"][0] == "WARN"
        r = subprocess.run([sys.executable, "-m", "torchft_b200.doctor", "--require-gpu"], capture_output=True, text=True, timeout=120)
        assert r.returncode == 1 and "[FAIL] cuda" in r.stdout
    # an unreachable lighthouse is a FAIL line, not a crash
    assert dict((n, lvl) for lvl, n, _ in doctor.run(lighthouse="http://127.0.0.1:9"))["lighthouse"] == "FAIL"


def test_future_timeout_releases_the_result_once_the_future_completed():
    """A completed (hence cancelled) timer must not keep the guarded future's value alive until its deadline: with a 60 s
    op timeout every collective of every step would otherwise pin its tensor for a minute (found as +1 full gradient of
    device memory per step under FSDP2)."""
    import gc
    import weakref
    from datetime import timedelta

    from torch.futures import Future

    from torchft_b200.futures import future_timeout

    class Payload:
        pass

    fut: Future = Future()
    guarded = future_timeout(fut, timedelta(seconds=300))
    p = Payload()
    ref = weakref.ref(p)
    fut.set_result(p)
    assert guarded.wait() is p
    del p, fut, guarded
    gc.collect()
    assert ref() is None, "timer heap still references the completed future's value"
