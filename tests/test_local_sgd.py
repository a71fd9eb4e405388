"""LocalSGD / DiLoCo: unit tests against an autospec'd Manager (call counts, restore-on-failure,
bucketization, gradient sign, fragment schedule) and integration tests with real Lighthouse +
Gloo replicas (healthy, recovery, commit failure). Mirrors the reference's local_sgd_test.py /
local_sgd_integ_test.py scenarios; the oracle for integration is identical global state."""

import copy
from concurrent.futures import ThreadPoolExecutor
from datetime import timedelta
from typing import Any, Dict, List
from unittest.mock import MagicMock, create_autospec

import pytest
import torch
from torch import nn, optim
from torch.distributed import TCPStore

from torchft_b200.coordination import LighthouseServer
from torchft_b200.local_sgd import DiLoCo, LocalSGD
from torchft_b200.manager import Manager
from torchft_b200.process_group import FakeProcessGroupWrapper, ProcessGroupGloo
from torchft_b200.work import DummyWork


class SimpleModel(nn.Module):
    def __init__(self) -> None:
        super().__init__()
        self.model = nn.Sequential(nn.Linear(3, 4), nn.ReLU(), nn.Linear(4, 5), nn.Sigmoid())

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.model(x)


def _params(m: nn.Module) -> Dict[str, torch.Tensor]:
    return {n: p.detach().clone() for n, p in m.named_parameters()}


def _mock_manager(use_async_quorum=False) -> MagicMock:
    manager = create_autospec(Manager)
    manager._use_async_quorum = use_async_quorum
    manager.should_commit.return_value = True
    manager.current_step.return_value = 0
    manager.allreduce.side_effect = lambda t, **kw: DummyWork(t)
    manager.supports_fused_delta.return_value = False
    return manager


def test_local_sgd_healthy_call_counts():
    model = SimpleModel()
    opt = optim.SGD(model.parameters(), lr=0.1)
    manager = _mock_manager()
    with LocalSGD(manager, model, opt, sync_every=2) as lsgd:
        inp = torch.rand(2, 3)
        for i in range(2):
            opt.zero_grad()
            model(inp).mean().backward()
            opt.step()
            if i == 0:
                assert lsgd._local_step == 1 and manager.start_quorum.call_count == 0
        assert lsgd._local_step == 0
        assert manager.start_quorum.call_count == 1 and manager.should_commit.call_count == 1
        assert manager.allreduce.call_count >= 1  # flat buffer(s), not one per parameter
        assert manager.disallow_state_dict_read.call_count == 2 and manager.allow_state_dict_read.call_count == 2
    assert len(opt._optimizer_step_post_hooks) == 0  # hooks removed on exit


def test_local_sgd_failed_commit_keeps_local_weights():
    model = SimpleModel()
    opt = optim.SGD(model.parameters(), lr=0.1)
    manager = _mock_manager()
    manager.should_commit.return_value = False
    manager.allreduce.side_effect = lambda t, **kw: (t.zero_(), DummyWork(t))[1]  # "averaged" = zeros
    with LocalSGD(manager, model, opt, sync_every=1):
        opt.zero_grad()
        model(torch.rand(2, 3)).mean().backward()
        opt.step()
    assert any(p.abs().sum() > 0 for p in model.parameters())  # zeros were NOT applied


def test_diloco_requires_sync_quorum_and_validates_args():
    model = SimpleModel()
    inner, outer = optim.AdamW(model.parameters()), optim.SGD(model.parameters(), lr=0.7)
    with pytest.raises(ValueError, match="synchronous quorum"):
        DiLoCo(_mock_manager(use_async_quorum=True), [model], inner, outer, sync_every=2)
    m = _mock_manager()
    with pytest.raises(ValueError):
        DiLoCo(m, [model, model], inner, outer, sync_every=1)
    with pytest.raises(ValueError):
        DiLoCo(m, [model, model], inner, outer, sync_every=3)
    with pytest.raises(ValueError):
        DiLoCo(m, [model], inner, outer, sync_every=2, fragment_sync_delay=2)
    with pytest.raises(ValueError):
        DiLoCo(m, [model], inner, outer, sync_every=2, fragment_update_alpha=1.5)


def test_diloco_healthy_outer_step_and_backup():
    torch.manual_seed(0)
    model = SimpleModel()
    inner = optim.AdamW(model.parameters(), lr=1e-2)
    outer = optim.SGD(model.parameters(), lr=1.0)  # lr 1 + identity allreduce => global = local
    manager = _mock_manager()
    with DiLoCo(manager, [model], inner, outer, sync_every=2, backup_device=torch.device("cpu"), pin_memory=False) as d:
        frag = d._fragments[0]
        initial = _params(model)
        for n, t in frag.original_parameters.items():
            torch.testing.assert_close(t, initial[n])
        inp = torch.rand(2, 3)
        for _ in range(2):
            inner.zero_grad()
            model(inp).mean().backward()
            inner.step()
        assert manager.start_quorum.call_count == 1 and manager.should_commit.call_count == 1
        after = _params(model)
        # pseudo-gradient = original - local, SGD(lr=1): new = original - (original - local) = local
        # backup updated to the new global weights
        for n, t in frag.original_parameters.items():
            torch.testing.assert_close(t, after[n])
        assert any(not torch.equal(initial[n], after[n]) for n in initial)
    key = "StreamingDiLoCoFragment_0"
    assert manager.register_state_dict_fn.call_args[0][0] == key


def test_diloco_gradient_sign_and_failed_commit_restores():
    model = SimpleModel()
    inner = optim.SGD(model.parameters(), lr=0.5)
    outer = optim.SGD(model.parameters(), lr=1.0)
    manager = _mock_manager()
    captured: List[torch.Tensor] = []
    manager.allreduce.side_effect = lambda t, **kw: (captured.append(t.clone()), DummyWork(t))[1]
    manager.should_commit.return_value = False
    with DiLoCo(manager, [model], inner, outer, sync_every=1, backup_device=torch.device("cpu"), pin_memory=False):
        before = _params(model)
        inner.zero_grad()
        model(torch.rand(2, 3)).mean().backward()
        grads = {n: p.grad.clone() for n, p in model.named_parameters()}
        inner.step()
        after = _params(model)
    # failed commit: parameters reset to the last global copy
    for n in before:
        torch.testing.assert_close(after[n], before[n])
    # pseudo-gradient = original - local = lr * grad  (same sign as the gradient)
    flat = torch.cat([c.reshape(-1) for c in captured])
    expect = torch.cat([0.5 * grads[n].reshape(-1) for n, _ in model.named_parameters()])
    assert flat.numel() >= expect.numel()
    assert torch.allclose(flat[flat != 0].sort().values, expect[expect != 0].sort().values, atol=1e-6)


@pytest.mark.parametrize("use_bucketization,cap_mb", [(False, None), (True, 1), (True, 1e-4)])
def test_diloco_bucketization_results_identical(use_bucketization, cap_mb):
    torch.manual_seed(1)
    model = SimpleModel()
    inner = optim.SGD(model.parameters(), lr=0.1)
    outer = optim.SGD(model.parameters(), lr=0.7, momentum=0.9, nesterov=True)
    manager = _mock_manager()
    manager.allreduce.side_effect = lambda t, **kw: (t.mul_(2.0), DummyWork(t))[1]  # fake "sum of 2 replicas"
    with DiLoCo(manager, [model], inner, outer, sync_every=1, backup_device=torch.device("cpu"), pin_memory=False,
                use_bucketization=use_bucketization, bucket_cap_mb=cap_mb):
        torch.manual_seed(5)
        inner.zero_grad()
        model(torch.rand(2, 3)).mean().backward()
        inner.step()
    out = _params(model)
    # reference result computed without any bucketing
    torch.manual_seed(1)
    ref = SimpleModel()
    ref_inner = optim.SGD(ref.parameters(), lr=0.1)
    ref_outer = optim.SGD(ref.parameters(), lr=0.7, momentum=0.9, nesterov=True)
    orig = _params(ref)
    torch.manual_seed(5)
    ref_inner.zero_grad()
    ref(torch.rand(2, 3)).mean().backward()
    ref_inner.step()
    with torch.no_grad():
        for n, p in ref.named_parameters():
            g = (orig[n] - p.data) * 2.0
            p.data.copy_(orig[n])
            p.grad = g
    ref_outer.step()
    for n, p in ref.named_parameters():
        torch.testing.assert_close(out[n], p.detach())
    if use_bucketization and cap_mb == 1e-4:
        assert manager.allreduce.call_count > 1  # tiny cap => several buckets


def test_streaming_diloco_fragment_schedule():
    m1, m2 = nn.Linear(3, 3), nn.Linear(3, 3)
    model = nn.Sequential(m1, m2)
    inner = optim.SGD(model.parameters(), lr=0.1)
    outers = [optim.SGD(m1.parameters(), lr=1.0), optim.SGD(m2.parameters(), lr=1.0)]
    manager = _mock_manager()
    step = {"v": 0}
    manager.current_step.side_effect = lambda: step["v"]
    manager.should_commit.side_effect = lambda *a, **k: (step.__setitem__("v", step["v"] + 1), True)[1]
    synced: List[int] = []
    with DiLoCo(manager, [m1, m2], inner, outers, sync_every=4, backup_device=torch.device("cpu"), pin_memory=False,
                fragment_sync_delay=1) as d:
        for i, f in enumerate(d._fragments):
            orig = f.perform_sync
            f.perform_sync = (lambda o=orig, i=i: (synced.append(i), o())[1])  # type: ignore[method-assign]
        for _ in range(8):
            inner.zero_grad()
            model(torch.rand(2, 3)).mean().backward()
            inner.step()
    assert synced == [0, 1, 0, 1]  # one fragment every sync_every / n_fragments steps, round-robin
    assert manager.start_quorum.call_count == 4


# ------------------------------------------------------------------ integration
def _replica(lh_addr: str, rid: int, algo: str, steps: int, fail_allreduce_at: int, out: Dict[int, Any], use_quant=False,
             start_after=None, signal_at=None):
    if start_after is not None:  # late joiner (upscale scenario)
        assert start_after.wait(60)
    store = TCPStore("127.0.0.1", 0, is_master=True, wait_for_workers=False)
    torch.manual_seed(100 + rid)
    model = SimpleModel()
    pg = FakeProcessGroupWrapper(ProcessGroupGloo(timeout=timedelta(seconds=10)))
    inner = optim.SGD(model.parameters(), lr=0.05)
    outer = optim.SGD(model.parameters(), lr=0.7, momentum=0.9, nesterov=True)
    state: Dict[str, Any] = {}
    manager = Manager(pg=pg, min_replica_size=2, use_async_quorum=False,
                      load_state_dict=lambda sd: (model.load_state_dict(sd["model"]), inner.load_state_dict(sd["inner"])),
                      state_dict=lambda: {"model": model.state_dict(), "inner": inner.state_dict()},
                      replica_id=f"rep_{rid}", store_addr="127.0.0.1", store_port=store.port, rank=0, world_size=1,
                      lighthouse_addr=lh_addr, timeout=timedelta(seconds=10), quorum_timeout=timedelta(seconds=20))
    try:
        ctx = (LocalSGD(manager, model, inner, sync_every=2) if algo == "local_sgd" else
               DiLoCo(manager, [model], inner, outer, sync_every=2, backup_device=torch.device("cpu"), pin_memory=False))
        gen = torch.Generator().manual_seed(3)
        injected = False
        with ctx as algo_obj:
            while manager.current_step() < steps:
                if fail_allreduce_at >= 0 and manager.current_step() == fail_allreduce_at and not injected and rid == 0:
                    pg.report_future_error(RuntimeError("injected"))
                    injected = True
                inner.zero_grad()
                model(torch.rand(4, 3, generator=gen)).mean().backward()
                inner.step()
                if signal_at is not None and manager.current_step() >= signal_at[0]:
                    signal_at[1].set()
            res = {"params": _params(model), "step": manager.current_step()}
            if algo == "diloco":
                res["original"] = {n: t.clone() for n, t in algo_obj._fragments[0].original_parameters.items()}
                res["outer"] = copy.deepcopy(outer.state_dict())
            out[rid] = res
    finally:
        manager.shutdown(wait=False)
        pg.shutdown()


@pytest.mark.parametrize("algo", ["local_sgd", "diloco"])
@pytest.mark.parametrize("fail_at", [-1, 1])
def test_integration_two_replicas(algo, fail_at):
    lh = LighthouseServer(bind="[::]:0", min_replicas=2, join_timeout_ms=200)
    out: Dict[int, Any] = {}
    try:
        with ThreadPoolExecutor(max_workers=2) as ex:
            futs = [ex.submit(_replica, lh.address(), r, algo, 3, fail_at, out) for r in range(2)]
            for f in futs:
                f.result(timeout=120)
    finally:
        lh.shutdown()
    assert out[0]["step"] == out[1]["step"] == 3
    key = "original" if algo == "diloco" else "params"
    for n in out[0][key]:
        torch.testing.assert_close(out[0][key][n], out[1][key][n])
    if algo == "diloco":
        s0, s1 = out[0]["outer"]["state"], out[1]["outer"]["state"]
        for k in s0:
            torch.testing.assert_close(s0[k]["momentum_buffer"], s1[k]["momentum_buffer"])


@pytest.mark.parametrize("algo", ["local_sgd", "diloco"])
def test_integration_upscale_third_replica_joins(algo):
    """Two replicas sync for a while, a third joins at outer step 2 (reference: local_sgd_integ_test 'upscale'). The healing
    payload must carry the algorithm's own state too (DiLoCo: backup weights + outer optimizer, registered with the Manager),
    so that all three end with identical global state."""
    import threading

    lh = LighthouseServer(bind="[::]:0", min_replicas=2, join_timeout_ms=300)
    out: Dict[int, Any] = {}
    go = threading.Event()
    steps = 6
    try:
        with ThreadPoolExecutor(max_workers=3) as ex:
            futs = [ex.submit(_replica, lh.address(), 0, algo, steps, -1, out, False, None, (2, go)),
                    ex.submit(_replica, lh.address(), 1, algo, steps, -1, out),
                    ex.submit(_replica, lh.address(), 2, algo, steps, -1, out, False, go, None)]
            for f in futs:
                f.result(timeout=180)
    finally:
        lh.shutdown()
    assert out[0]["step"] == out[1]["step"] == out[2]["step"] == steps
    key = "original" if algo == "diloco" else "params"
    for r in (1, 2):
        for n in out[0][key]:
            torch.testing.assert_close(out[0][key][n], out[r][key][n])
    if algo == "diloco":
        for r in (1, 2):
            s0, sr = out[0]["outer"]["state"], out[r]["outer"]["state"]
            for k in s0:
                torch.testing.assert_close(s0[k]["momentum_buffer"], sr[k]["momentum_buffer"])


class _Crash(Exception):
    pass


def _streaming_replica(lh_addr: str, rid: int, steps: int, crash_at: int, out: Dict[int, Any], attempts: Dict[int, int]):
    """Streaming DiLoCo (2 fragments, staggered syncs) replica that can crash once and restart from scratch."""
    for attempt in range(2):
        attempts[rid] = attempt + 1
        store = TCPStore("127.0.0.1", 0, is_master=True, wait_for_workers=False)
        torch.manual_seed(200 + rid)
        f1, f2 = nn.Linear(3, 4), nn.Linear(4, 2)
        model = nn.Sequential(f1, nn.ReLU(), f2)
        pg = ProcessGroupGloo(timeout=timedelta(seconds=10))
        inner = optim.SGD(model.parameters(), lr=0.05)
        outers = [optim.SGD(f1.parameters(), lr=0.7, momentum=0.9, nesterov=True),
                  optim.SGD(f2.parameters(), lr=0.7, momentum=0.9, nesterov=True)]
        manager = Manager(pg=pg, min_replica_size=2, use_async_quorum=False,
                          load_state_dict=lambda sd: (model.load_state_dict(sd["model"]), inner.load_state_dict(sd["inner"])),
                          state_dict=lambda: {"model": model.state_dict(), "inner": inner.state_dict()},
                          replica_id=f"srep_{rid}", store_addr="127.0.0.1", store_port=store.port, rank=0, world_size=1,
                          lighthouse_addr=lh_addr, timeout=timedelta(seconds=10), quorum_timeout=timedelta(seconds=30))
        try:
            gen = torch.Generator().manual_seed(5)
            with DiLoCo(manager, [f1, f2], inner, outers, sync_every=4, backup_device=torch.device("cpu"), pin_memory=False,
                        fragment_sync_delay=1) as d:
                while manager.current_step() < steps:
                    if attempt == 0 and crash_at >= 0 and manager.current_step() == crash_at:
                        raise _Crash()
                    inner.zero_grad()
                    model(torch.rand(4, 3, generator=gen)).mean().backward()
                    inner.step()
                out[rid] = {"step": manager.current_step(),
                            "original": [{n: t.clone() for n, t in f.original_parameters.items()} for f in d._fragments],
                            "outer": [copy.deepcopy(o.state_dict()) for o in outers]}
            return
        except _Crash:
            continue
        finally:
            manager.shutdown(wait=False)
            pg.shutdown()


def test_integration_streaming_diloco_recovery_after_crash():
    """Replica 1 dies mid-run (between fragment syncs) and restarts with fresh weights: it must heal BOTH fragments'
    backup weights and outer optimizers from replica 0 and finish with identical global state
    (reference: local_sgd_integ_test 'streaming recovery')."""
    lh = LighthouseServer(bind="[::]:0", min_replicas=2, join_timeout_ms=300, heartbeat_timeout_ms=1000)
    out: Dict[int, Any] = {}
    attempts: Dict[int, int] = {}
    steps = 6
    try:
        with ThreadPoolExecutor(max_workers=2) as ex:
            futs = [ex.submit(_streaming_replica, lh.address(), 0, steps, -1, out, attempts),
                    ex.submit(_streaming_replica, lh.address(), 1, steps, 3, out, attempts)]
            for f in futs:
                f.result(timeout=180)
    finally:
        lh.shutdown()
    assert attempts == {0: 1, 1: 2}
    assert out[0]["step"] == out[1]["step"] == steps
    for frag in range(2):
        for n in out[0]["original"][frag]:
            torch.testing.assert_close(out[0]["original"][frag][n], out[1]["original"][frag][n])
        s0, s1 = out[0]["outer"][frag]["state"], out[1]["outer"][frag]["state"]
        assert set(s0) == set(s1) and len(s0) > 0
        for k in s0:
            torch.testing.assert_close(s0[k]["momentum_buffer"], s1[k]["momentum_buffer"])
