"""Manager state-machine unit tests (control plane and process group mocked).

Scenario coverage follows the reference's manager_test.py (happy path, sync/async
heal, zero contribution while healing, error latch + recovery, pg.errored,
FIXED_WITH_SPARES, allow_heal=False, wrap_future timeouts, numerics per op,
timeouts plumbed, max_retries, state-dict locking).
"""

import time
from datetime import timedelta
from typing import Optional
from unittest.mock import MagicMock, create_autospec, patch

import pytest
import torch
from torch.distributed import ReduceOp, TCPStore

from torchft_b200._C import QuorumResult
from torchft_b200.checkpointing.transport import CheckpointTransport
from torchft_b200.manager import MANAGER_ADDR_KEY, REPLICA_ID_KEY, Manager, WorldSizeMode, extract_trailing_digits
from torchft_b200.process_group import ProcessGroup
from torchft_b200.work import DummyWork


def make_quorum(**kw) -> QuorumResult:
    q = QuorumResult()
    q.quorum_id = 123
    q.replica_rank = 1
    q.replica_world_size = 2
    q.recover_src_manager_address = "manager address"
    q.store_address = "store_addr:1234"
    q.max_step = 1
    q.max_replica_rank = 1
    q.max_world_size = 2
    q.heal = False
    q.replica_ids = ["replica_0", "replica_1"]
    for k, v in kw.items():
        setattr(q, k, v)
    return q


class Harness:
    def __init__(self, client_cls: MagicMock, use_async_quorum=True, min_replica_size=2, world_size_mode=WorldSizeMode.DYNAMIC,
                 timeout=timedelta(seconds=10), init_sync=True, max_retries: Optional[int] = None) -> None:
        self.store = TCPStore("127.0.0.1", 0, is_master=True, wait_for_workers=False)
        self.store.set(MANAGER_ADDR_KEY, "dummy")
        self.store.set(REPLICA_ID_KEY, "dummy_id")
        self.pg = create_autospec(ProcessGroup)
        self.pg.errored.return_value = None
        self.pg.allreduce.side_effect = lambda tensors, opts: DummyWork(tensors)
        # autospec'd mocks have no fused path
        del self.pg.allreduce_native
        self.transport = create_autospec(CheckpointTransport)
        self.transport.metadata.return_value = "meta"
        self.load_state_dict = MagicMock()
        self.manager = Manager(
            pg=self.pg, min_replica_size=min_replica_size, load_state_dict=self.load_state_dict,
            state_dict=lambda: {"user": "state"}, replica_id="test_replica", store_addr="127.0.0.1",
            store_port=self.store.port, rank=1, world_size=2, use_async_quorum=use_async_quorum,
            world_size_mode=world_size_mode, timeout=timeout, checkpoint_transport=self.transport,
            init_sync=init_sync, max_retries=max_retries)
        self.client = client_cls.return_value
        self.client.should_commit.side_effect = lambda rank, step, vote, timeout: vote

    def close(self):
        self.manager.shutdown(wait=False)


@pytest.fixture
def client_cls():
    with patch("torchft_b200.manager.ManagerClient", autospec=True) as c:
        yield c


def test_extract_trailing_digits():
    assert extract_trailing_digits("replica_12") == 12
    assert extract_trailing_digits("abc") == 0
    assert extract_trailing_digits("7") == 7


def test_state_dict_roundtrip(client_cls):
    h = Harness(client_cls)
    try:
        assert h.manager.state_dict() == {"step": 0, "batches_committed": 0}
        h.manager.load_state_dict({"step": 1234, "batches_committed": 2345})
        assert h.manager.current_step() == 1234 and h.manager.batches_committed() == 2345
    finally:
        h.close()


def test_user_state_dict_registration(client_cls):
    h = Harness(client_cls)
    try:
        assert h.manager._manager_state_dict() == {"user": {"default": {"user": "state"}}, "torchft": {"step": 0, "batches_committed": 0}}
        h.manager.register_state_dict_fn("extra", MagicMock(), lambda: {"new": 1})
        assert h.manager._manager_state_dict()["user"]["extra"] == {"new": 1}
        with pytest.raises(AssertionError):
            h.manager.register_state_dict_fn("extra", MagicMock(), lambda: 0)
    finally:
        h.close()


def test_quorum_happy_path(client_cls):
    h = Harness(client_cls)
    m = h.manager
    try:
        h.client._quorum.return_value = make_quorum()
        assert m._quorum_id == -1 and m.current_step() == 0 and m.batches_committed() == 0
        m.start_quorum()
        m.allreduce(torch.tensor([1.0])).wait()
        assert m.should_commit()
        assert m._quorum_id == 123 and m.current_step() == 1 and m.batches_committed() == 2
        assert h.pg.configure.call_count == 1
        args = h.pg.configure.call_args[0]
        assert args[0] == "store_addr:1234/torchft/123/1" and args[2:5] == (1, 2, 123)
        assert h.pg.allreduce.call_count == 1
        # same quorum id next step: no reconfigure
        m.start_quorum()
        assert m.should_commit()
        assert h.pg.configure.call_count == 1 and m.current_step() == 2
    finally:
        h.close()


def test_quorum_heal_sync(client_cls):
    h = Harness(client_cls, use_async_quorum=False)
    m = h.manager
    try:
        h.client._quorum.return_value = make_quorum(recover_src_replica_rank=0, max_step=20, max_replica_rank=None, heal=True)
        h.client._checkpoint_metadata.return_value = "src meta"
        h.transport.recv_checkpoint.return_value = {"user": {"default": {"w": 1}}, "torchft": {"step": 20, "batches_committed": 0}}
        m.start_quorum()
        assert not m._healing  # healed eagerly
        assert m.current_step() == 20
        h.load_state_dict.assert_called_once_with({"w": 1})
        # sync quorum: a healed replica participates immediately with the full world
        assert m.is_participating() and m.num_participants() == 2
        h.transport.recv_checkpoint.assert_called_once()
        assert h.transport.recv_checkpoint.call_args.kwargs["metadata"] == "src meta"
        m.allreduce(torch.tensor([1.0])).wait()
        assert m.should_commit() and m.current_step() == 21
    finally:
        h.close()


def test_quorum_heal_async_zero_contribution(client_cls):
    h = Harness(client_cls, use_async_quorum=True, min_replica_size=1)
    m = h.manager
    try:
        h.client._quorum.return_value = make_quorum(recover_src_replica_rank=0, max_step=20, max_replica_rank=None,
                                                    max_world_size=1, heal=True)
        h.client._checkpoint_metadata.return_value = "src meta"
        h.transport.recv_checkpoint.return_value = {"user": {"default": {"w": 2}}, "torchft": {"step": 20, "batches_committed": 0}}
        m.start_quorum()
        m.wait_quorum()
        assert m._healing and not m.is_participating() and m.num_participants() == 1
        grad = torch.tensor([1.0, 2.0])
        m.allreduce(grad).wait()
        assert torch.equal(grad, torch.zeros(2))  # healer contributes zeros
        h.load_state_dict.assert_not_called()  # applied on the main thread at commit
        assert m.should_commit()
        h.load_state_dict.assert_called_once_with({"w": 2})
        assert m.current_step() == 21
        # next step: full member
        h.client._quorum.return_value = make_quorum(max_step=21, replica_rank=1, max_replica_rank=1)
        m.start_quorum()
        m.wait_quorum()
        assert m.is_participating() and not m._healing
    finally:
        h.close()


def test_send_checkpoint_to_recovering_peers(client_cls):
    h = Harness(client_cls)
    m = h.manager
    try:
        h.client._quorum.return_value = make_quorum(recover_dst_replica_ranks=[0, 3], max_step=7)
        m.start_quorum()
        m.wait_quorum()
        h.transport.send_checkpoint.assert_called_once()
        kw = h.transport.send_checkpoint.call_args.kwargs
        assert kw["dst_ranks"] == [0, 3] and kw["step"] == 7 and kw["state_dict"]["torchft"] == {"step": 0, "batches_committed": 0}
        assert m.should_commit()
        h.transport.disallow_checkpoint.assert_called()
    finally:
        h.close()


def test_allow_heal_false(client_cls):
    h = Harness(client_cls)
    m = h.manager
    try:
        h.client._quorum.return_value = make_quorum(heal=True, recover_src_replica_rank=0, max_replica_rank=None, max_step=20)
        m.start_quorum(allow_heal=False)
        m.wait_quorum()
        assert not m._healing
        h.transport.recv_checkpoint.assert_not_called()
        assert not m.is_participating()  # behind and not healed -> contributes zeros
    finally:
        h.close()


def test_not_enough_participants(client_cls):
    h = Harness(client_cls, min_replica_size=2)
    m = h.manager
    try:
        h.client._quorum.return_value = make_quorum(replica_world_size=1, max_world_size=1, max_replica_rank=0, replica_rank=0)
        m.start_quorum()
        assert m.num_participants() == 1
        assert not m.should_commit()
        assert m.current_step() == 0 and m._commit_failures == 1
        # commit_failures travels with the next quorum request
        m.start_quorum()
        m.wait_quorum()
        assert h.client._quorum.call_args.kwargs["commit_failures"] == 1
    finally:
        h.close()


def test_allreduce_error_latch_and_recovery(client_cls):
    h = Harness(client_cls)
    m = h.manager
    try:
        h.client._quorum.return_value = make_quorum()
        m.start_quorum()
        h.pg.allreduce.side_effect = RuntimeError("boom")
        t = torch.tensor([1.0])
        w = m.allreduce(t)
        assert isinstance(w, DummyWork) and m.errored() is not None
        h.pg.allreduce.reset_mock()
        assert isinstance(m.allreduce(t), DummyWork)  # latched: no further collectives
        h.pg.allreduce.assert_not_called()
        assert not m.should_commit()
        # next step with a bumped quorum id: error cleared, group reconfigured
        h.pg.allreduce.side_effect = lambda tensors, opts: DummyWork(tensors)
        h.client._quorum.return_value = make_quorum(quorum_id=124)
        m.start_quorum()
        assert m.errored() is None
        m.allreduce(t).wait()
        assert m.should_commit()
        assert h.pg.configure.call_count == 2
    finally:
        h.close()


def test_future_error_is_swallowed(client_cls):
    h = Harness(client_cls)
    m = h.manager
    try:
        h.client._quorum.return_value = make_quorum()
        m.start_quorum()

        class Failing(DummyWork):
            def get_future(self):
                f = torch.futures.Future()
                f.set_exception(RuntimeError("async failure"))
                return f

        h.pg.allreduce.side_effect = lambda tensors, opts: Failing(tensors)
        w = m.allreduce(torch.tensor([1.0]))
        assert w.wait() is True  # never raises
        assert m.errored() is not None and not m.should_commit()
    finally:
        h.close()


def test_pg_errored_blocks_commit(client_cls):
    h = Harness(client_cls)
    m = h.manager
    try:
        h.client._quorum.return_value = make_quorum()
        m.start_quorum()
        h.pg.errored.return_value = RuntimeError("aborted")
        assert not m.should_commit()
        assert "aborted" in str(m.errored())
    finally:
        h.close()


def test_configure_error_reported(client_cls):
    h = Harness(client_cls)
    m = h.manager
    try:
        h.client._quorum.return_value = make_quorum()
        h.pg.configure.side_effect = RuntimeError("configure failed")
        m.start_quorum()
        m.wait_quorum()
        assert m.errored() is not None and not m.should_commit()
    finally:
        h.close()


def test_checkpoint_error_reported(client_cls):
    h = Harness(client_cls)
    m = h.manager
    try:
        h.client._quorum.return_value = make_quorum(heal=True, recover_src_replica_rank=0, max_replica_rank=None, max_step=5)
        h.client._checkpoint_metadata.return_value = "m"
        h.transport.recv_checkpoint.side_effect = RuntimeError("fetch failed")
        m.start_quorum()
        m.wait_quorum()
        assert m.errored() is not None
        assert not m.should_commit()
    finally:
        h.close()


def test_fixed_with_spares(client_cls):
    h = Harness(client_cls, min_replica_size=2, world_size_mode=WorldSizeMode.FIXED_WITH_SPARES)
    m = h.manager
    try:
        h.client._quorum.return_value = make_quorum(replica_rank=2, replica_world_size=3, max_replica_rank=2, max_world_size=3)
        m.start_quorum()
        assert m.num_participants() == 2 and m.participating_rank() is None and not m.is_participating()
        g = torch.tensor([3.0])
        m.allreduce(g).wait()
        assert torch.equal(g, torch.zeros(1))
        h.client._quorum.return_value = make_quorum(replica_rank=1, replica_world_size=3, max_replica_rank=1, max_world_size=3)
        m.start_quorum()
        assert m.participating_rank() == 1 and m.is_participating()
    finally:
        h.close()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.int32])
@pytest.mark.parametrize("op", [ReduceOp.SUM, ReduceOp.AVG, ReduceOp.MAX])
def test_allreduce_numerics(client_cls, dtype, op):
    h = Harness(client_cls)
    m = h.manager
    try:
        h.client._quorum.return_value = make_quorum()
        m.start_quorum()
        t = torch.tensor([2, 4], dtype=dtype)
        if op == ReduceOp.AVG and not dtype.is_floating_point:
            with pytest.raises(ValueError, match="floating point"):
                m.allreduce(t, reduce_op=op)
            return
        m.allreduce(t, reduce_op=op).wait()
        # the mocked PG leaves data untouched: only AVG divides (by num_participants = 2)
        expect = torch.tensor([1, 2], dtype=dtype) if op == ReduceOp.AVG else torch.tensor([2, 4], dtype=dtype)
        assert torch.equal(t, expect)
        pg_op = h.pg.allreduce.call_args[0][1].reduceOp
        assert pg_op == (ReduceOp.SUM if op == ReduceOp.AVG else op)
    finally:
        h.close()


def test_wrap_future_default_and_timeout(client_cls):
    h = Harness(client_cls, timeout=timedelta(seconds=10))
    m = h.manager
    try:
        f = torch.futures.Future()
        w = m.wrap_future(f, 2)
        f.set_exception(RuntimeError("x"))
        assert w.wait() == 2 and m.errored() is not None
        m._errored = None
        f2 = torch.futures.Future()
        w2 = m.wrap_future(f2, "default", timeout=timedelta(milliseconds=50))
        assert w2.wait() == "default"
        assert "did not complete" in str(m.errored())
    finally:
        h.close()


def test_timeouts_plumbed(client_cls):
    h = Harness(client_cls)
    m = h.manager
    try:
        h.client._quorum.return_value = make_quorum()
        m.start_quorum(timeout=timedelta(seconds=12))
        m.wait_quorum()
        assert h.client._quorum.call_args.kwargs["timeout"] == timedelta(seconds=12)
        m.should_commit(timeout=timedelta(seconds=23))
        assert h.client.should_commit.call_args.kwargs["timeout"] == timedelta(seconds=23)
    finally:
        h.close()


def test_env_timeout_overrides(client_cls, monkeypatch):
    monkeypatch.setenv("TORCHFT_TIMEOUT_SEC", "7")
    monkeypatch.setenv("TORCHFT_QUORUM_TIMEOUT_SEC", "8")
    monkeypatch.setenv("TORCHFT_CONNECT_TIMEOUT_SEC", "9")
    monkeypatch.setenv("TORCHFT_QUORUM_RETRIES", "4")
    h = Harness(client_cls)
    try:
        m = h.manager
        assert (m._timeout, m._quorum_timeout, m._connect_timeout) == (timedelta(seconds=7), timedelta(seconds=8), timedelta(seconds=9))
        assert m._quorum_retries == 4
    finally:
        h.close()


def test_init_sync_flag_forwarded(client_cls):
    h = Harness(client_cls, init_sync=False)
    try:
        h.client._quorum.return_value = make_quorum()
        h.manager.start_quorum()
        h.manager.wait_quorum()
        assert h.client._quorum.call_args.kwargs["init_sync"] is False
    finally:
        h.close()


def test_max_retries(client_cls):
    h = Harness(client_cls, max_retries=2)
    m = h.manager
    try:
        h.client._quorum.return_value = make_quorum()
        h.client.should_commit.side_effect = lambda rank, step, vote, timeout: False
        for _ in range(2):
            m.start_quorum()
            assert not m.should_commit()
        m.start_quorum()
        with pytest.raises(RuntimeError, match="exceeding max_retries=2"):
            m.should_commit()
        # a success resets the counter
        h.client.should_commit.side_effect = lambda rank, step, vote, timeout: True
        m.start_quorum()
        assert m.should_commit() and m._commit_failures == 0
    finally:
        h.close()


def test_state_dict_read_lock(client_cls):
    h = Harness(client_cls, timeout=timedelta(milliseconds=200))
    m = h.manager
    try:
        m.disallow_state_dict_read()
        m.disallow_state_dict_read()  # idempotent
        with pytest.raises(TimeoutError):
            m._manager_state_dict()
        m.allow_state_dict_read()
        assert "user" in m._manager_state_dict()
    finally:
        h.close()


def test_commit_gate_tracks_verdict(client_cls):
    h = Harness(client_cls)
    m = h.manager
    try:
        gate = m.commit_gate()
        h.client._quorum.return_value = make_quorum()
        m.start_quorum()
        assert m.should_commit() and int(gate.item()) == 1
        h.client.should_commit.side_effect = lambda rank, step, vote, timeout: False
        m.start_quorum()
        assert not m.should_commit() and int(gate.item()) == 0
    finally:
        h.close()


def test_managed_work_lazy_callbacks(client_cls):
    from torchft_b200.manager import _ManagedWork

    h = Harness(client_cls)
    m = h.manager
    try:
        t = torch.tensor([1.0])
        calls = []
        w = _ManagedWork(m, DummyWork(t), t)
        f1 = w.get_future().then(lambda f: (calls.append("a"), f.value() * 2)[1])
        f2 = f1.then(lambda f: (calls.append("b"), f.value() + 1)[1])
        f3 = f2.then(lambda f: (calls.append("c"), str(f.value().item()))[1])  # type-changing callback
        assert calls == []  # lazy: nothing ran yet
        assert w.wait() is True
        assert calls == ["a", "b", "c"]
        assert f3.wait() == "3.0"
        with pytest.raises(NotImplementedError):
            f1.value()
    finally:
        h.close()


class _FailingWork(DummyWork):
    """Inner work whose wait() raises (a collective that failed after launch)."""

    def wait(self, timeout=None):  # type: ignore[override]
        raise RuntimeError("inner collective failed")


def test_managed_work_wait_never_raises_and_reports(client_cls):
    from torchft_b200.manager import _ManagedWork

    h = Harness(client_cls)
    m = h.manager
    try:
        t = torch.tensor([1.0])
        w = _ManagedWork(m, _FailingWork(t), t)
        assert m.errored() is None
        assert w.wait() is False          # swallowed ...
        assert m.errored() is not None    # ... but latched: this step will not commit
        assert "inner collective failed" in str(m.errored().original_exception)
    finally:
        h.close()


def test_managed_work_callback_exception_falls_back_to_default(client_cls):
    from torchft_b200.manager import _ManagedWork

    h = Harness(client_cls)
    m = h.manager
    try:
        t = torch.tensor([4.0])
        w = _ManagedWork(m, DummyWork(t), t)
        order = []

        def boom(f):
            order.append("boom")
            raise ValueError("callback exploded")

        fut = w.get_future().then(lambda f: (order.append("first"), f.value() + 1)[1]).then(boom).then(
            lambda f: order.append("never"))
        assert w.wait() is True           # the pipeline's failure is swallowed by wrap_future ...
        assert order == ["first", "boom"]  # ... later callbacks do not run
        assert fut.wait() is t            # ... and the future resolves to the default (the original tensor)
        assert "callback exploded" in str(m.errored().original_exception)
    finally:
        h.close()


def test_managed_work_other_entry_points_materialize_the_pipeline(client_cls):
    from torchft_b200.manager import _ManagedWork

    h = Harness(client_cls)
    m = h.manager
    try:
        for entry in ("block_current_stream", "synchronize"):
            t = torch.tensor([2.0])
            ran = []
            w = _ManagedWork(m, DummyWork(t), t)
            fut = w.get_future().then(lambda f: (ran.append(1), f.value() * 3)[1])
            getattr(w, entry)()
            assert ran == [1], entry      # callbacks ran without an explicit wait()
            assert fut.wait().item() == 6.0
            assert ran == [1]             # and only once
        # callbacks registered on different handles of the same work run in registration order
        t = torch.tensor([1.0])
        w = _ManagedWork(m, DummyWork(t), t)
        seq = []
        w.get_future().then(lambda f: seq.append("x") or f.value())
        w.get_future().then(lambda f: seq.append("y") or f.value())
        w.wait()
        assert seq == ["x", "y"]
    finally:
        h.close()
