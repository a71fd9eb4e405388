"""Process-group tests on CPU: world-size-1 API sweep, multi-rank numerics over real Gloo
(rank threads sharing one store), resiliency (a dying rank must surface as an error on the
survivors within the timeout), wrappers, subprocess ("Baby") groups, c10d registration."""

import threading
import time
from concurrent.futures import ThreadPoolExecutor
from datetime import timedelta
from typing import Any, Callable, List
from unittest.mock import Mock

import pytest
import torch
import torch.distributed as dist
from torch.distributed import ReduceOp, TCPStore
from torch.distributed.distributed_c10d import (
    AllgatherOptions,
    AllreduceOptions,
    AllToAllOptions,
    BarrierOptions,
    BroadcastOptions,
    ReduceScatterOptions,
)

from torchft_b200.baby import ProcessGroupBabyGloo
from torchft_b200.manager import Manager
from torchft_b200.process_group import ErrorSwallowingProcessGroupWrapper, FakeProcessGroupWrapper, ManagedProcessGroup, ProcessGroup, ProcessGroupDummy, ProcessGroupGloo, create_store_client


def _store():
    return TCPStore("127.0.0.1", 0, is_master=True, wait_for_workers=False)


def run_all_collectives(pg: ProcessGroup, device="cpu", skip=()) -> List[str]:
    """Every collective on a world-size-1 group: shapes preserved, works complete."""
    t = torch.arange(6, dtype=torch.float32, device=device).reshape(2, 3)
    ran = []

    def check(name, work, *tensors):
        work.wait()
        fut = work.get_future()
        fut.wait()
        ran.append(name)

    ar = AllreduceOptions()
    ar.reduceOp = ReduceOp.SUM
    if "allreduce" not in skip:
        x = t.clone()
        check("allreduce", pg.allreduce([x], ar))
        assert torch.equal(x, t)
        from torch.distributed.distributed_c10d import AllreduceCoalescedOptions

        check("allreduce_coalesced", pg.allreduce_coalesced([t.clone(), t.clone()], AllreduceCoalescedOptions()))
    if "allgather" not in skip:
        out = [[torch.zeros_like(t)]]
        check("allgather", pg.allgather(out, [t], AllgatherOptions()))
        assert torch.equal(out[0][0], t)
        o = torch.zeros_like(t)
        check("allgather_into_tensor_coalesced", pg.allgather_into_tensor_coalesced([o], [t], AllgatherOptions()))
        assert torch.equal(o, t)
    if "broadcast" not in skip:
        b = BroadcastOptions()
        b.rootRank = 0
        check("broadcast", pg.broadcast([t.clone()], b))
        check("broadcast_one", pg.broadcast_one(t.clone(), 0))
    if "barrier" not in skip:
        check("barrier", pg.barrier(BarrierOptions()))
    if "alltoall_base" not in skip:
        o = torch.zeros_like(t)
        check("alltoall_base", pg.alltoall_base(o, t, [], [], AllToAllOptions()))
        assert torch.equal(o, t)
    if "reduce_scatter" not in skip:
        o = torch.zeros_like(t)
        check("reduce_scatter", pg.reduce_scatter([o], [[t]], ReduceScatterOptions()))
        check("reduce_scatter_tensor_coalesced", pg.reduce_scatter_tensor_coalesced([o], [t], ReduceScatterOptions()))
    return ran


def test_dummy_pg():
    pg = ProcessGroupDummy(0, 1)
    ran = run_all_collectives(pg)
    assert "allreduce" in ran and "reduce_scatter" in ran
    assert pg.size() == 1 and pg.wait_count > 0 and pg.get_future_count > 0
    pg.configure("addr", "rid", 0, 1)
    assert pg.configure_count == 1
    pg.send([torch.ones(1)], 0, 1).wait()
    pg.recv([torch.ones(1)], 0, 1).wait()


def test_gloo_world1_apis():
    store = _store()
    pg = ProcessGroupGloo(timeout=timedelta(seconds=10))
    pg.configure(f"127.0.0.1:{store.port}/a/0", "r0", 0, 1, quorum_id=3, group_rank=0, group_world_size=1, global_ranks=[0])
    assert pg.size() == 1 and pg.getBackendName() == "torchft-gloo"
    ran = run_all_collectives(pg, skip=("reduce_scatter",))
    assert {"allreduce", "allgather", "broadcast", "barrier", "alltoall_base"} <= set(ran)
    with pytest.raises(RuntimeError, match="does not support reduce_scatter"):
        pg.reduce_scatter([], [], ReduceScatterOptions())
    # reconfigure replaces the inner group
    inner = pg.parent
    pg.configure(f"127.0.0.1:{store.port}/a/1", "r0", 0, 1)
    assert pg.parent is not inner
    pg.shutdown()
    with pytest.raises(RuntimeError, match="not initialized"):
        pg.parent


def test_create_store_client_prefix_isolation():
    store = _store()
    a = create_store_client(f"127.0.0.1:{store.port}/p/1", timedelta(seconds=5))
    b = create_store_client(f"127.0.0.1:{store.port}/p/2", timedelta(seconds=5))
    a.set("k", "1")
    b.set("k", "2")
    assert a.get("k") == b"1" and b.get("k") == b"2"


def _ranks(world: int, fn: Callable[[int, ProcessGroup], Any], make=lambda: ProcessGroupGloo(timeout=timedelta(seconds=10)),
           prefix="m") -> List[Any]:
    store = _store()
    pgs = [make() for _ in range(world)]

    def one(rank):
        pg = pgs[rank]
        pg.configure(f"127.0.0.1:{store.port}/{prefix}/0", f"rep{rank}", rank, world)
        return fn(rank, pg)

    with ThreadPoolExecutor(max_workers=world) as ex:
        out = list(ex.map(one, range(world)))
    for pg in pgs:
        pg.shutdown()
    return out


def test_gloo_multirank_numerics():
    W = 3

    def body(rank, pg):
        res = {}
        t = torch.full((4,), float(rank + 1))
        o = AllreduceOptions()
        o.reduceOp = ReduceOp.SUM
        pg.allreduce([t], o).wait()
        res["allreduce"] = t.clone()
        outs = [[torch.zeros(2) for _ in range(W)]]
        pg.allgather(outs, [torch.full((2,), float(rank))], AllgatherOptions()).wait()
        res["allgather"] = torch.stack(outs[0])
        b = torch.full((3,), float(rank))
        bo = BroadcastOptions()
        bo.rootRank = 1
        pg.broadcast([b], bo).wait()
        res["broadcast"] = b
        inp = torch.arange(W, dtype=torch.float32) + 10 * rank
        out = torch.zeros(W)
        pg.alltoall_base(out, inp, [], [], AllToAllOptions()).wait()
        res["alltoall"] = out
        pg.barrier(BarrierOptions()).wait()
        if rank == 0:
            pg.send([torch.tensor([42.0])], 2, 7).wait()
        if rank == 2:
            r = torch.zeros(1)
            pg.recv([r], 0, 7).wait()
            res["recv"] = r
        return res

    outs = _ranks(W, body)
    for rank, r in enumerate(outs):
        assert torch.equal(r["allreduce"], torch.full((4,), 6.0))
        assert torch.equal(r["allgather"], torch.tensor([[0.0, 0.0], [1.0, 1.0], [2.0, 2.0]]))
        assert torch.equal(r["broadcast"], torch.full((3,), 1.0))
        assert torch.equal(r["alltoall"], torch.tensor([rank + 0.0, rank + 10.0, rank + 20.0]))
    assert outs[2]["recv"].item() == 42.0


def test_gloo_resiliency_peer_shutdown():
    """After the last rank disappears, survivors' collectives must fail fast (not hang)."""
    W = 3
    store = _store()
    pgs = [ProcessGroupGloo(timeout=timedelta(seconds=2)) for _ in range(W)]
    barrier = threading.Barrier(W)

    def one(rank):
        pg = pgs[rank]
        pg.configure(f"127.0.0.1:{store.port}/res/0", f"rep{rank}", rank, W)
        o = AllreduceOptions()
        o.reduceOp = ReduceOp.SUM
        t = torch.ones(2)
        pg.allreduce([t], o).wait()
        barrier.wait()
        if rank == W - 1:
            pg.shutdown()
            return "left"
        t0 = time.time()
        try:
            pg.allreduce([torch.ones(2)], o).wait()
        except Exception as e:  # noqa: BLE001
            return ("error", time.time() - t0, str(e))
        return ("no error", time.time() - t0, "")

    with ThreadPoolExecutor(max_workers=W) as ex:
        out = list(ex.map(one, range(W)))
    assert out[W - 1] == "left"
    for r in out[: W - 1]:
        assert r[0] == "error", r
        assert r[1] < 10.0
    # survivors can form a smaller group again
    def again(rank):
        pgs[rank].configure(f"127.0.0.1:{store.port}/res/1", f"rep{rank}", rank, W - 1)
        t = torch.ones(1)
        o = AllreduceOptions()
        o.reduceOp = ReduceOp.SUM
        pgs[rank].allreduce([t], o).wait()
        return t.item()

    with ThreadPoolExecutor(max_workers=W - 1) as ex:
        assert list(ex.map(again, range(W - 1))) == [2.0, 2.0]


def test_error_swallowing_wrapper():
    inner = ProcessGroupDummy(0, 1)
    pg = ErrorSwallowingProcessGroupWrapper(inner)
    t = torch.ones(2)
    o = AllreduceOptions()
    assert pg.allreduce([t], o).wait() is True and pg.error() is None
    err = RuntimeError("bad")
    pg.report_error(err)
    assert pg.error() is err and pg.errored() is err
    w = pg.allreduce([t], o)
    assert w.wait() is True  # no-op dummy while errored
    pg.configure("a", "r", 0, 1)
    assert pg.error() is None and inner.configure_count == 1


def test_fake_pg_injects_future_error():
    inner = ProcessGroupDummy(0, 1)
    pg = FakeProcessGroupWrapper(inner)
    o = AllreduceOptions()
    pg.report_future_error(RuntimeError("injected"))
    w = pg.allreduce([torch.ones(1)], o)
    with pytest.raises(RuntimeError, match="injected"):
        w.get_future().wait()
    # only the next collective is affected
    pg.allreduce([torch.ones(1)], o).get_future().wait()


def test_managed_process_group_delegates_to_manager():
    manager = Mock(spec=Manager)
    manager.num_participants.return_value = 123
    manager._pg = ProcessGroupDummy(0, 1)
    pg = ManagedProcessGroup(manager)
    t = torch.zeros(2)
    o = AllreduceOptions()
    o.reduceOp = ReduceOp.AVG
    pg.allreduce([t], o)
    manager.allreduce.assert_called_once()
    assert manager.allreduce.call_args.kwargs["reduce_op"] == ReduceOp.AVG
    assert pg.size() == 123
    assert pg.getBackendName() == "torchft-dummy"


def test_baby_gloo_multirank_and_respawn():
    W = 2

    def body(rank, pg):
        o = AllreduceOptions()
        o.reduceOp = ReduceOp.SUM
        t = torch.full((3,), float(rank + 1))
        w = pg.allreduce([t], o)
        w.wait()
        fut_t = torch.full((2,), 1.0)
        fut = pg.allreduce([fut_t], o).get_future()
        fut.wait()
        assert pg.num_active_work() == 0
        return t.clone(), fut_t.clone(), pg._proc.pid

    store = _store()
    pgs = [ProcessGroupBabyGloo(timeout=timedelta(seconds=20)) for _ in range(W)]

    def one(rank, gen):
        pgs[rank].configure(f"127.0.0.1:{store.port}/baby/{gen}", f"rep{rank}", rank, W)
        return body(rank, pgs[rank])

    with ThreadPoolExecutor(max_workers=W) as ex:
        first = list(ex.map(lambda r: one(r, 0), range(W)))
        procs = [pg._proc for pg in pgs]
        second = list(ex.map(lambda r: one(r, 1), range(W)))
    for (t, f, _pid) in first + second:
        assert torch.equal(t, torch.full((3,), 3.0)) and torch.equal(f, torch.full((2,), 2.0))
    # reconfigure killed the old children
    for p in procs:
        assert not p.is_alive()
    assert {x[2] for x in first}.isdisjoint({x[2] for x in second})
    for pg in pgs:
        pg.shutdown()


def test_baby_gloo_timeout_kills_child():
    store = _store()
    pg = ProcessGroupBabyGloo(timeout=timedelta(seconds=1))
    # world of 2 with nobody else: the child's rendezvous cannot complete
    with pytest.raises((TimeoutError, RuntimeError, Exception)):
        pg.configure(f"127.0.0.1:{store.port}/babyto/0", "rep0", 0, 2)
    pg.shutdown()


def test_register_with_c10d_and_functional_collectives():
    store = _store()
    if not dist.is_initialized():
        dist.init_process_group("gloo", store=dist.PrefixStore("default", store), rank=0, world_size=1)
    try:
        pg = ProcessGroupDummy(0, 1)
        registered = pg.register("test_dummy_reg")
        assert registered is not None
        t = torch.ones(4)
        dist.all_reduce(t, group=registered)
        from torch.distributed import _functional_collectives as fc

        out = fc.all_reduce(torch.ones(2), "sum", registered)
        assert torch.equal(torch.as_tensor(out), torch.ones(2))
    finally:
        dist.destroy_process_group()
