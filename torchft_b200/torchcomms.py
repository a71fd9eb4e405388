"""Adapter: a reconfigurable *communicator object* as a fault-tolerant ProcessGroup.

Parity with the reference's ``torchft/torchcomms.py:116-324`` (``ProcessGroupTorchComms``):
the communicator is created once (``enable_reconfigure=True``) and SURVIVES quorum
changes — ``configure()`` only swaps its membership:

1. publish ``comm.get_init_handle()`` under ``torchcomms_init_handle/{rank}`` in the
   quorum's store prefix,
2. gather the handles of all ``world_size`` ranks in rank order,
3. ``comm.reconfigure(uuid=quorum_id, init_handles=[...], timeout=...)`` and wait.

That is the same idea ``ProcessGroupB200`` implements natively for one NVSwitch domain
(peer-memory remap instead of communicator re-creation); this adapter exists for
communicators we do not own. The ``torchcomms`` package is not in this image, so the
adapter is duck-typed: anything with ``get_init_handle / reconfigure / all_reduce / …``
works (the unit test drives it with an in-process communicator), and
:func:`new_torchcomm` imports the real package lazily.

Unlike the reference there is no per-collective method body: one table maps each c10d
entry point to (communicator method, argument picker), and single-tensor-only ops are
validated in one place.
"""

from __future__ import annotations

import logging
from datetime import timedelta
from typing import Any, List, Optional, Tuple

import torch
from torch.distributed.distributed_c10d import ReduceOp, Work
from torch.futures import Future

from torchft_b200.process_group import ProcessGroup, create_store_client

logger = logging.getLogger(__name__)

__all__ = ["ProcessGroupTorchComms", "new_torchcomm"]

_HANDLE_KEY = "torchcomms_init_handle/{}"


def new_torchcomm(backend: str, device: torch.device, name: str = "torchft_b200", **kwargs: Any) -> Any:
    """Create a reconfigurable ``torchcomms.TorchComm`` (raises ImportError when the package is absent)."""

    try:
        import torchcomms  # type: ignore[import-not-found]
    except ImportError as e:  # pragma: no cover - not in this image
        raise ImportError(
            "torchcomms is not installed; on one NVSwitch domain use torchft_b200.ProcessGroupB200, "
            "which reconfigures in place natively"
        ) from e
    return torchcomms.new_comm(backend, device, name=name, enable_reconfigure=True, **kwargs)  # pragma: no cover


class _CommWork(Work):
    """c10d ``Work`` over a communicator work handle.

    The future is resolved eagerly (stream-ordered backends: the result tensors are
    valid for later work on the launch stream; host backends must call ``wait()``),
    which is what DDP-style ``get_future().then(...)`` chaining needs.
    """

    def __init__(self, inner: Any, value: object, device: torch.device) -> None:
        super().__init__()
        self._inner = inner
        self._fut: Future = Future(devices=[] if device.type == "cpu" else [device])
        self._fut.set_result(value)

    def wait(self, timeout: Optional[timedelta] = None) -> bool:
        if self._inner is not None:
            self._inner.wait()
        return True

    def is_completed(self) -> bool:
        return True if self._inner is None else bool(self._inner.is_completed())

    def get_future(self) -> Future:
        return self._fut


def _one(x: Any, what: str) -> Any:
    if isinstance(x, (list, tuple)):
        if len(x) != 1:
            raise ValueError(f"ProcessGroupTorchComms.{what}: exactly one tensor per call, got {len(x)}")
        return x[0]
    return x


def _op_of(comm_module_ops: Any, opts: Any) -> Any:
    op = opts if isinstance(opts, (ReduceOp, ReduceOp.RedOpType)) else getattr(opts, "reduceOp", ReduceOp.SUM)
    name = str(op).rsplit(".", 1)[-1].upper()
    return getattr(comm_module_ops, name, getattr(comm_module_ops, "SUM"))


class ProcessGroupTorchComms(ProcessGroup):
    """``ProcessGroup`` backed by one long-lived reconfigurable communicator."""

    def __init__(self, comm: Any, timeout: timedelta = timedelta(seconds=60), reduce_ops: Any = None) -> None:
        super().__init__(0, 1)
        self._comm: Any = comm
        self._timeout = timeout
        self._world = 1
        self._rank = 0
        self._backend = str(comm.get_backend()) if hasattr(comm, "get_backend") else type(comm).__name__
        self._device = comm.get_device() if hasattr(comm, "get_device") else torch.device("cpu")
        if reduce_ops is None:
            try:
                import torchcomms  # type: ignore[import-not-found]

                reduce_ops = torchcomms.ReduceOp
            except ImportError:
                reduce_ops = ReduceOp  # duck-typed communicators take c10d ops directly
        self._ops = reduce_ops
        self._error: Optional[Exception] = None

    # -- lifecycle ---------------------------------------------------------
    def configure(self, store_addr: str, replica_id: str, rank: int, world_size: int, quorum_id: Optional[int] = None,
                  group_rank: Optional[int] = None, group_world_size: Optional[int] = None,
                  global_ranks: Optional[List[int]] = None) -> None:
        if self._comm is None:
            raise RuntimeError("ProcessGroupTorchComms: communicator was finalized")
        store = create_store_client(store_addr, timeout=self._timeout)
        store.set(_HANDLE_KEY.format(rank), self._comm.get_init_handle())
        keys = [_HANDLE_KEY.format(r) for r in range(world_size)]
        store.wait(keys, self._timeout)
        handles = [store.get(k).decode("utf-8") for k in keys]
        self._comm.reconfigure(uuid=quorum_id, init_handles=handles, timeout=self._timeout).wait()
        self._rank, self._world = rank, world_size
        self._error = None

    @property
    def comm(self) -> Any:
        if self._comm is None:
            raise RuntimeError("ProcessGroupTorchComms: communicator was finalized")
        return self._comm

    def size(self) -> int:
        return self._world

    def getBackendName(self) -> str:
        return f"torchcomms:{self._backend}"

    def abort(self) -> None:
        comm, self._comm = self._comm, None
        if comm is not None:
            try:
                comm.finalize()
            except Exception:  # noqa: BLE001 - abort must not raise
                logger.debug("finalize raised during abort", exc_info=True)

    def shutdown(self) -> None:
        self.abort()

    def errored(self) -> Optional[Exception]:
        return self._error

    def set_timeout(self, timeout: timedelta) -> None:
        self._timeout = timeout

    # -- collectives: one table instead of a body per op -------------------------
    def _run(self, what: str, method: str, args: Tuple[Any, ...], result: object) -> Work:
        try:
            inner = getattr(self.comm, method)(*args, async_op=True)
        except Exception as e:  # noqa: BLE001
            self._error = e
            raise
        return _CommWork(inner, result, self._device)

    def allreduce(self, tensors: List[torch.Tensor], opts: Any) -> Work:
        return self._run("allreduce", "all_reduce", (_one(tensors, "allreduce"), _op_of(self._ops, opts)), tensors)

    def allreduce_coalesced(self, tensors: List[torch.Tensor], opts: Any) -> Work:
        return self.allreduce(tensors, opts)

    def allgather(self, output_tensors: List[List[torch.Tensor]], input_tensor: List[torch.Tensor], opts: Any) -> Work:
        return self._run("allgather", "all_gather", (_one(output_tensors, "allgather"), _one(input_tensor, "allgather")), output_tensors)

    def allgather_into_tensor_coalesced(self, output_tensors: List[torch.Tensor], input_tensors: List[torch.Tensor], opts: Any) -> Work:
        w = "allgather_into_tensor_coalesced"
        return self._run(w, "all_gather_single", (_one(output_tensors, w), _one(input_tensors, w)), output_tensors)

    def broadcast(self, tensor_list: List[torch.Tensor], opts: Any) -> Work:
        return self._run("broadcast", "broadcast", (_one(tensor_list, "broadcast"), int(opts.rootRank)), tensor_list)

    def reduce_scatter(self, output_tensors: List[torch.Tensor], input_tensors: List[List[torch.Tensor]], opts: Any) -> Work:
        w = "reduce_scatter"
        return self._run(w, "reduce_scatter", (_one(output_tensors, w), _one(input_tensors, w), _op_of(self._ops, opts)), output_tensors)

    def reduce_scatter_tensor_coalesced(self, output_tensors: List[torch.Tensor], input_tensors: List[torch.Tensor], opts: Any) -> Work:
        w = "reduce_scatter_tensor_coalesced"
        return self._run(w, "reduce_scatter_single", (_one(output_tensors, w), _one(input_tensors, w), _op_of(self._ops, opts)), output_tensors)

    def alltoall_base(self, output_buffer: torch.Tensor, input_buffer: torch.Tensor, output_split_sizes: List[int],
                      input_split_sizes: List[int], opts: Any) -> Work:
        return self._run("alltoall_base", "all_to_all_single", (output_buffer, input_buffer), output_buffer)

    def barrier(self, opts: Any = None) -> Work:
        return self._run("barrier", "barrier", (), None)

    def send(self, tensors: List[torch.Tensor], dst_rank: int, tag: int) -> Work:
        return self._run("send", "send", (_one(tensors, "send"), dst_rank), tensors)

    def recv(self, tensors: List[torch.Tensor], src_rank: int, tag: int) -> Work:
        return self._run("recv", "recv", (_one(tensors, "recv"), src_rank), tensors)

    def __repr__(self) -> str:
        return f"ProcessGroupTorchComms(backend={self._backend!r}, device={self._device}, world={self._world}, live={self._comm is not None})"
