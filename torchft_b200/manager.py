"""Per-rank fault-tolerance state machine.

One ``Manager`` per training process. Per step (reference
semantics: /root/reference/torchft/manager.py:148-1053):

1. ``start_quorum()``  - async (overlaps the forward pass): group barrier +
   Lighthouse quorum through the C++ ``ManagerServer``; on a quorum-id change
   the process group is *reconfigured* (peer-memory remap on ``ProcessGroupB200``);
   replicas behind ``max_step`` heal from an up-to-date peer via the checkpoint
   transport (NVLink P2P by default on CUDA).
2. ``allreduce()``     - fault-tolerant gradient reduction; errors are swallowed and
   latched, never raised. On ``ProcessGroupB200`` the 1/num_participants scale and
   the non-participant zero contribution are fused into the all-reduce kernel.
3. ``should_commit()`` - intra-group AND barrier; the optimizer steps only when
   every rank of the group saw no error and enough replicas participated.

B200-specific departures from the reference:

* the commit path synchronises only the streams that carry collectives/recovery
  (events), not the whole device;
* ``commit_gate`` exposes the verdict as a device int32 so a fused optimizer
  kernel can be gated without further host round trips.
"""

from __future__ import annotations

import concurrent.futures
import logging
import os
import socket
import traceback
import uuid
import weakref
from concurrent.futures import ThreadPoolExecutor
from datetime import timedelta
from enum import Enum
from typing import TYPE_CHECKING, Callable, Dict, List, Optional, TypeVar, cast

import torch
from torch.distributed import ReduceOp, TCPStore, Work
from torch.distributed.distributed_c10d import AllreduceOptions
from torch.futures import Future

from torchft_b200._C import ManagerClient, ManagerServer
from torchft_b200.checkpointing._rwlock import RWLock
from torchft_b200.checkpointing.transport import CheckpointTransport
from torchft_b200.futures import future_timeout
from torchft_b200.utils import get_stream_context, synchronize
from torchft_b200.work import DummyWork

if TYPE_CHECKING:
    from torchft_b200.process_group import ProcessGroup

MANAGER_ADDR_KEY = "manager_addr"
MANAGER_PORT_ENV = "TORCHFT_MANAGER_PORT"
REPLICA_ID_KEY = "replica_id"

# env overrides win over constructor arguments (reference: manager.py:78-87,251-268)
TIMEOUT_SEC_ENV = "TORCHFT_TIMEOUT_SEC"
QUORUM_TIMEOUT_SEC_ENV = "TORCHFT_QUORUM_TIMEOUT_SEC"
CONNECT_TIMEOUT_SEC_ENV = "TORCHFT_CONNECT_TIMEOUT_SEC"
QUORUM_RETRIES_ENV = "TORCHFT_QUORUM_RETRIES"
TORCH_FR_DUMP_TEMP_FILE_ENV = "TORCH_FR_DUMP_TEMP_FILE"

T = TypeVar("T")
S = TypeVar("S")


def get_timeout(timeout_sec_env: Optional[str], default_timeout: timedelta) -> timedelta:
    """Environment value (integer seconds) if present, else the default."""
    return timedelta(seconds=int(timeout_sec_env)) if timeout_sec_env is not None else default_timeout


def extract_trailing_digits(s: str) -> int:
    """``"replica_12"`` -> 12; 0 when the string does not end in digits."""

    digits = ""
    for ch in reversed(s):
        if not ch.isdigit():
            break
        digits = ch + digits
    return int(digits) if digits else 0


class WorldSizeMode(Enum):
    """How the reduction is normalised when more than ``min_replica_size`` replicas are alive.

    DYNAMIC: use every healthy replica, divide by the live count.
    FIXED_WITH_SPARES: exactly ``min_replica_size`` replicas contribute; the rest are
    hot spares that add zeros.
    """

    DYNAMIC = 0
    FIXED_WITH_SPARES = 1


class ExceptionWithTraceback(Exception):
    """Wraps the first error reported in a step together with its formatted traceback (reference: manager.py:141-145)."""

    def __init__(self, e: Exception) -> None:
        self.original_exception = e
        self.stack_trace = traceback.format_exc()
        super().__init__(f"{e}\n{self.stack_trace}")


class Manager:
    """Fault-tolerant training-loop manager (see module docstring).

    The replica group's TCPStore (``store_addr:store_port`` or ``MASTER_ADDR`` /
    ``MASTER_PORT``) must already be running. When you persist periodic
    checkpoints, save and restore :meth:`state_dict` alongside the model.
    """

    def __init__(
        self,
        pg: "ProcessGroup",
        load_state_dict: Optional[Callable[[T], None]],
        state_dict: Optional[Callable[[], T]],
        min_replica_size: int,
        use_async_quorum: bool = True,
        timeout: timedelta = timedelta(seconds=60),
        quorum_timeout: timedelta = timedelta(seconds=60),
        connect_timeout: timedelta = timedelta(seconds=60),
        rank: Optional[int] = None,
        world_size: Optional[int] = None,
        world_size_mode: WorldSizeMode = WorldSizeMode.DYNAMIC,
        store_addr: Optional[str] = None,
        store_port: Optional[int] = None,
        lighthouse_addr: Optional[str] = None,
        replica_id: Optional[str] = None,
        port: Optional[int] = None,
        hostname: str = socket.gethostname(),
        heartbeat_interval: timedelta = timedelta(milliseconds=100),
        checkpoint_transport: Optional[CheckpointTransport[Dict[str, T]]] = None,
        init_sync: bool = True,
        max_retries: Optional[int] = None,
        quorum_retries: int = 0,
    ) -> None:
        """
        Args:
            pg: reconfigurable process group spanning the replica dimension
            load_state_dict / state_dict: user state hooks used to heal a recovering replica
            min_replica_size: minimum number of participating replicas for a step to commit
            use_async_quorum: compute the quorum in the background during the forward pass
            timeout: default timeout for collectives, should_commit, checkpoint ops, wrap_future
            quorum_timeout: how long to wait for a quorum (set ~1h for LocalSGD/DiLoCo)
            connect_timeout: timeout for establishing control-plane connections
            rank / world_size: rank and size WITHIN the replica group (env RANK / WORLD_SIZE)
            store_addr / store_port: the replica group's TCPStore (env MASTER_ADDR / MASTER_PORT)
            lighthouse_addr: (group rank 0) lighthouse address (env TORCHFT_LIGHTHOUSE)
            replica_id: (group rank 0) human-readable id; a uuid suffix makes restarts unique
            port: (group rank 0) manager server port (env TORCHFT_MANAGER_PORT, else ephemeral)
            checkpoint_transport: heal transport; default P2P over NVLink on CUDA, HTTP on CPU
            init_sync: force a step-0 weight sync from the primary replica
            max_retries: raise after this many consecutive failed commits (None = never)
            quorum_retries: lighthouse quorum retries before the manager gives up
        """
        self.quorum_logger = logging.getLogger("torchft_quorums")
        self.commits_logger = logging.getLogger("torchft_commits")
        self.errors_logger = logging.getLogger("torchft_errors")

        self._load_state_dict_fns: Dict[str, Callable[[object], None]] = {}
        self._user_state_dicts: Dict[str, Callable[[], object]] = {}
        self._original_fr_dump_temp_file = os.environ.get(TORCH_FR_DUMP_TEMP_FILE_ENV)
        self._replica_id = replica_id

        self._timeout = get_timeout(os.environ.get(TIMEOUT_SEC_ENV), timeout)
        self._quorum_timeout = get_timeout(os.environ.get(QUORUM_TIMEOUT_SEC_ENV), quorum_timeout)
        self._connect_timeout = get_timeout(os.environ.get(CONNECT_TIMEOUT_SEC_ENV), connect_timeout)
        self._quorum_retries = int(os.environ.get(QUORUM_RETRIES_ENV, str(quorum_retries)))

        self._state_dict_lock = RWLock(timeout=self._timeout.total_seconds())
        self._is_state_dict_read_allowed = True
        if load_state_dict and state_dict:
            self.register_state_dict_fn("default", load_state_dict, state_dict)

        self._pending_state_dict: Optional[Dict[str, object]] = None
        self._use_async_quorum = use_async_quorum
        self._replica_world_size_mode = world_size_mode
        self._init_sync = init_sync
        self._max_retries = max_retries
        self._commit_failures = 0

        store_addr = store_addr or os.environ["MASTER_ADDR"]
        store_port = store_port or int(os.environ["MASTER_PORT"])
        self._group_rank = rank if rank is not None else int(os.environ["RANK"])
        self._group_world_size = world_size or int(os.environ["WORLD_SIZE"])
        self._min_replica_size = min_replica_size

        if checkpoint_transport is None:
            checkpoint_transport = self._default_transport()
        self._checkpoint_transport: CheckpointTransport[Dict[str, T]] = checkpoint_transport

        self._executor = ThreadPoolExecutor(max_workers=1, thread_name_prefix="async_quorum")
        self._quorum_future: Optional[concurrent.futures.Future] = None

        self._store = TCPStore(host_name=store_addr, port=store_port, is_master=False, wait_for_workers=False)
        self._pg = pg
        self._manager: Optional[ManagerServer] = None

        self._recovery_stream: Optional[torch.cuda.Stream] = torch.cuda.Stream() if torch.cuda.is_available() else None
        self._recovery_event: Optional[torch.cuda.Event] = None
        self._commit_gate: Optional[torch.Tensor] = None

        if self._group_rank == 0:
            if port is None:
                port = int(os.environ.get(MANAGER_PORT_ENV, 0))
            lighthouse_addr = lighthouse_addr or os.environ["TORCHFT_LIGHTHOUSE"]
            # uuid suffix: a fast restart must not be mistaken for the old incarnation
            suffix = str(uuid.uuid4())
            replica_id = suffix if not replica_id else f"{replica_id}:{suffix}"
            self._manager = ManagerServer(
                replica_id=replica_id,
                lighthouse_addr=lighthouse_addr,
                hostname=hostname,
                bind=f"[::]:{port}",
                store_addr=f"{store_addr}:{store_port}",
                world_size=self._group_world_size,
                heartbeat_interval=heartbeat_interval,
                connect_timeout=self._connect_timeout,
                quorum_retries=self._quorum_retries,
            )
            self._store.set(MANAGER_ADDR_KEY, self._manager.address())
            self._store.set(REPLICA_ID_KEY, replica_id)

        addr = self._store.get(MANAGER_ADDR_KEY).decode("utf-8")
        self._client = ManagerClient(addr, connect_timeout=self._connect_timeout)
        full_replica_id = self._store.get(REPLICA_ID_KEY).decode("utf-8")
        self._logger = _ManagerLogger(self, full_replica_id or "", self._group_rank)

        self._step = 0
        self._quorum_id = -1
        self._errored: Optional[ExceptionWithTraceback] = None
        self._healing = False
        self._batches_committed = 0
        self._participating_replica_rank: Optional[int] = None
        self._participating_replica_world_size = 0

        self._global_rank = (
            self._group_rank
            if self._replica_id is None
            else extract_trailing_digits(self._replica_id) * self._group_world_size + self._group_rank
        )
        self._update_fr_path()

    def _default_transport(self) -> CheckpointTransport[Dict[str, T]]:
        if torch.cuda.is_available() and os.environ.get("TORCHFT_B200_TRANSPORT", "p2p") == "p2p":
            from torchft_b200.checkpointing.p2p_transport import P2PTransport

            # receive IN PLACE into the live tensors: a heal allocates nothing (a second copy
            # of an 8B model + optimizer state does not fit next to the first in 180 GB)
            return P2PTransport(timeout=self._timeout, state_dict=self._heal_targets)
        from torchft_b200.checkpointing.http_transport import HTTPTransport

        return HTTPTransport(timeout=self._timeout, num_chunks=0)

    def _heal_targets(self) -> Dict[str, object]:
        """Same pytree shape as ``_manager_state_dict`` (destinations for an in-place heal)."""
        return {"user": {k: fn() for k, fn in self._user_state_dicts.items()}, "torchft": self.state_dict()}

    # ------------------------------------------------------------ state dict
    def allow_state_dict_read(self) -> None:
        if not self._is_state_dict_read_allowed:
            self._is_state_dict_read_allowed = True
            self._state_dict_lock.w_release()

    def disallow_state_dict_read(self) -> None:
        if self._is_state_dict_read_allowed:
            self._is_state_dict_read_allowed = False
            self._state_dict_lock.w_acquire()

    def register_state_dict_fn(self, key: str, load_state_dict: Callable[[T], None], state_dict: Callable[[], T]) -> None:
        assert key not in self._load_state_dict_fns and key not in self._user_state_dicts, f"duplicate state_dict key {key}"
        self._load_state_dict_fns[key] = cast(Callable[[object], None], load_state_dict)
        self._user_state_dicts[key] = state_dict

    def set_state_dict_fns(self, load_state_dict: Callable[[T], None], state_dict: Callable[[], T]) -> None:
        self._logger.warn("`set_state_dict_fns` is deprecated, please use `register_state_dict_fn` instead")
        self.register_state_dict_fn("set_state_dict_fns", load_state_dict, state_dict)

    def shutdown(self, wait: bool = True) -> None:
        """Stop the checkpoint transport, the manager server and the quorum thread."""
        self._checkpoint_transport.shutdown(wait=wait)
        if self._manager is not None:
            self._manager.shutdown()
        self._executor.shutdown(wait=wait)

    # -------------------------------------------------------------- allreduce
    @torch.profiler.record_function("torchft::manager::allreduce")
    def allreduce(self, tensor: torch.Tensor, should_quantize: bool = False, reduce_op: ReduceOp = ReduceOp.AVG) -> Work:
        """Fault-tolerant all-reduce across the participating replicas.

        AVG divides by ``num_participants()``. Errors never raise: the first one is
        latched (``errored()``), the returned work completes, later calls are
        no-ops, and ``should_commit`` will return False so the (possibly
        corrupted) tensor is discarded.
        """
        if self.errored():
            return DummyWork(tensor)
        self.wait_quorum()
        num_participants = self.num_participants()
        participating = self.is_participating()

        pg_op = reduce_op
        if reduce_op == ReduceOp.AVG:
            if not torch.is_floating_point(tensor):
                raise ValueError("average reduce op is only supported for floating point tensors")
            pg_op = ReduceOp.SUM

        try:
            native = self._native_allreduce(tensor, should_quantize, reduce_op, num_participants, participating)
            if native is not None:
                return _ManagedWork(self, native, tensor)

            if not participating:
                tensor.zero_()
            if should_quantize and tensor.is_cuda:
                from torchft_b200.collectives import allreduce_quantized

                work = allreduce_quantized([tensor], pg_op, self._pg, torch.cuda.current_stream())
            else:
                opts = AllreduceOptions()
                opts.reduceOp = pg_op
                work = self._pg.allreduce([tensor], opts)

            managed = _ManagedWork(self, work, tensor)
            if reduce_op == ReduceOp.AVG:

                @torch.profiler.record_function("torchft::manager::allreduce::callback")
                def normalize(fut: Future) -> torch.Tensor:
                    tensor.div_(num_participants)
                    return tensor

                managed.get_future().then(normalize)
            return managed
        except Exception as e:  # noqa: BLE001
            self._logger.exception(f"got exception in all reduce -- skipping remaining: {e}")
            self.report_error(e)
            return DummyWork(tensor)

    def _native_allreduce(self, tensor: torch.Tensor, should_quantize: bool, reduce_op: ReduceOp,
                          num_participants: int, participating: bool) -> Optional[Work]:
        """Fused path on ProcessGroupB200: scale + zero-contribution live inside the kernel."""
        pg = self._pg
        fused = getattr(pg, "allreduce_native", None)
        if fused is None or not tensor.is_cuda or not pg._native_ok(tensor):  # type: ignore[attr-defined]
            return None
        from torchft_b200.ops import _native

        ops = {ReduceOp.SUM: _native.OP_SUM, ReduceOp.AVG: _native.OP_SUM, ReduceOp.MAX: _native.OP_MAX, ReduceOp.MIN: _native.OP_MIN}
        if reduce_op not in ops:
            return None
        scale = 1.0 / max(num_participants, 1) if reduce_op == ReduceOp.AVG else 1.0
        if should_quantize and reduce_op in (ReduceOp.SUM, ReduceOp.AVG):
            return pg.allreduce_q8(tensor, tensor, None, scale=scale, contribute=participating)  # type: ignore[attr-defined]
        return fused(tensor, op=ops[reduce_op], scale=scale, contribute=participating)

    def alloc_symmetric(self, name: str, nbytes: int) -> Optional[torch.Tensor]:
        """Peer-visible buffer from the process group (``ProcessGroupB200.alloc_symmetric``) or ``None``
        when the group has no such memory. Tensors carved from it are all-reduced in place, zero-copy.
        Call with identical arguments on every replica before the next quorum."""
        fn = getattr(self._pg, "alloc_symmetric", None)
        return fn(name, nbytes) if callable(fn) else None

    def supports_fused_delta(self) -> bool:
        """True when ``allreduce_delta`` runs as ONE fused kernel (ProcessGroupB200)."""
        return hasattr(self._pg, "allreduce_q8") and torch.cuda.is_available()

    @torch.profiler.record_function("torchft::manager::allreduce_delta")
    def allreduce_delta(self, out: torch.Tensor, a: torch.Tensor, b: torch.Tensor, should_quantize: bool = True) -> Work:
        """``out = AVG over replicas of (a - b)`` -- DiLoCo's pseudo-gradient reduction.

        On ProcessGroupB200 with ``should_quantize`` the subtraction, fp8 quantisation,
        exchange, fp32 reduction and dequantisation are one kernel; elsewhere this is
        ``torch.sub`` followed by :meth:`allreduce`. ``out`` may alias ``a``. Same
        error-latching contract as :meth:`allreduce`.
        """
        if self.errored():
            return DummyWork(out)
        self.wait_quorum()
        n = self.num_participants()
        participating = self.is_participating()
        try:
            pg = self._pg
            if should_quantize and self.supports_fused_delta() and all(pg._native_ok(t) for t in (out, a, b)):  # type: ignore[attr-defined]
                work = pg.allreduce_q8(out, a, b, scale=1.0 / max(n, 1), contribute=participating)  # type: ignore[attr-defined]
                return _ManagedWork(self, work, out)
        except Exception as e:  # noqa: BLE001
            self._logger.exception(f"got exception in allreduce_delta -- skipping remaining: {e}")
            self.report_error(e)
            return DummyWork(out)
        torch.sub(a, b, out=out)
        return self.allreduce(out, should_quantize=should_quantize)

    def report_error(self, e: Exception) -> None:
        """Latch an error: the current step will not commit and the group reconfigures next step."""
        self._errored = ExceptionWithTraceback(e)

    def errored(self) -> Optional[ExceptionWithTraceback]:
        return self._errored

    def wrap_future(self, fut: Future, default: T, timeout: Optional[timedelta] = None) -> Future:
        """Future that never fails: errors/timeouts are reported to the manager and replaced by ``default``."""
        fut = future_timeout(fut, timeout or self._timeout)
        stream = torch.cuda.current_stream() if torch.cuda.is_available() else None

        def swallow(f: Future) -> T:
            with get_stream_context(stream):
                try:
                    return f.value()
                except Exception as e:  # noqa: BLE001
                    self._logger.exception(f"got exception in future -- skipping remaining: {e}")
                    self.report_error(e)
                    return default

        return fut.then(swallow)

    # ----------------------------------------------------------------- quorum
    def start_quorum(self, allow_heal: bool = True, shrink_only: bool = False, timeout: Optional[timedelta] = None) -> None:
        """Begin a new step: compute the quorum (async by default) and ready the group.

        Call before the forward pass (``OptimizerWrapper.zero_grad`` does). All
        replicas must pass the same ``allow_heal``.
        """
        if self._quorum_future is not None:
            self._quorum_future.result()
        self._errored = None
        self._healing = False
        self._quorum_future = self._executor.submit(
            self._async_quorum,
            allow_heal=allow_heal,
            shrink_only=shrink_only,
            quorum_timeout=timeout or self._quorum_timeout,
            curr_device=torch.cuda.current_device() if torch.cuda.is_available() else -1,
        )
        if not self._use_async_quorum:
            self.wait_quorum()
            if self._healing:
                # sync quorum: heal before the forward pass so this replica counts immediately
                self._apply_pending_state_dict()
                self._healing = False

    @torch.profiler.record_function("torchft::manager::wait_quorum")
    def wait_quorum(self) -> None:
        """Block until the quorum (and PG reconfiguration) for this step is done."""
        assert self._quorum_future is not None, "must call start_quorum before wait_quorum"
        self._quorum_future.result()

    @torch.profiler.record_function("torchft::manager::_async_quorum")
    def _async_quorum(self, allow_heal: bool, shrink_only: bool, quorum_timeout: timedelta, curr_device: int) -> None:
        try:
            torch.multiprocessing._set_thread_name("torchft_quorum")
        except Exception:  # pragma: no cover
            pass
        if curr_device >= 0 and torch.cuda.is_available():
            torch.cuda.set_device(curr_device)

        with torch.profiler.record_function("torchft::manager::_client::_quorum"):
            quorum = self._client._quorum(
                group_rank=self._group_rank,
                step=self._step,
                checkpoint_metadata=self._checkpoint_transport.metadata(),
                shrink_only=shrink_only,
                timeout=quorum_timeout,
                init_sync=self._init_sync,
                commit_failures=self._commit_failures,
            )

        quorum_id = quorum.quorum_id
        max_step = quorum.max_step
        heal = quorum.heal
        ranks_in_quorum = [
            extract_trailing_digits(rid.split(":")[0]) * self._group_world_size + self._group_rank
            for rid in quorum.replica_ids
        ]

        # async quorum: only replicas already at max_step contribute this step;
        # sync quorum (or no healing): everybody in the quorum does.
        if self._use_async_quorum or not allow_heal:
            self._participating_replica_rank = quorum.max_replica_rank
            self._participating_replica_world_size = quorum.max_world_size
        else:
            self._participating_replica_rank = quorum.replica_rank
            self._participating_replica_world_size = quorum.replica_world_size

        if self._replica_world_size_mode == WorldSizeMode.FIXED_WITH_SPARES:
            self._participating_replica_world_size = min(self._participating_replica_world_size, self._min_replica_size)
            if self._participating_replica_rank is not None and self._participating_replica_rank >= self._min_replica_size:
                self._participating_replica_rank = None

        if quorum_id != self._quorum_id:
            self.quorum_logger.info("", extra={
                "job_id": os.environ.get("JOB_ID", "unknown"), "replica_id": self._replica_id,
                "rank": self._group_rank, "quorum_id": quorum_id, "step": max_step})
            store_prefixed_addr = f"{quorum.store_address}/torchft/{quorum_id}/{self._group_rank}"
            self._logger.info(f"reconfiguring for {quorum_id=} {store_prefixed_addr=}")
            try:
                self._quorum_id = quorum_id
                with torch.profiler.record_function("torchft::manager::_pg::configure"):
                    self._pg.configure(
                        store_prefixed_addr,
                        self._replica_id if self._replica_id is not None else "0",
                        quorum.replica_rank,
                        quorum.replica_world_size,
                        quorum_id,
                        self._group_rank,
                        self._group_world_size,
                        ranks_in_quorum,
                    )
                self._update_fr_path()
                reset = getattr(torch._C._distributed_c10d, "_reset_fr_recording_nccl", None)
                if reset is not None and "nccl" in self._pg.getBackendName():
                    reset()
            except Exception as e:  # noqa: BLE001
                self._logger.exception(f"got exception in pg configure: {e}")
                self.report_error(e)
                return

        if allow_heal:
            with get_stream_context(self._recovery_stream):
                try:
                    if quorum.recover_dst_replica_ranks:
                        self._logger.info(f"peers need recovery from us {quorum.recover_dst_replica_ranks}")
                        with torch.profiler.record_function("torchft::manager::_checkpoint_transport::send_checkpoint"):
                            self._checkpoint_transport.send_checkpoint(
                                dst_ranks=quorum.recover_dst_replica_ranks,
                                step=max_step,
                                state_dict=self._manager_state_dict(),
                                timeout=self._timeout,
                            )
                    if heal:
                        self._healing = True
                        src_addr = quorum.recover_src_manager_address
                        self._logger.info(f"healing required, fetching checkpoint metadata from {src_addr=} {max_step=}")
                        src_client = ManagerClient(src_addr, connect_timeout=self._connect_timeout)
                        checkpoint_metadata = src_client._checkpoint_metadata(self._group_rank, timeout=self._timeout)
                        src_rank = quorum.recover_src_replica_rank
                        assert src_rank is not None, "must have a recover rank when healing"
                        with torch.profiler.record_function("torchft::manager::_checkpoint_transport::recv_checkpoint"):
                            # staged here; the user part is applied on the main thread
                            self._pending_state_dict = self._checkpoint_transport.recv_checkpoint(
                                src_rank=src_rank, metadata=checkpoint_metadata, step=max_step, timeout=self._timeout)
                        self.load_state_dict(cast(Dict[str, int], self._pending_state_dict["torchft"]))
                        self._step = max_step
                except Exception as e:  # noqa: BLE001
                    self._logger.exception(f"got exception in recovery: {e}")
                    self.report_error(e)
                self._recovery_event = (
                    torch.cuda.current_stream().record_event() if self._recovery_stream is not None else None
                )

    def _update_fr_path(self) -> None:
        """Flight-recorder dumps go to ``<TORCH_FR_DUMP_TEMP_FILE>_quorum_<id>/<global_rank>``."""
        if self._original_fr_dump_temp_file is not None:
            folder = f"{self._original_fr_dump_temp_file}_quorum_{self._quorum_id}"
            os.makedirs(folder, exist_ok=True)
            os.environ[TORCH_FR_DUMP_TEMP_FILE_ENV] = f"{folder}/{self._global_rank}"

    def _apply_pending_state_dict(self) -> None:
        assert self._healing, "must be in healing state"
        assert self._quorum_future is not None, "must call start_quorum before should_commit"
        self._quorum_future.result()
        pending = self._pending_state_dict
        if pending is None:
            assert self.errored(), "checkpoint was not staged and no error occured"
            return
        self._logger.info("applying pending state dict")
        assert len(self._load_state_dict_fns) > 0, "user load_state_dict is not initialized."
        user = cast(Dict[str, object], pending["user"])
        for key, fn in self._load_state_dict_fns.items():
            fn(user[key])
        self._pending_state_dict = None
        self._logger.info("Loaded state dict.")

    # ----------------------------------------------------------------- commit
    @torch.profiler.record_function("torchft::manager::should_commit")
    def should_commit(self, timeout: Optional[timedelta] = None) -> bool:
        """Decide (identically on every rank of the group) whether to step the optimizer.

        Call once per step after backward and before ``optimizer.step()``; only step
        when this returns True. Raises ``RuntimeError`` after more than ``max_retries``
        consecutive failures.
        """
        with torch.profiler.record_function("torchft::manager::should_commit::recovery_stream::synchronize"):
            if self._recovery_event is not None:
                self._recovery_event.synchronize()
                self._recovery_event = None
        with torch.profiler.record_function("torchft::manager::should_commit::current_stream::synchronize"):
            if torch.cuda.is_available():
                synchronize()

        if err := self._pg.errored():
            self.report_error(err)

        if self._healing:
            self._apply_pending_state_dict()

        enough_replicas = self.num_participants() >= self._min_replica_size
        local_should_commit = enough_replicas and self._errored is None
        should_commit = self._client.should_commit(self._group_rank, self._step, local_should_commit,
                                                   timeout=timeout or self._timeout)
        self._logger.info(f"should_commit={should_commit} enough_replicas={enough_replicas}, errored={self._errored}")
        self.commits_logger.info("", extra={
            "job_id": os.environ.get("JOB_ID", "unknown"), "replica_id": self._replica_id, "rank": self._group_rank,
            "quorum_id": self._quorum_id, "step": self._step, "commit_result": should_commit})

        self._checkpoint_transport.disallow_checkpoint()

        if self._commit_gate is not None:
            self._commit_gate.fill_(1 if should_commit else 0)

        if should_commit:
            self._step += 1
            self._batches_committed += self.num_participants()
            self._commit_failures = 0
        else:
            self._commit_failures += 1
            if self._max_retries is not None and self._commit_failures > self._max_retries:
                msg = (f"should_commit failed {self._commit_failures} times consecutively, "
                       f"exceeding max_retries={self._max_retries}")
                self._logger.exception(msg)
                raise RuntimeError(msg)
        return should_commit

    def commit_gate(self) -> torch.Tensor:
        """Device int32 that ``should_commit`` sets to 1/0: lets a fused optimizer kernel be gated
        on the verdict (``FlatAdamW.step(gate=manager.commit_gate())``)."""
        if self._commit_gate is None:
            dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
            self._commit_gate = torch.zeros(1, dtype=torch.int32, device=dev)
        return self._commit_gate

    # ------------------------------------------------------------------ state
    def load_state_dict(self, state_dict: Dict[str, int]) -> None:
        self._step = state_dict["step"]
        self._batches_committed = state_dict["batches_committed"]

    def _manager_state_dict(self) -> Dict[str, object]:
        with self._state_dict_lock.r_lock():
            assert len(self._user_state_dicts) > 0, "user state_dict is not initialized."
            return {"user": {k: fn() for k, fn in self._user_state_dicts.items()}, "torchft": self.state_dict()}

    def state_dict(self) -> Dict[str, int]:
        """``{"step", "batches_committed"}`` -- persist with your periodic checkpoints."""
        return {"step": self._step, "batches_committed": self._batches_committed}

    def current_step(self) -> int:
        return self._step

    def batches_committed(self) -> int:
        """Total batches committed across all replicas and steps."""
        return self._batches_committed

    def participating_rank(self) -> Optional[int]:
        """This replica's rank among the participants (None if not participating). Blocks on the quorum."""
        if self._quorum_future is None:
            return None
        self.wait_quorum()
        return self._participating_replica_rank

    def num_participants(self) -> int:
        """Number of replicas contributing to this step. Blocks on the quorum."""
        if self._quorum_future is None:
            return 0
        self.wait_quorum()
        assert self._participating_replica_world_size >= 0, "internal error"
        return self._participating_replica_world_size

    def is_participating(self) -> bool:
        if self._participating_replica_rank is None:
            return False
        if self._healing:
            assert self._use_async_quorum
            return False
        return True


class _ManagerLogger:
    def __init__(self, manager: Manager, replica_id: str, group_rank: int) -> None:
        self._logger = logging.getLogger(__name__)
        self._replica_id, self._group_rank, self._manager = replica_id, group_rank, manager

    def prefix(self) -> str:
        return f"[{self._replica_id}/{self._group_rank} - step {self._manager.current_step()}]"

    def info(self, msg: str) -> None:
        self._logger.info(f"{self.prefix()} {msg}")

    def warn(self, msg: str) -> None:
        self._logger.warning(f"{self.prefix()} {msg}")

    def exception(self, msg: str) -> None:
        self._logger.exception(f"{self.prefix()} {msg}")


# ------------------------------------------------------------------ managed work
class _ValueFuture(Future):
    """Minimal future handed to chained callbacks: only ``value()`` is meaningful (never blocks)."""

    def __init__(self, value: object) -> None:
        super().__init__()
        self._v = value

    def value(self) -> object:
        return self._v

    def wait(self) -> object:
        return self._v


class _ManagedFuture(Future):
    """Handle into a :class:`_ManagedWork` callback pipeline.

    ``then(cb)`` only *records* ``cb`` (lazily, nothing runs and the host never
    blocks); the pipeline executes when the work is waited on. ``wait()`` resolves
    the pipeline and returns its final value.
    """

    def __init__(self, work: "weakref.ReferenceType[_ManagedWork]") -> None:
        super().__init__()
        self._work = work

    def _owner(self) -> "_ManagedWork":
        w = self._work()
        assert w is not None, "managed work was garbage collected"
        return w

    def then(self, callback: Callable[[Future], S]) -> Future:  # type: ignore[override]
        w = self._owner()
        w._callbacks.append(callback)
        return _ManagedFuture(self._work)

    def wait(self) -> object:  # type: ignore[override]
        w = self._owner()
        w._materialize()
        assert w._final is not None
        return w._final.wait()

    def value(self) -> object:  # type: ignore[override]
        raise NotImplementedError("use wait(); this future only builds the callback pipeline")

    def done(self) -> bool:  # type: ignore[override]
        raise NotImplementedError("use wait(); this future only builds the callback pipeline")

    def add_done_callback(self, callback: Callable[[Future], None]) -> None:  # type: ignore[override]
        raise NotImplementedError("use then(); this future only builds the callback pipeline")

    def set_result(self, result: object) -> None:  # type: ignore[override]
        raise NotImplementedError("managed futures are completed by their work")

    def set_exception(self, result: object) -> None:  # type: ignore[override]
        raise NotImplementedError("managed futures are completed by their work")


class _ManagedWork(Work):
    """``Work`` whose ``wait()`` never raises and whose continuation callbacks run on the
    launching stream AFTER a stream dependency on the collective has been established.

    Callbacks registered through ``get_future().then(...)`` are kept in a list and
    composed into one continuation the first time the work is waited on; that
    continuation is routed through ``Manager.wrap_future`` so a failure or timeout
    anywhere in the pipeline is reported to the manager and swallowed.
    """

    def __init__(self, manager: Manager, work: Work, value: object) -> None:
        super().__init__()
        self._manager = manager
        self._inner = work
        self._value = value
        self._callbacks: List[Callable[[Future], object]] = []
        self._final: Optional[Future] = None
        self._stream: Optional[torch.cuda.Stream] = torch.cuda.current_stream() if torch.cuda.is_available() else None

    def _materialize(self) -> None:
        if self._final is not None:
            return
        fut = self._inner.get_future()
        callbacks, value, stream = list(self._callbacks), self._value, self._stream

        def pipeline(f: Future) -> object:
            with get_stream_context(stream):
                f.wait()  # stream dependency on the collective, not a host block for CUDA futures
                v = value
                for cb in callbacks:
                    v = cb(_ValueFuture(v))
                return v

        self._final = self._manager.wrap_future(fut.then(pipeline), value)

    def _assert_same_stream(self) -> None:
        if self._stream is not None:
            assert self._stream == torch.cuda.current_stream(), "wait on the stream that launched the collective"

    def wait(self, timeout: Optional[timedelta] = None) -> bool:
        self._assert_same_stream()
        try:
            with get_stream_context(self._stream):
                self._inner.wait()
                self._materialize()
                assert self._final is not None
                self._final.wait()
            return True
        except Exception as e:  # noqa: BLE001
            self._manager._logger.exception(f"got exception waiting for work {e}")
            self._manager.report_error(e)
            return False

    def block_current_stream(self, timeout: Optional[timedelta] = None) -> None:
        self._assert_same_stream()
        with get_stream_context(self._stream):
            blk = getattr(self._inner, "block_current_stream", None)
            if blk is not None and self._stream is not None:
                blk()
            else:  # no accelerator stream to block: fall back to a host wait
                self._inner.wait()
        self._materialize()

    def synchronize(self) -> None:
        self._assert_same_stream()
        if torch.cuda.is_available():
            self.block_current_stream()
        else:
            self._materialize()

    def get_future(self) -> Future:
        return _ManagedFuture(weakref.ref(self))
