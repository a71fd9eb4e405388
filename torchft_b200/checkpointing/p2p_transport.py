"""NVLink peer-to-peer live-recovery transport (B200-native default on one node).

The reference heals a rejoining replica by copying the source's state_dict
GPU -> CPU, serving it over HTTP/TCP, and copying CPU -> GPU on the receiver
(/root/reference/torchft/checkpointing/http_transport.py:219-284), or by one
NCCL send/recv per tensor (pg_transport.py:214-303). Here the bytes never leave
HBM/NVLink and no collective is involved:

1. ``send_checkpoint`` builds a *manifest*: pytree spec, small non-tensor leaves
   (pickled), and for every CUDA tensor the CUDA-IPC handle of the allocation
   that contains it plus (offset, nbytes, dtype, shape, stride). Nothing is
   copied. The manifest is served by a tiny HTTP endpoint (control plane).
2. ``recv_checkpoint`` fetches the manifest, maps the source allocations into
   its address space, and launches ONE ``heal_copy`` kernel that streams every
   tensor with 16-byte peer loads into its destination -- optionally *in place*
   into the receiver's existing tensors (``state_dict=`` callable), so a heal
   allocates nothing.
3. The source keeps training state immutable until ``disallow_checkpoint``,
   which waits for in-flight pulls (reader lock) exactly like the HTTP transport.

Roofline: bytes / 770 GB/s (one NVLink direction); the source's SMs are not used.
"""

from __future__ import annotations

import io
import os
import logging
import pickle
import socket
import threading
from datetime import timedelta
from http.server import BaseHTTPRequestHandler, ThreadingHTTPServer
from typing import Any, Callable, Dict, Generic, List, Optional, Sequence, Tuple, TypeVar

import torch
from torch.utils import _pytree as pytree

from torchft_b200.checkpointing._rwlock import RWLock
from torchft_b200.checkpointing.transport import CheckpointTransport
from torchft_b200.ops import _native

logger = logging.getLogger(__name__)


from torchft_b200.checkpointing.transport import advertise_host as _advertise_host  # noqa: E402

T = TypeVar("T")

CHUNK_BYTES = 4 << 20


def device_copy(entries: Sequence[Tuple[int, int, int]], stream: Optional[torch.cuda.Stream] = None,
                blocks: int = 128, chunk_bytes: int = CHUNK_BYTES, bulk: Optional[bool] = None) -> None:
    """Copy ``(src_ptr, dst_ptr, nbytes)`` ranges with one kernel launch. Either side may be a mapped peer pointer;
    the kernel issues the NVLink traffic.

    ``bulk`` selects the TMA variant (``heal_copy_bulk_kernel``: one thread per CTA drives a ring of ``cp.async.bulk``
    global->shared->global transfers, no data in registers) instead of the LSU variant (``heal_copy_kernel``: 8 x 16-byte
    loads in flight per thread). Default: env ``TORCHFT_B200_HEAL_BULK`` (off). Bulk needs 16-byte aligned ranges.
    """

    K = _native.load()
    if bulk is None:
        bulk = os.environ.get("TORCHFT_B200_HEAL_BULK", "0") == "1"
    rows, chunk0 = [], 0
    for src, dst, n in entries:
        if n <= 0:
            continue
        rows.append((src, dst, n, chunk0))
        chunk0 += (n + chunk_bytes - 1) // chunk_bytes
        if bulk and ((src | dst) & 15):
            bulk = False
    if not rows:
        return
    table = torch.tensor(rows, dtype=torch.int64).cuda(non_blocking=False)
    s = stream if stream is not None else torch.cuda.current_stream()
    with torch.cuda.stream(s):
        if bulk and chunk_bytes % 16 == 0:
            # 64 KiB of shared memory per CTA -> 3 CTAs per SM; enough CTAs to put > 2 MB (NVLink) / > 8 MB (HBM) in flight
            K.heal_copy_bulk(table.data_ptr(), len(rows), chunk0, chunk_bytes, max(1, min(max(blocks, 296), chunk0)), int(s.cuda_stream))
        else:
            K.heal_copy(table.data_ptr(), len(rows), chunk0, chunk_bytes, max(1, min(blocks, chunk0)), int(s.cuda_stream))
        table.record_stream(s)


class _Server(ThreadingHTTPServer):
    address_family = socket.AF_INET6
    daemon_threads = True
    request_queue_size = 256

    def server_bind(self) -> None:
        try:
            self.socket.setsockopt(socket.IPPROTO_IPV6, socket.IPV6_V6ONLY, 0)
        except OSError:
            pass
        super().server_bind()


class P2PTransport(CheckpointTransport[T], Generic[T]):
    """Receiver-pull heal over NVLink. Falls back to host pickling for CPU tensors."""

    def __init__(self, timeout: timedelta = timedelta(seconds=60),
                 state_dict: Optional[Callable[[], T]] = None, blocks: int = 128) -> None:
        self._timeout = timeout
        self._inplace_state_dict = state_dict
        self._blocks = blocks
        self._lock = RWLock(timeout=timeout.total_seconds())
        self._lock.w_acquire()
        self._allowed = False
        self._step = -1
        self._manifest: Optional[bytes] = None
        self._keepalive: Any = None
        self._sessions: Dict[str, Any] = {}  # session id -> its open connection (closing it releases its read lock)
        self._sess_lock = threading.Lock()
        self._lease_s = max(1.0, timeout.total_seconds())
        self.last_recv_bytes = 0
        self.last_recv_ms = 0.0
        transport = self

        class Handler(BaseHTTPRequestHandler):
            protocol_version = "HTTP/1.1"

            def log_message(self, fmt: str, *args: Any) -> None:
                logger.debug("p2p_transport: " + fmt, *args)

            def _send(self, code: int, body: bytes) -> None:
                self.send_response(code)
                self.send_header("Content-Length", str(len(body)))
                self.end_headers()
                self.wfile.write(body)

            def do_GET(self) -> None:  # noqa: N802
                # /manifest/{step}/{session}: the read lock is held for as long as THIS connection lives. The receiver
                # answers with one byte when its NVLink pull is done; a receiver that dies or is partitioned closes
                # (or times out) the socket and the lock is released either way -- a leaked reader can never wedge the
                # source's disallow_checkpoint (the reference's HTTP transport has the same property
                # because it only holds the lock during the request, http_transport.py:38-298).
                parts = self.path.strip("/").split("/")
                try:
                    if len(parts) == 3 and parts[0] == "manifest":
                        step, sess = int(parts[1]), parts[2]
                        transport._lock.r_acquire()
                        try:
                            if step != transport._step or transport._manifest is None:
                                return self._send(400, f"invalid checkpoint requested: serving {transport._step} but got {step}".encode())
                            with transport._sess_lock:
                                transport._sessions[sess] = self.connection
                            try:
                                self._send(200, transport._manifest)
                                self.connection.settimeout(transport._lease_s)
                                try:
                                    self.connection.recv(1)  # b"D" = done, b"" = peer went away, timeout = lease over
                                except OSError:
                                    pass
                            finally:
                                with transport._sess_lock:
                                    transport._sessions.pop(sess, None)
                                self.close_connection = True
                        finally:
                            transport._lock.r_release()
                        return
                    self._send(404, b"unknown path")
                except TimeoutError as e:
                    self._send(503, f"checkpoint not available: {e}".encode())
                except (BrokenPipeError, ConnectionResetError):
                    pass

        self._server = _Server(("::", 0), Handler)
        self._thread = threading.Thread(target=self._server.serve_forever, name="tft_p2p_ckpt", daemon=True)
        self._thread.start()

    # ------------------------------------------------------------------ source
    def metadata(self) -> str:
        port = self._server.socket.getsockname()[1]
        return f"http://{_advertise_host()}:{port}"

    def send_checkpoint(self, dst_ranks: List[int], step: int, state_dict: T, timeout: timedelta) -> None:
        K = _native.load()
        leaves, spec = pytree.tree_flatten(state_dict)
        items: List[Dict[str, Any]] = []
        handles: Dict[int, str] = {}
        keep = []
        # make sure everything the state_dict views has been produced
        if torch.cuda.is_available():
            torch.cuda.current_stream().synchronize()
        for leaf in leaves:
            host = isinstance(leaf, torch.Tensor) and not leaf.is_cuda and leaf.numel() * leaf.element_size() >= 4096
            if (isinstance(leaf, torch.Tensor) and leaf.is_cuda) or host:
                t = leaf.detach()
                if isinstance(t, torch.distributed.tensor.DTensor):  # type: ignore[attr-defined]
                    t = t.to_local()
                if host:
                    # CPU tensors (e.g. DiLoCo's pinned backup weights) ride the same NVLink pull: stage them on the GPU
                    # once instead of embedding them in the manifest (that doubled them as hex in memory)
                    t = t.to(torch.device("cuda", torch.cuda.current_device()), non_blocking=True)
                # bytes actually backing this view
                if not t.is_contiguous():
                    t = t.contiguous()
                keep.append(t)
                ptr = t.data_ptr()
                nbytes = t.numel() * t.element_size()
                base, size = K.address_range(ptr) if nbytes else (0, 0)
                if nbytes and base not in handles:
                    handles[base] = K.ipc_get_handle(base).hex()
                items.append({"k": "cuda", "base": base, "off": ptr - base, "nbytes": nbytes,
                              "dtype": str(t.dtype).replace("torch.", ""), "shape": list(t.shape),
                              "device": t.device.index, "host": host})
            else:
                items.append({"k": "obj", "data": pickle.dumps(leaf)})
        if torch.cuda.is_available():
            torch.cuda.current_stream().synchronize()  # staged host tensors have landed
        self._keepalive = keep
        self._manifest = pickle.dumps({"spec": spec, "items": items, "handles": handles, "pid": _pid(),
                                       "host": socket.gethostname()})
        self._step = step
        if not self._allowed:
            self._allowed = True
            self._lock.w_release()

    def _drop_sessions(self) -> None:
        """Cut the connections of sessions that are still open: their handlers wake up and release the read lock."""
        with self._sess_lock:
            conns = list(self._sessions.items())
        for sess, conn in conns:
            logger.warning("p2p_transport: dropping stale heal session %s", sess)
            try:
                conn.shutdown(socket.SHUT_RDWR)
            except OSError:
                pass

    def disallow_checkpoint(self) -> None:
        if not self._allowed:
            return
        try:
            self._lock.w_acquire()  # waits for sessions still pulling
        except TimeoutError:
            # a reader outlived the lock timeout (receiver wedged or partitioned): evict it rather than failing the
            # training step, and only give up if even that does not free the lock
            self._drop_sessions()
            self._lock.w_acquire()
        self._allowed = False  # flipped only once the write lock is really held, so the two can never disagree
        self._manifest = None
        self._keepalive = None

    # ---------------------------------------------------------------- receiver
    def recv_checkpoint(self, src_rank: int, metadata: str, step: int, timeout: timedelta) -> T:
        K = _native.load()
        import uuid

        import http.client
        from urllib.parse import urlparse

        sess = uuid.uuid4().hex
        u = urlparse(metadata)
        conn = http.client.HTTPConnection(u.hostname, u.port, timeout=timeout.total_seconds())
        try:
            conn.request("GET", f"/manifest/{step}/{sess}")
            resp = conn.getresponse()
            body = resp.read()
        except (socket.timeout, TimeoutError) as e:
            conn.close()
            raise TimeoutError(f"checkpoint manifest from rank {src_rank} timed out: {e}") from e
        except OSError as e:
            conn.close()
            raise RuntimeError(f"checkpoint fetch from rank {src_rank} failed: {e}") from e
        if resp.status != 200:
            conn.close()
            if resp.status == 503:
                raise TimeoutError(f"checkpoint fetch from rank {src_rank} failed: {resp.status} {body.decode(errors='replace')}")
            raise RuntimeError(f"checkpoint fetch from rank {src_rank} failed: {resp.status} {body.decode(errors='replace')}")
        man = pickle.loads(body)
        if man.get("host", socket.gethostname()) != socket.gethostname():
            conn.close()
            raise RuntimeError(f"P2PTransport only works inside one host (NVLink/CUDA-IPC): source is on {man['host']!r}, "
                               f"we are on {socket.gethostname()!r}; use HTTPTransport or PGTransport across hosts")
        opened: Dict[int, int] = {}
        try:
            same_process = man["pid"] == _pid()
            for base, hx in man["handles"].items():
                opened[base] = base if same_process else K.ipc_open_handle(bytes.fromhex(hx))
            # In-place targets are matched by KEY PATH, not by position: a freshly restarted replica's
            # state_dict usually has a different shape than the survivor's (e.g. a torch optimizer
            # creates its per-parameter state lazily, so it is empty before the first step). Leaves
            # without a matching target are simply allocated.
            dst_leaves: Optional[List[Any]] = None
            if self._inplace_state_dict is not None:
                targets = {pytree.keystr(kp): leaf
                           for kp, leaf in pytree.tree_flatten_with_path(self._inplace_state_dict())[0]}
                index_tree = pytree.tree_unflatten(list(range(len(man["items"]))), man["spec"])
                dst_leaves = [None] * len(man["items"])
                for kp, idx in pytree.tree_flatten_with_path(index_tree)[0]:
                    dst_leaves[idx] = targets.get(pytree.keystr(kp))
            out: List[Any] = []
            host_moves: List[Tuple[int, torch.Tensor, Optional[torch.Tensor]]] = []
            entries: List[Tuple[int, int, int]] = []
            total = 0
            dev = torch.device("cuda", torch.cuda.current_device())
            for i, it in enumerate(man["items"]):
                if it["k"] == "obj":
                    data = it["data"]
                    out.append(pickle.loads(bytes.fromhex(data) if isinstance(data, str) else data))
                    continue
                dtype = getattr(torch, it["dtype"])
                dst = None
                host_target = None
                if it.get("host") and dst_leaves is not None and isinstance(dst_leaves[i], torch.Tensor) and not dst_leaves[i].is_cuda:
                    host_target = dst_leaves[i]
                if dst_leaves is not None and isinstance(dst_leaves[i], torch.Tensor):
                    cand = dst_leaves[i]
                    cand_local = cand.to_local() if hasattr(cand, "to_local") else cand
                    if cand_local.is_cuda and cand_local.is_contiguous() and cand_local.dtype == dtype and list(cand_local.shape) == it["shape"]:
                        dst = cand
                        dptr = cand_local.data_ptr()
                if dst is None:
                    dst = torch.empty(it["shape"], dtype=dtype, device=dev)
                    dptr = dst.data_ptr()
                if it["nbytes"]:
                    entries.append((opened[it["base"]] + it["off"], dptr, it["nbytes"]))
                    total += it["nbytes"]
                if it.get("host"):
                    host_moves.append((len(out), dst, host_target))
                out.append(dst)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            device_copy(entries, blocks=self._blocks)
            e.record()
            e.synchronize()
            self.last_recv_bytes, self.last_recv_ms = total, s.elapsed_time(e)
            for idx, staged, target in host_moves:  # leaves that were CPU tensors on the source go back to the host
                if target is not None and target.shape == staged.shape and target.dtype == staged.dtype:
                    target.copy_(staged)
                    out[idx] = target
                else:
                    out[idx] = staged.cpu()
            return pytree.tree_unflatten(out, man["spec"])
        finally:
            for base, p in opened.items():
                if p != base:
                    try:
                        K.ipc_close_handle(p)
                    except RuntimeError:
                        pass
            try:
                conn.sock.sendall(b"D")  # done: the source releases this session's read lock
            except Exception:  # noqa: BLE001 - closing the socket releases it as well
                pass
            conn.close()

    def shutdown(self, wait: bool = True) -> None:
        self._server.shutdown()
        self._server.server_close()
        if wait:
            self._thread.join(timeout=5)


def _pid() -> int:
    import os

    return os.getpid()
