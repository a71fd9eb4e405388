"""In-process failure injection for chaos testing.

Covers the failure classes of the reference's monarch example
(``examples/monarch/utils/failure.py:24-78``: SEGFAULT, KILL_PROC, COMMS abort, DEADLOCK of the
GIL; KILL_SLURM is an orchestrator-side kill, see ``examples/orchestrator``) plus one that only
exists on this data plane: ``STALL_PEER`` — the process stays alive and keeps heart-beating but
stops participating in collectives, which is exactly the case the bounded in-kernel waits of
``csrc/kernels/common.cuh`` (and the Lighthouse's heartbeat timeout) have to turn into a latched
error on the healthy replicas.

A trainer opts in with ``FailureInjector(manager).start()``; the injector listens on a loopback
TCP port (written to ``$TORCHFT_FAILURE_PORT_FILE`` or logged) for one-line commands, or is
driven directly with :meth:`FailureInjector.inject`. Nothing is installed unless asked for.
"""

from __future__ import annotations

import ctypes
import enum
import logging
import os
import socket
import threading
import time
from typing import Any, Optional

logger = logging.getLogger(__name__)

__all__ = ["Failure", "FailureInjector", "send_failure"]

PORT_FILE_ENV = "TORCHFT_FAILURE_PORT_FILE"


class Failure(enum.Enum):
    """Injectable failure kinds (reference: examples/monarch/utils/failure.py:24-30, plus ``STALL_PEER``)."""

    SEGFAULT = "segfault"      # SIGSEGV in native code
    KILL_PROC = "kill_proc"    # immediate exit(1), no cleanup
    COMMS = "comms"            # abort the fault-tolerant process group under the trainer
    DEADLOCK = "deadlock"      # hold the GIL in a native sleep: heartbeats continue (C++ thread), training stops
    STALL_PEER = "stall_peer"  # stop participating in collectives without dying


class FailureInjector:
    """In-process fault injector: ``inject(kind)`` directly, or ``start()`` to accept one-word commands on a loopback port."""

    def __init__(self, manager: Any = None, pg: Any = None, deadlock_secs: int = 70) -> None:
        self._manager = manager
        self._pg = pg if pg is not None else getattr(manager, "_pg", None)
        self._deadlock_secs = deadlock_secs
        self._sock: Optional[socket.socket] = None
        self._thread: Optional[threading.Thread] = None
        self.port: Optional[int] = None
        self.stalled = threading.Event()

    # -- the failures ------------------------------------------------------
    def inject(self, failure: Failure) -> None:
        logger.warning("failure injection: %s (pid %d)", failure.name, os.getpid())
        if failure is Failure.SEGFAULT:
            ctypes.string_at(0)  # read address 0 from native code
        elif failure is Failure.KILL_PROC:
            os._exit(1)
        elif failure is Failure.COMMS:
            if self._pg is None:
                raise RuntimeError("COMMS failure needs a process group")
            self._pg.abort()
        elif failure is Failure.DEADLOCK:
            # PyDLL keeps the GIL held across the call, freezing every Python thread
            libc = ctypes.PyDLL(None)
            libc.sleep.argtypes = (ctypes.c_uint,)
            libc.sleep(self._deadlock_secs)
        elif failure is Failure.STALL_PEER:
            self.stalled.set()
        else:  # pragma: no cover
            raise ValueError(failure)

    def maybe_stall(self, poll_s: float = 0.05) -> None:
        """Call once per step from the training loop: blocks forever once STALL_PEER was injected."""
        while self.stalled.is_set():
            time.sleep(poll_s)

    # -- remote trigger ------------------------------------------------------
    def start(self, port: int = 0) -> "FailureInjector":
        s = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
        s.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        s.bind(("127.0.0.1", port))
        s.listen(4)
        self._sock, self.port = s, s.getsockname()[1]
        path = os.environ.get(PORT_FILE_ENV)
        if path:
            with open(path, "w") as f:
                f.write(str(self.port))
        logger.info("failure injector listening on 127.0.0.1:%d", self.port)
        self._thread = threading.Thread(target=self._serve, name="tft_failure_injector", daemon=True)
        self._thread.start()
        return self

    def _serve(self) -> None:
        sock = self._sock
        while sock is not None:
            try:
                conn, _ = sock.accept()
            except OSError:
                return
            with conn:
                try:
                    word = conn.recv(64).decode().strip().lower()
                    failure = Failure(word)
                    conn.sendall(b"ok\n")
                except (ValueError, OSError):
                    try:
                        conn.sendall(b"unknown failure\n")
                    except OSError:
                        pass
                    continue
            self.inject(failure)

    def stop(self) -> None:
        if self._sock is not None:
            try:
                self._sock.close()
            finally:
                self._sock = None


def send_failure(port: int, failure: Failure, timeout: float = 5.0) -> str:
    """Ask the injector on ``127.0.0.1:port`` to fail; returns its acknowledgement."""
    with socket.create_connection(("127.0.0.1", port), timeout=timeout) as s:
        s.sendall(failure.value.encode() + b"\n")
        try:
            return s.recv(64).decode().strip()
        except OSError:
            return ""
