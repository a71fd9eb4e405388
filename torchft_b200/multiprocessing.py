"""Request/response channel to a helper process that cannot hang its owner.

Role of the reference's ``torchft/multiprocessing.py:16-38`` (``_MonitoredPipe``) for the
subprocess-hosted process groups in :mod:`torchft_b200.baby`, with two additions the
subprocess design needs in practice:

* **liveness**: while waiting, the channel polls in short slices and asks an optional
  ``alive()`` probe (normally ``Process.is_alive``); a dead peer surfaces at once as
  ``RuntimeError`` instead of after the full timeout;
* **remote failures keep their traceback**: the peer ships failures as a :class:`RemoteFailure`
  envelope (exception + formatted remote traceback) built by :func:`failure_of`; raising it
  chains the remote text as ``__cause__``. Bare exception objects are still re-raised for
  compatibility with peers that send them directly.
"""

from __future__ import annotations

import time
import traceback
from dataclasses import dataclass
from datetime import timedelta
from typing import Any, Callable, Optional, Union

__all__ = ["MonitoredPipe", "RemoteFailure", "failure_of"]

_SLICE_S = 0.05


class RemoteTraceback(Exception):
    """Carries the text of a traceback that happened in another process."""


@dataclass
class RemoteFailure:
    """What a helper process sends instead of a result when a command failed."""

    error: BaseException
    remote_traceback: str = ""

    def throw(self) -> None:
        cause = RemoteTraceback("\n" + self.remote_traceback) if self.remote_traceback else None
        raise self.error from cause


def failure_of(e: BaseException) -> RemoteFailure:
    """Envelope for ``e`` including the traceback of the current ``except`` block."""
    return RemoteFailure(e, traceback.format_exc())


class MonitoredPipe:
    """One end of a ``multiprocessing`` pipe with deadline-bounded, liveness-aware ``recv``."""

    def __init__(self, conn: Any, alive: Optional[Callable[[], bool]] = None) -> None:
        self._conn = conn
        self._alive = alive

    def send(self, obj: object) -> None:
        self._conn.send(obj)

    def recv(self, timeout: Union[float, timedelta]) -> object:
        """Next message, or raise: ``TimeoutError`` when the deadline passes, ``RuntimeError`` when the
        peer process is gone, the peer's own exception when it reported a failure."""
        budget = timeout.total_seconds() if isinstance(timeout, timedelta) else float(timeout)
        deadline = time.monotonic() + budget
        while True:
            left = deadline - time.monotonic()
            if self._conn.poll(max(0.0, min(left, _SLICE_S))):
                break
            if self._alive is not None and not self._alive() and not self._conn.poll(0):
                raise RuntimeError("peer process exited without answering")
            if left <= 0:
                raise TimeoutError(f"pipe.recv() timed out after {budget} seconds")
        msg = self._conn.recv()
        if isinstance(msg, RemoteFailure):
            msg.throw()
        if isinstance(msg, BaseException):
            raise msg
        return msg

    def close(self) -> None:
        self._conn.close()

    def closed(self) -> bool:
        return bool(self._conn.closed)


_MonitoredPipe = MonitoredPipe  # name used by the reference
