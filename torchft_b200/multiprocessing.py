"""Pipe with timeouts and exception transport for subprocess-hosted process groups
(reference: torchft/multiprocessing.py:16-38)."""

from __future__ import annotations

from datetime import timedelta
from multiprocessing.connection import Connection
from typing import Union


class _MonitoredPipe:
    def __init__(self, pipe: "Connection[object, object]") -> None:  # type: ignore[type-arg]
        self._pipe = pipe

    def send(self, obj: object) -> None:
        self._pipe.send(obj)

    def recv(self, timeout: Union[float, timedelta]) -> object:
        """Receive one object; raises ``TimeoutError`` if nothing arrives in time and
        re-raises any exception object the peer sent."""
        secs = timeout.total_seconds() if isinstance(timeout, timedelta) else float(timeout)
        if not self._pipe.poll(secs):
            raise TimeoutError(f"pipe.recv() timed out after {secs} seconds")
        out = self._pipe.recv()
        if isinstance(out, Exception):
            raise out
        return out

    def close(self) -> None:
        self._pipe.close()

    def closed(self) -> bool:
        return self._pipe.closed
