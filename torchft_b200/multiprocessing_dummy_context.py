"""Thread-backed stand-in for a ``multiprocessing`` context.

Lets subprocess-hosted ("Baby") process groups run their child loop in a THREAD, which makes
them debuggable and usable in unit tests without spawning (reference:
torchft/multiprocessing_dummy_context.py). Only the subset the Baby groups use is provided:
``Process`` (start/join/kill/is_alive/pid) and ``Pipe``.

    class ThreadedBabyGloo(ProcessGroupBabyGloo):
        def _mp_context(self):
            return multiprocessing_dummy_context.get_context()
"""

from __future__ import annotations

import itertools
import queue
import threading
from typing import Any, Callable, Optional, Tuple

_PIDS = itertools.count(1_000_000)


class _Conn:
    def __init__(self, rx: "queue.Queue[Any]", tx: "queue.Queue[Any]") -> None:
        self._rx, self._tx = rx, tx
        self.closed = False

    def send(self, obj: Any) -> None:
        if self.closed:
            raise OSError("connection closed")
        self._tx.put(obj)

    def recv(self) -> Any:
        while True:
            if self.closed:
                raise EOFError("connection closed")
            try:
                item = self._rx.get(timeout=0.1)
            except queue.Empty:
                continue
            if item is _EOF:
                raise EOFError("peer closed")
            return item

    def poll(self, timeout: Optional[float] = 0.0) -> bool:
        if not self._rx.empty():
            return True
        if not timeout:
            return False
        try:
            item = self._rx.get(timeout=timeout)
        except queue.Empty:
            return False
        # put it back at the front: single consumer, so a private stash is enough
        q2: "queue.Queue[Any]" = queue.Queue()
        q2.put(item)
        while not self._rx.empty():
            q2.put(self._rx.get())
        while not q2.empty():
            self._rx.put(q2.get())
        return True

    def close(self) -> None:
        if not self.closed:
            self.closed = True
            self._tx.put(_EOF)


_EOF = object()


class _ThreadProcess:
    def __init__(self, target: Callable[..., Any], args: Tuple[Any, ...] = (), daemon: bool = True, **_: Any) -> None:
        self._thread = threading.Thread(target=target, args=args, daemon=daemon, name="tft_dummy_proc")
        self.pid = next(_PIDS)
        self.exitcode: Optional[int] = None

    def start(self) -> None:
        self._thread.start()

    def join(self, timeout: Optional[float] = None) -> None:
        self._thread.join(timeout)
        if not self._thread.is_alive():
            self.exitcode = 0

    def is_alive(self) -> bool:
        return self._thread.is_alive()

    def kill(self) -> None:
        # threads cannot be killed: closing the pipes (done by the caller) ends the child loop
        self.exitcode = -9

    terminate = kill


class _Context:
    Process = _ThreadProcess

    @staticmethod
    def Pipe(duplex: bool = True) -> Tuple[_Conn, _Conn]:  # noqa: N802
        a: "queue.Queue[Any]" = queue.Queue()
        b: "queue.Queue[Any]" = queue.Queue()
        if duplex:
            return _Conn(a, b), _Conn(b, a)
        # (receive end, send end) like multiprocessing.Pipe(duplex=False)
        return _Conn(a, queue.Queue()), _Conn(queue.Queue(), a)


def get_context(method: Optional[str] = None) -> _Context:
    return _Context()
