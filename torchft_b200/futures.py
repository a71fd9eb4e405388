"""User-space timeouts for futures, CUDA streams and blocking regions.

Parity with the reference's ``torchft/futures.py:50-354`` (``future_timeout``,
``future_wait``, ``stream_timeout``, ``context_timeout`` and a watchdog that
kills the process when the timeout machinery itself wedges), with a different
engine: instead of an asyncio event loop there is ONE timer thread draining a
deadline heap. CUDA events used by ``stream_timeout`` are only created and
destroyed on the caller's thread (a queue hands finished events back), which
avoids CUDA calls from the timer thread racing with a poisoned context.
"""

from __future__ import annotations

import heapq
import itertools
import os
import sys
import threading
import time
from contextlib import contextmanager
from datetime import timedelta
from typing import Callable, Generator, List, Optional, Tuple, TypeVar

import torch
from torch.futures import Future

T = TypeVar("T")

WATCHDOG_TIMEOUT_SEC_ENV = "TORCHFT_WATCHDOG_TIMEOUT_SEC"


class _Handle:
    """Cancellable timer entry. It owns the callback so that ``cancel()`` can drop it at once: the heap entry itself only
    leaves the heap at its deadline, and a callback closure typically references the future being guarded (and through
    it the result tensor) -- kept until the deadline, every collective of every step would pin its tensor for a full op
    timeout (FSDP2 hands the manager a fresh gradient shard per parameter group per step: 16 GB per step at 8B)."""

    __slots__ = ("cancelled", "fired", "cb")

    def __init__(self, cb: Optional[Callable[[], None]] = None) -> None:
        self.cancelled = False
        self.fired = False
        self.cb = cb

    def cancel(self) -> None:
        self.cancelled = True
        self.cb = None


class _TimeoutManager:
    """One daemon thread, one heap of (deadline, seq, handle); the handle carries the callback."""

    def __init__(self) -> None:
        self._cv = threading.Condition()
        self._heap: List[Tuple[float, int, _Handle]] = []
        self._seq = itertools.count()
        self._thread: Optional[threading.Thread] = None
        self._watchdog: Optional[threading.Thread] = None
        self._last_beat = time.monotonic()
        self._shutdown = False
        # events whose streams finished: destroyed on a user thread, never on the timer thread
        self._dead_events: List[object] = []

    # -- lifecycle -------------------------------------------------------
    def _ensure_started(self) -> None:
        if self._thread is not None and self._thread.is_alive():
            return
        with self._cv:
            if self._thread is not None and self._thread.is_alive():
                return
            self._shutdown = False
            self._last_beat = time.monotonic()
            self._thread = threading.Thread(target=self._run, name="tft_timeouts", daemon=True)
            self._thread.start()
            self._watchdog = threading.Thread(target=self._watch, name="tft_timeout_watchdog", daemon=True)
            self._watchdog.start()

    def shutdown(self) -> None:
        with self._cv:
            self._shutdown = True
            self._cv.notify_all()
        t = self._thread
        if t is not None:
            t.join(timeout=2)
        self._thread = None

    # -- timer thread ----------------------------------------------------
    def _run(self) -> None:
        while True:
            with self._cv:
                if self._shutdown:
                    return
                self._last_beat = time.monotonic()
                now = time.monotonic()
                due: List[Tuple[_Handle, Callable[[], None]]] = []
                while self._heap and self._heap[0][0] <= now:
                    _, _, h = heapq.heappop(self._heap)
                    cb = h.cb
                    if not h.cancelled and cb is not None:
                        due.append((h, cb))
                if not due:
                    wait = 1.0
                    if self._heap:
                        wait = max(0.0, min(wait, self._heap[0][0] - now))
                    self._cv.wait(wait)
                    continue
            for h, cb in due:
                h.fired = True
                h.cb = None
                try:
                    cb()
                except Exception:  # pragma: no cover - callbacks must not kill the thread
                    import traceback

                    traceback.print_exc()

    def _watch(self) -> None:
        limit = float(os.environ.get(WATCHDOG_TIMEOUT_SEC_ENV, "30"))
        while True:
            time.sleep(min(limit / 2, 1.0))
            with self._cv:
                if self._shutdown:
                    return
                stuck = time.monotonic() - self._last_beat
            if stuck > limit:
                sys.stderr.write(f"torchft_b200: timeout thread stuck for {stuck:.1f}s (> {limit}s); exiting\n")
                sys.stderr.flush()
                try:
                    sys.exit(1)
                except SystemExit:
                    os._exit(1)  # sys.exit from a non-main thread only ends the thread

    # -- api ---------------------------------------------------------------
    def call_later(self, timeout: timedelta, cb: Callable[[], None]) -> _Handle:
        self._ensure_started()
        h = _Handle(cb)
        with self._cv:
            heapq.heappush(self._heap, (time.monotonic() + timeout.total_seconds(), next(self._seq), h))
            self._cv.notify_all()
        return h

    def register(self, fut: "Future[T]", timeout: timedelta) -> "Future[T]":
        """Future that mirrors ``fut`` but fails with ``TimeoutError`` after ``timeout``."""
        out: Future[T] = Future()
        lock = threading.Lock()
        done = [False]

        def finish(setter: Callable[[], None]) -> None:
            with lock:
                if done[0]:
                    return
                done[0] = True
            setter()

        handle = self.call_later(
            timeout, lambda: finish(lambda: out.set_exception(TimeoutError(f"future did not complete within {timeout}")))
        )

        def relay(f: "Future[T]") -> None:
            handle.cancel()
            try:
                v = f.value()
            except Exception as e:  # noqa: BLE001
                finish(lambda: out.set_exception(e))
            else:
                finish(lambda: out.set_result(v))

        fut.add_done_callback(relay)
        return out

    def stream_timeout(self, callback: Callable[[], None], timeout: timedelta) -> None:
        """Fire ``callback`` unless everything enqueued so far on the current stream finishes in ``timeout``."""
        self._reap_events()
        if not torch.cuda.is_available():
            return
        event = torch.cuda.Event()
        event.record()

        def check() -> None:
            try:
                finished = event.query()
            except Exception:  # context already broken: treat as not finished
                finished = False
            if not finished:
                callback()
            with self._cv:
                self._dead_events.append(event)

        self.call_later(timeout, check)

    def _reap_events(self) -> None:
        with self._cv:
            dead, self._dead_events = self._dead_events, []
        del dead  # CUDA events released here, on a user thread

    @contextmanager
    def context_timeout(self, callback: Callable[[], None], timeout: timedelta) -> Generator[None, None, None]:
        h = self.call_later(timeout, callback)
        try:
            yield
        finally:
            h.cancel()


_TIMEOUT_MANAGER = _TimeoutManager()


def future_timeout(fut: "Future[T]", timeout: timedelta) -> "Future[T]":
    """Return a future that completes like ``fut`` or raises ``TimeoutError`` after ``timeout``."""
    return _TIMEOUT_MANAGER.register(fut, timeout)


def future_wait(fut: "Future[T]", timeout: timedelta) -> T:
    """Block for ``fut`` at most ``timeout``; raises ``TimeoutError`` (the future itself is untouched)."""

    ev = threading.Event()
    fut.add_done_callback(lambda _f: ev.set())
    if not ev.wait(timeout.total_seconds()):
        raise TimeoutError(f"future did not complete within {timeout}")
    return fut.wait()


def stream_timeout(callback: Callable[[], None], timeout: timedelta) -> None:
    """Call ``callback`` if the work currently enqueued on this CUDA stream has not finished after ``timeout``."""
    _TIMEOUT_MANAGER.stream_timeout(callback, timeout)


@contextmanager
def context_timeout(callback: Callable[[], None], timeout: timedelta) -> Generator[None, None, None]:
    """Run the body; if it is still running after ``timeout`` call ``callback`` (e.g. ``pg.abort``)."""
    with _TIMEOUT_MANAGER.context_timeout(callback, timeout):
        yield
