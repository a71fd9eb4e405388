"""Trivial ``Work`` objects (reference: torchft/work.py:15-26)."""

from __future__ import annotations

from datetime import timedelta
from typing import Optional

import torch
from torch.distributed import Work


class DummyWork(Work):
    """Already-completed work whose future resolves to ``result`` (used after an error is latched)."""

    def __init__(self, result: object) -> None:
        super().__init__()
        self._result = result
        self._fut: torch.futures.Future[object] = torch.futures.Future()
        self._fut.set_result(result)

    def wait(self, timeout: Optional[timedelta] = None) -> bool:
        return True

    def get_future(self) -> torch.futures.Future[object]:
        return self._fut

    def is_completed(self) -> bool:  # pragma: no cover - trivial
        return True


_DummyWork = DummyWork  # reference-compatible alias
