"""Abstract live-checkpoint transport used to heal a recovering replica.

Same contract as the reference's ``CheckpointTransport``
(/root/reference/torchft/checkpointing/transport.py:14-68): the up-to-date
replica calls :meth:`send_checkpoint` for the ranks assigned to it by the
quorum, the healing replica calls :meth:`recv_checkpoint` with the source's
:meth:`metadata` string, and the source stops serving in
:meth:`disallow_checkpoint` (called from ``Manager.should_commit``) before its
optimizer mutates the state.
"""

from __future__ import annotations

from abc import ABC, abstractmethod
from datetime import timedelta
from typing import Generic, List, TypeVar

T = TypeVar("T")


class CheckpointTransport(Generic[T], ABC):
    """Interface of a live-heal transport (reference: checkpointing/transport.py:14-68): the source calls
    ``send_checkpoint`` / ``disallow_checkpoint``, the healing replica ``recv_checkpoint`` with the source's ``metadata()``."""

    @abstractmethod
    def metadata(self) -> str:
        """Opaque string a remote transport needs to fetch from this one (e.g. a URL)."""

    @abstractmethod
    def send_checkpoint(self, dst_ranks: List[int], step: int, state_dict: T, timeout: timedelta) -> None:
        """Make ``state_dict`` for ``step`` available to ``dst_ranks`` (may be asynchronous)."""

    def disallow_checkpoint(self) -> None:
        """Block until in-flight transfers finish; afterwards the state may be mutated."""

    @abstractmethod
    def recv_checkpoint(self, src_rank: int, metadata: str, step: int, timeout: timedelta) -> T:
        """Fetch the checkpoint for ``step`` from ``src_rank``."""

    def shutdown(self, wait: bool = True) -> None:
        """Release sockets/threads."""
