"""Ownership map for state that is *partitioned across replica groups* yet must survive their loss.

Groundwork for the fault-tolerant partitioned optimizer of ROADMAP.md §1 (not wired into the trainer yet):
optimizer state is cut into ``V`` virtual shards and every shard lives on ``k`` replicas. Ownership is a PURE
function of (shard, set of live replica ids) — rendezvous / highest-random-weight hashing — so every replica derives
the same map from the quorum alone, with no extra coordination round, and membership changes move the provable
minimum of data:

* a replica that leaves only causes the shards it owned to gain one new owner each;
* a replica that joins only *takes over* shards (existing owners never trade shards among themselves);
* a shard is lost only if all ``k`` of its owners disappear in the same quorum transition.

``plan_transition`` turns two consecutive quorums into the copy list ``(shard, source replica, destination replica)``
that the heal-copy kernel would execute, plus the shards that have to be re-initialised (or restored from a durable
checkpoint) because nobody who held them survived.
"""

from __future__ import annotations

import hashlib
from dataclasses import dataclass
from typing import Dict, List, Sequence, Tuple

__all__ = ["ShardMap", "Transition"]


def _weight(shard: int, replica_id: str) -> int:
    h = hashlib.blake2b(f"{shard}\x00{replica_id}".encode(), digest_size=8).digest()
    return int.from_bytes(h, "big")


@dataclass(frozen=True)
class Transition:
    """Result of :meth:`ShardMap.plan_transition`: copies to perform, shards nobody survived with, untouched pairs."""

    copies: Tuple[Tuple[int, str, str], ...]  # (shard, source replica id, destination replica id)
    lost: Tuple[int, ...]                     # shards none of whose previous owners is still alive
    kept: int                                 # (shard, owner) pairs that did not move


class ShardMap:
    """``num_shards`` virtual shards, each held by ``replication`` replicas chosen by rendezvous hashing over replica ids."""

    def __init__(self, num_shards: int = 256, replication: int = 2) -> None:
        if num_shards < 1 or replication < 1:
            raise ValueError("num_shards and replication must be >= 1")
        self.num_shards = num_shards
        self.replication = replication

    def owners(self, shard: int, members: Sequence[str]) -> Tuple[str, ...]:
        """The ``min(k, len(members))`` owners of ``shard``, primary first. Order of ``members`` is irrelevant."""
        ranked = sorted(set(members), key=lambda rid: (-_weight(shard, rid), rid))
        return tuple(ranked[: self.replication])

    def assignment(self, members: Sequence[str]) -> Dict[str, List[int]]:
        """replica id -> shards it holds (as primary or backup), for load inspection and allocation."""
        out: Dict[str, List[int]] = {rid: [] for rid in set(members)}
        for s in range(self.num_shards):
            for rid in self.owners(s, members):
                out[rid].append(s)
        return out

    def primaries(self, members: Sequence[str]) -> Dict[str, List[int]]:
        """replica id -> shards it is PRIMARY for (the replica that runs the update and publishes the result)."""
        out: Dict[str, List[int]] = {rid: [] for rid in set(members)}
        if not out:
            return out
        for s in range(self.num_shards):
            out[self.owners(s, members)[0]].append(s)
        return out

    def plan_transition(self, old_members: Sequence[str], new_members: Sequence[str]) -> Transition:
        """Copies needed so that every owner under ``new_members`` holds its shards, given who held them before."""
        alive = set(new_members)
        copies: List[Tuple[int, str, str]] = []
        lost: List[int] = []
        kept = 0
        for s in range(self.num_shards):
            before = self.owners(s, old_members) if old_members else ()
            after = self.owners(s, new_members)
            sources = [rid for rid in before if rid in alive]  # survivors that hold the shard, best-ranked first
            for dst in after:
                if dst in before:
                    kept += 1
                elif sources:
                    # spread reads over the surviving holders deterministically
                    copies.append((s, sources[_weight(s, dst) % len(sources)], dst))
                elif before:
                    if s not in lost:
                        lost.append(s)
                # no previous quorum at all: everybody initialises its shards locally, nothing to copy
        return Transition(tuple(copies), tuple(lost), kept)
