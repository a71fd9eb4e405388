"""Low-level coordination API: the C++ control plane, importable on its own.

Build custom fault-tolerance algorithms directly on the Lighthouse / Manager quorum protocol with these
(``LighthouseServer``/``LighthouseClient``: global quorum authority; ``ManagerServer``/``ManagerClient``:
per-replica-group barrier and quorum proxy; ``quorum_compute``/``compute_quorum_results``: the pure
decision procedures);
most users want :class:`torchft_b200.Manager` instead. Same surface as the reference's
``torchft/coordination.py:23-39`` (a re-export of its native module) plus two conveniences that
only make sense for our wire protocol: :func:`lighthouse_status` and :func:`wait_for_lighthouse`.
"""

from __future__ import annotations

import json
import time
import urllib.error
import urllib.request
from datetime import timedelta
from typing import Any, Dict

from torchft_b200 import _C as _native

# name -> what it is (also used to fill in a docstring when the binding has none)
_SURFACE: Dict[str, str] = {
    "LighthouseServer": "Global quorum authority (one per job): heartbeats, quorum computation, dashboard.",
    "LighthouseClient": "Client of a LighthouseServer: quorum() and heartbeat().",
    "ManagerServer": "Per-replica-group server hosted by group rank 0: group barrier for quorum and should_commit.",
    "ManagerClient": "Client every rank of a replica group uses to talk to its ManagerServer.",
    "Quorum": "A formed quorum: id, participants, creation time.",
    "QuorumMember": "One replica group inside a Quorum.",
    "QuorumResult": "A rank's view of a quorum: replica rank/world size, recovery source and destinations, store.",
    "Timestamp": "Seconds + nanos creation time of a Quorum.",
    "quorum_compute": "Pure decision procedure of the Lighthouse: is there a valid quorum now, and who is in it.",
    "compute_quorum_results": "Pure decision procedure of the ManagerServer: per-rank recovery assignment.",
    "lighthouse_main": "Entry point of the `torchft_b200_lighthouse` command line.",
}

for _name, _what in _SURFACE.items():
    _obj = getattr(_native, _name)
    if not getattr(_obj, "__doc__", None):
        try:
            _obj.__doc__ = _what
        except (AttributeError, TypeError):  # builtin function objects are read-only
            pass
    globals()[_name] = _obj
del _name, _what, _obj

__all__ = [*_SURFACE, "lighthouse_status", "wait_for_lighthouse"]


def _http_base(addr: str) -> str:
    return (addr if "://" in addr else f"http://{addr}").rstrip("/")


def lighthouse_status(addr: str, timeout: timedelta = timedelta(seconds=5)) -> Dict[str, Any]:
    """Machine-readable dashboard of a Lighthouse (``GET /status.json``): quorum id, previous quorum with
    per-member step / recovering flag, replicas currently waiting, heartbeat ages."""
    with urllib.request.urlopen(_http_base(addr) + "/status.json", timeout=timeout.total_seconds()) as r:
        return json.loads(r.read().decode())


def wait_for_lighthouse(addr: str, timeout: timedelta = timedelta(seconds=60), poll_s: float = 0.2) -> Dict[str, Any]:
    """Block until the Lighthouse at ``addr`` answers (launch scripts start it next to the trainers);
    returns its first status. Raises ``TimeoutError``."""

    deadline = time.monotonic() + timeout.total_seconds()
    last: Exception = TimeoutError("not tried")
    while time.monotonic() < deadline:
        try:
            return lighthouse_status(addr, timedelta(seconds=min(2.0, max(0.1, deadline - time.monotonic()))))
        except (urllib.error.URLError, OSError, ValueError) as e:
            last = e
            time.sleep(poll_s)
    raise TimeoutError(f"lighthouse at {addr} did not answer within {timeout}: {last}")
