"""Low-level coordination API: the C++ control-plane classes.

Use these directly to build custom fault-tolerance algorithms on top of the
Lighthouse / Manager protocol; most users want :class:`torchft_b200.Manager`.
Parity: /root/reference/torchft/coordination.py:23-39 (re-export of the pyo3 module).

* ``LighthouseServer`` / ``LighthouseClient`` -- global quorum authority and its client.
* ``ManagerServer`` / ``ManagerClient``       -- per-replica-group barrier + quorum proxy.
* ``Quorum``, ``QuorumMember``, ``QuorumResult``, ``Timestamp`` -- message types.
* ``quorum_compute`` / ``compute_quorum_results`` -- the pure decision procedures.
"""

from torchft_b200._C import (  # noqa: F401
    LighthouseClient,
    LighthouseServer,
    ManagerClient,
    ManagerServer,
    Quorum,
    QuorumMember,
    QuorumResult,
    Timestamp,
    compute_quorum_results,
    lighthouse_main,
    quorum_compute,
)

__all__ = [
    "LighthouseClient",
    "LighthouseServer",
    "ManagerClient",
    "ManagerServer",
    "Quorum",
    "QuorumMember",
    "QuorumResult",
    "Timestamp",
    "compute_quorum_results",
    "quorum_compute",
    "lighthouse_main",
]
