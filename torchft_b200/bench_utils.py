"""Small helpers shared by ``bench.py``, ``__graft_entry__.smoke`` and the GPU tests."""

from __future__ import annotations

from datetime import timedelta
from typing import Any, Optional

import torch


def local_lighthouse(min_replicas: int = 1, join_timeout_ms: int = 100) -> Any:
    from torchft_b200.coordination import LighthouseServer

    return LighthouseServer(bind="[::]:0", min_replicas=min_replicas, join_timeout_ms=join_timeout_ms)


def loopback(addr: str) -> str:
    """``http://<hostname>:port`` -> ``http://127.0.0.1:port`` (container hostnames may not resolve)."""

    host = addr.split("//")[1].rsplit(":", 1)[0]
    return addr.replace(host, "127.0.0.1")


def ft_smoke_step(model: Optional[Any] = None, cfg: Optional[Any] = None, steps: int = 2) -> float:
    """A tiny but COMPLETE fault-tolerant training step on cuda:0: Lighthouse + Manager (C++ control
    plane) + ProcessGroupB200 + fused-kernel Llama forward/backward + gated AdamW, through the
    public trainer API including the pinned H2D input copy and the D2H loss read."""
    from torchft_b200.parallel.trainer import FaultTolerantTrainer

    lh = local_lighthouse()
    trainer = None
    try:
        trainer = FaultTolerantTrainer("llama3_debug", loopback(lh.address()), replica_id="smoke_0",
                                       timeout=timedelta(seconds=30), bucket_mb=1.0)
        c = trainer.cfg
        tok = torch.randint(0, c.vocab_size, (2, 128)).pin_memory()
        tgt = torch.randint(0, c.vocab_size, (2, 128)).pin_memory()
        loss = 0.0
        for _ in range(steps):
            loss = trainer.step(tok, tgt)
        assert trainer.manager.current_step() == steps, "steps did not commit"
        return loss
    finally:
        if trainer is not None:
            trainer.shutdown()
        lh.shutdown()
