"""High-level fault-tolerant trainer: the "one call a user makes" per step.

Wires together, B200-first:

    ProcessGroupB200 (NVLink peer-memory collectives, remap-on-quorum)
      -> Manager (quorum / heal / commit protocol; C++ control plane)
      -> FlatParams (one flat bf16 parameter buffer, one flat SYMMETRIC gradient buffer)
      -> FlatDistributedDataParallel (per-bucket fused all-reduce overlapped with backward)
      -> OptimizerWrapper(FlatAdamW) (single-launch AdamW, stepped only on commit)

This is the HSDP configuration of BASELINE.json with shard degree 1: every GPU is
one replica group holding the full model (180 GB of HBM3e fits Llama-3-8B weights,
gradients, fp32 master weights and Adam state), and fault tolerance lives on the
replicated dimension exactly as in the reference (README.md:37-42).
"""

from __future__ import annotations

import dataclasses
import os
from datetime import timedelta
from typing import Any, Dict, Optional

import torch
from torch import nn
from torch.distributed import TCPStore

from torchft_b200.ddp import FlatDistributedDataParallel
from torchft_b200.manager import Manager
from torchft_b200.models.llama import CONFIGS, FlatParams, Llama, LlamaConfig
from torchft_b200.optim import OptimizerWrapper
from torchft_b200.ops.fused import FlatAdamW


class FaultTolerantTrainer:
    """Llama trainer over one replica group per GPU.

    Args:
        model: config name in ``models.llama.CONFIGS`` or a ``LlamaConfig``
        lighthouse_addr: address of a running Lighthouse
        replica_id: this replica group's name (``"replica_3"``); trailing digits give its global rank
        min_replica_size: minimum participating replicas for a step to commit
        backend: ``"b200"`` (native peer-memory kernels) or ``"nccl"`` (reference-equivalent baseline)
        bucket_mb: gradient bucket size for the overlapped all-reduce
        should_quantize: fused fp8 gradient all-reduce
        overlap_optimizer: run AdamW layer by layer on a side stream so the next forward starts as soon
            as its first layers are updated (default: env ``TORCHFT_B200_OVERLAP_OPT``, else on)
        optimizer_blocks: CTA cap of the overlapped AdamW launches (0 = kernel default)
    """

    def __init__(self, model: str | LlamaConfig, lighthouse_addr: str, replica_id: str = "replica_0",
                 min_replica_size: int = 1, backend: str = "b200", bucket_mb: float = 512.0,
                 should_quantize: bool = False, lr: float = 3e-4, seed: int = 0,
                 timeout: timedelta = timedelta(seconds=60), device: Optional[torch.device] = None,
                 activation_checkpoint: Optional[str] = None, init_sync: bool = False,
                 use_async_quorum: bool = True, overlap_optimizer: Optional[bool] = None,
                 optimizer_blocks: int = 0) -> None:
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        torch.cuda.set_device(self.device)
        cfg = CONFIGS[model] if isinstance(model, str) else model
        if activation_checkpoint is not None:
            cfg = dataclasses.replace(cfg, activation_checkpoint=activation_checkpoint)
        self.cfg = cfg
        self.backend = backend

        # the replica group's own store (group world size 1 => this process hosts it)
        self._store = TCPStore("127.0.0.1", 0, is_master=True, wait_for_workers=False)

        if backend == "b200":
            from torchft_b200.parallel.process_group_b200 import ProcessGroupB200

            self.pg: Any = ProcessGroupB200(timeout=timeout, device=self.device)
        elif backend == "nccl":
            from torchft_b200.process_group import ProcessGroupNCCL

            self.pg = ProcessGroupNCCL(timeout=timeout)
        else:
            raise ValueError(f"unknown backend {backend}")

        # build on meta, materialise directly into the flat buffers, then initialise in place
        self.model = Llama(cfg, device="meta")
        numel_bytes = sum((p.numel() + FlatParams.ALIGN - 1) // FlatParams.ALIGN * FlatParams.ALIGN
                          for p in self.model.parameters()) * 2
        if backend == "b200":
            grad_alloc = lambda n: self.pg.alloc_symmetric("grads", numel_bytes).view(torch.bfloat16)  # noqa: E731
        else:
            grad_alloc = None
        self.flat = FlatParams(self.model, grad_alloc=grad_alloc, device=self.device)
        self.model.init_weights(seed)
        self.inner_optim = FlatAdamW(self.flat.param, self.flat.grad, lr=lr)
        self.inner_optim.direct_grads = True  # step_device() drops p.grad; wgrad GEMMs fill the flat buffer

        self.manager = Manager(
            pg=self.pg,
            load_state_dict=self.load_state_dict,
            state_dict=self.state_dict,
            min_replica_size=min_replica_size,
            use_async_quorum=use_async_quorum,
            timeout=timeout,
            quorum_timeout=timeout,
            connect_timeout=timeout,
            rank=0,
            world_size=1,
            store_addr="127.0.0.1",
            store_port=self._store.port,
            lighthouse_addr=lighthouse_addr,
            replica_id=replica_id,
            init_sync=init_sync,
        )
        self.ddp = FlatDistributedDataParallel(self.manager, self.model, self.flat, bucket_mb=bucket_mb,
                                               should_quantize=should_quantize)
        self.optim = OptimizerWrapper(self.manager, self.inner_optim)
        if overlap_optimizer is None:
            overlap_optimizer = os.environ.get("TORCHFT_B200_OVERLAP_OPT", "1") != "0"
        self._opt_blocks = int(os.environ.get("TORCHFT_B200_OPT_BLOCKS", optimizer_blocks))
        self._opt_stream: Optional[torch.cuda.Stream] = None
        if overlap_optimizer:
            self._setup_optimizer_overlap()
        self._tok: Optional[torch.Tensor] = None
        self._tgt: Optional[torch.Tensor] = None

    # ------------------------------------------------- optimizer / forward overlap
    def _setup_optimizer_overlap(self) -> None:
        """Cut the flat AdamW update into one range per top-level module, in FORWARD order.

        The update is HBM-bound and the forward is tensor-core-bound, so instead of idling the tensor
        cores for the whole update (12 % of a Llama-3-8B step) the ranges run on a side stream and each
        module's forward only waits for the event of the range(s) holding its own parameters.
        """
        where = {id(p): (o, p.numel()) for p, o in zip(self.flat.params, self.flat.offsets)}
        ranges = []
        for stage in self.model.param_stages():
            mine = [where[id(p)] for p in stage]
            lo = min(o for o, _ in mine)
            hi = max((o + n + FlatParams.ALIGN - 1) // FlatParams.ALIGN * FlatParams.ALIGN for o, n in mine)
            ranges.append((lo, min(hi, self.flat.numel)))
        covered = sorted(ranges)
        assert covered[0][0] == 0 and covered[-1][1] == self.flat.numel and all(
            a[1] == b[0] for a, b in zip(covered, covered[1:])), "parameter stages must tile the flat buffer"
        self._opt_ranges = ranges
        self._opt_events = [torch.cuda.Event() for _ in ranges]
        self._opt_stream = torch.cuda.Stream(device=self.device)
        self._opt_pending = False

        def gate(stage: int) -> None:
            if self._opt_pending:
                torch.cuda.current_stream().wait_event(self._opt_events[stage])

        self.model.stage_hook = gate

    def _optimizer_step(self) -> None:
        if self._opt_stream is None:
            self.inner_optim.step()
            return
        self._opt_stream.wait_stream(torch.cuda.current_stream())
        self.inner_optim.step_ranges(self._opt_ranges, (self._opt_stream, self._opt_events), max_blocks=self._opt_blocks)
        self._opt_pending = True

    def _join_optimizer(self) -> None:
        """Make the current stream (and later host readers) see a fully applied update."""
        if self._opt_stream is not None and self._opt_pending:
            torch.cuda.current_stream().wait_stream(self._opt_stream)

    # -------------------------------------------------------------- heal hooks
    def state_dict(self) -> Dict[str, Any]:
        if self._opt_stream is not None:
            self._opt_stream.synchronize()  # heal senders read on their own stream: hand them a finished update
        o = self.inner_optim
        return {"param": self.flat.param, "master": o.master, "m": o.m, "v": o.v, "t": o.t}

    def load_state_dict(self, sd: Dict[str, Any]) -> None:
        o = self.inner_optim
        self._join_optimizer()
        with torch.no_grad():
            for name, dst in (("param", self.flat.param), ("master", o.master), ("m", o.m), ("v", o.v)):
                if sd[name].data_ptr() != dst.data_ptr():
                    dst.copy_(sd[name])
        o.t = int(sd["t"])

    # -------------------------------------------------------------------- step
    def step_device(self, tokens: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
        """One fault-tolerant optimisation step on device-resident inputs; returns the loss tensor (device)."""
        self.optim.zero_grad(set_to_none=True)  # start_quorum (async); no memset:
        self.flat.reset_grads()                  # wgrad GEMMs write the flat gradient buffer directly
        loss = self.ddp(tokens, targets)  # forward (fused kernels + cuBLAS + SDPA)
        loss.backward()                   # bucket all-reduces launch from grad hooks, overlapped
        self.ddp.finish()                 # current stream waits for the comm stream
        if self.manager.should_commit():  # OptimizerWrapper.step() semantics (optim.py), with the update
            self._optimizer_step()        # cut per layer so the next forward overlaps it
        return loss

    def step(self, tokens_cpu: torch.Tensor, targets_cpu: torch.Tensor) -> float:
        """End-to-end step: pinned-host inputs -> H2D -> train step -> D2H loss."""
        B, S = tokens_cpu.shape
        if self._tok is None or self._tok.shape != tokens_cpu.shape:
            self._tok = torch.empty((B, S), dtype=torch.int64, device=self.device)
            self._tgt = torch.empty((B, S), dtype=torch.int64, device=self.device)
        assert self._tgt is not None
        self._tok.copy_(tokens_cpu, non_blocking=True)
        self._tgt.copy_(targets_cpu, non_blocking=True)
        loss = self.step_device(self._tok, self._tgt)
        return float(loss.item())

    def shutdown(self) -> None:
        if self._opt_stream is not None:
            self._opt_stream.synchronize()
        self.manager.shutdown(wait=False)
        self.pg.shutdown()
