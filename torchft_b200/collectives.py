"""Quantised (fp8) collectives over ANY reconfigurable process group.

API parity with /root/reference/torchft/collectives.py (``allreduce_quantized``,
``reduce_scatter_quantized``, ``get_padded_sizes``, ``allocate_reduce_scatter_output``).
This is the *generic* path (used when the group is NCCL/Gloo-backed): our sm_100a
quantise / reduce / dequantise kernels around the group's ``alltoall_base`` and
``allgather``. On :class:`ProcessGroupB200` the whole pipeline is a single fused
kernel instead -- these functions detect that and delegate.
"""

from __future__ import annotations

from datetime import timedelta
from typing import Any, List, Optional, Tuple

import torch
import torch.distributed as dist
from torch.distributed import ReduceOp, Work
from torch.distributed.distributed_c10d import AllgatherOptions, AllreduceOptions, AllToAllOptions, ReduceScatterOptions
from torch.futures import Future

from torchft_b200 import quantization as Q
from torchft_b200.ops import _native


def _op_of(opts: Any) -> ReduceOp:
    return opts.reduceOp if isinstance(opts, (AllreduceOptions, ReduceScatterOptions)) else opts


class _StreamWork(Work):
    """Completed-on-stream work: ``wait`` joins ``stream`` into the caller's current stream."""

    def __init__(self, stream: Optional[torch.cuda.Stream], result: Any) -> None:
        super().__init__()
        self._event = stream.record_event() if stream is not None else None
        self._fut: Future = Future()
        self._fut.set_result(result)

    def wait(self, timeout: Optional[timedelta] = None) -> bool:
        if self._event is not None:
            torch.cuda.current_stream().wait_event(self._event)
        return True

    def get_future(self) -> Future:
        return self._fut


def get_padded_sizes(tensors: List[torch.Tensor], world_size: int) -> List[torch.Size]:
    """Shapes with dim 0 rounded up to a multiple of ``world_size`` (reference: collectives.py:51-75)."""
    out = []
    for t in tensors:
        s = list(t.shape) if t.dim() > 0 else [1]
        s[0] = (s[0] + world_size - 1) // world_size * world_size
        out.append(torch.Size(s))
    return out


def allocate_reduce_scatter_output(tensors: List[torch.Tensor], world_size: int) -> Tuple[torch.Tensor, List[torch.Size]]:
    """One flat output buffer holding this rank's 1/world_size row-slice of every (padded) tensor."""
    padded = get_padded_sizes(tensors, world_size)
    dt, dev = tensors[0].dtype, tensors[0].device
    for t in tensors:
        if t.dtype != dt or t.device != dev:
            raise ValueError("all tensors must share dtype and device")
    chunk = sum(s.numel() // world_size for s in padded)
    return torch.zeros(chunk, dtype=dt, device=dev), padded


def _check(tensors: List[torch.Tensor], op: ReduceOp) -> None:
    if op not in (ReduceOp.SUM, ReduceOp.AVG):
        raise NotImplementedError(f"quantized collectives support SUM and AVG only, got {op}")
    if not tensors or not all(t.is_cuda for t in tensors):
        raise ValueError("quantized collectives need CUDA tensors")


def allreduce_quantized(tensors: List[torch.Tensor], opts: AllreduceOptions | ReduceOp, process_group: dist.ProcessGroup,
                        sync_stream: Optional[torch.cuda.Stream] = None) -> Work:
    """In-place fp8 all-reduce of ``tensors`` (SUM or AVG).

    quantize -> all-to-all of row slices -> fp32 reduce + requantize of the local slice ->
    all-gather -> dequantize, all enqueued on ``sync_stream`` (a side stream by default).
    Expected mean relative error <= 0.04 (reference tolerance, collectives_test.py:186).
    """
    op = _op_of(opts)
    _check(tensors, op)
    world, rank = process_group.size(), process_group.rank() if hasattr(process_group, "rank") else dist.get_rank(process_group)
    fused = getattr(process_group, "allreduce_q8", None)
    if fused is not None:  # ProcessGroupB200: one kernel per tensor, no NCCL
        scale = 1.0 / world if op == ReduceOp.AVG else 1.0
        works = [fused(t, t, None, scale=scale) for t in tensors]
        for w in works[:-1]:
            w.wait()
        return works[-1]
    stream = sync_stream if sync_stream is not None else torch.cuda.Stream()
    stream.wait_stream(torch.cuda.current_stream())
    contig = [t if t.is_contiguous() else t.contiguous() for t in tensors]
    with torch.cuda.stream(stream):
        qbuf = Q.fused_quantize_into_fp8(contig, world)
        recv = torch.empty((world, qbuf.numel()), dtype=torch.uint8, device=qbuf.device)
        # Q8G keeps each rank's groups contiguous per tensor but tensors are laid out back to
        # back, so ship the whole buffer to every peer (all-gather) and let each rank reduce
        # only its own group range: same bytes on the wire as all-to-all + all-gather of slices
        # for world <= 2 and simpler bookkeeping; for larger worlds prefer ProcessGroupB200.
        ag = AllgatherOptions()
        process_group.allgather([list(recv.unbind(0))], [qbuf], ag).wait()
        bufs = list(recv.unbind(0))
        Q.fused_reduce_fp8(contig, bufs, world, rank, op)
        # gather every rank's reduced slice
        red = torch.empty_like(recv)
        process_group.allgather([list(red.unbind(0))], [bufs[rank]], ag).wait()
        final = Q.merge_reduced_slices(contig, list(red.unbind(0)), world)
        Q.fused_dequantize_from_fp8(contig, final, world)
        for t, c in zip(tensors, contig):
            if t.data_ptr() != c.data_ptr():
                t.copy_(c)
        for t in tensors:
            t.record_stream(stream)
    return _StreamWork(stream, tensors)


def reduce_scatter_quantized(output: torch.Tensor, inputs: List[torch.Tensor], opts: ReduceScatterOptions | ReduceOp,
                             process_group: dist.ProcessGroup, sync_stream: Optional[torch.cuda.Stream] = None) -> Work:
    """fp8 reduce-scatter: ``output`` receives this rank's row-slice of every (row-padded) input,
    concatenated (layout of :func:`allocate_reduce_scatter_output`)."""
    op = _op_of(opts)
    _check(inputs, op)
    world = process_group.size()
    rank = process_group.rank() if hasattr(process_group, "rank") else dist.get_rank(process_group)
    stream = sync_stream if sync_stream is not None else torch.cuda.Stream()
    stream.wait_stream(torch.cuda.current_stream())
    padded = get_padded_sizes(inputs, world)
    with torch.cuda.stream(stream):
        # pad rows so every rank owns an equal slice, reduce in fp8, keep only our slice
        work_tensors = []
        for t, ps in zip(inputs, padded):
            p = torch.zeros(ps, dtype=t.dtype, device=t.device)
            p.view(ps[0], -1)[: (t.shape[0] if t.dim() else 1)].copy_(t.reshape(t.shape[0] if t.dim() else 1, -1))
            work_tensors.append(p)
        allreduce_quantized(work_tensors, op, process_group, stream).wait()
        off = 0
        for p in work_tensors:
            rows = p.shape[0] // world
            sl = p[rank * rows : (rank + 1) * rows].reshape(-1)
            output[off : off + sl.numel()].copy_(sl)
            off += sl.numel()
        output.record_stream(stream)
    return _StreamWork(stream, output)
