"""World-size-N correctness of the peer-memory collectives on FEWER than N GPUs.

Ranks are mapped round-robin onto the visible GPUs (several processes per GPU; CUDA IPC works
between processes on one device and the GPU time-slices their spinning kernels), so the W=8
code paths — slice arithmetic, per-CTA barriers over 8 flag slots, the U=2 two-shot variant,
the fp8 exchange — can be validated on a 1/2/4-GPU box. Rendezvous is gloo only (NCCL refuses
two ranks on one device). Timings from this script are meaningless; it only checks results.

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 bench/comm_oversub.py
"""

from __future__ import annotations

import json
import os
import sys
from datetime import timedelta

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from torchft_b200.parallel.symm_mem import SymmetricComm  # noqa: E402


def main() -> None:
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    ngpu = torch.cuda.device_count()
    dev = torch.device("cuda", rank % ngpu)
    torch.cuda.set_device(dev)
    dist.init_process_group("gloo", timeout=timedelta(seconds=300))
    store = dist.distributed_c10d._get_default_store()
    comm = SymmetricComm(timeout=timedelta(seconds=60))
    symm = comm.alloc("t", 80 << 20)
    comm.configure(dist.PrefixStore("q1", store), rank, world, 1)
    fails = []
    cases = 0
    gen = torch.Generator(device="cuda").manual_seed(77 + rank)

    def ref_mean(x: torch.Tensor) -> torch.Tensor:
        r = x.float().cpu()
        dist.all_reduce(r)
        return (r / world).to(dev)

    for dtype, tol in ((torch.float32, 1e-5), (torch.bfloat16, 2e-2)):
        es = torch.empty(0, dtype=dtype).element_size()
        for n in (1, 9, 4096, 100_003, 1 << 20, (16 << 20) + 5):
            x = torch.randn(n, device=dev, generator=gen).to(dtype)
            ref = ref_mean(x)
            for path in ("symm", "staged"):
                for plan in (None, "oneshot", "twoshot"):
                    if plan == "oneshot" and n * es > (1 << 20):
                        continue
                    if path == "symm":
                        y = symm[: n * es].view(dtype)
                        y.copy_(x)
                    else:
                        y = x.clone()
                    comm._force_plan = None if plan is None else (0 if plan == "oneshot" else 1, 16)
                    comm.allreduce_(y, scale=1.0 / world)
                    torch.cuda.synchronize()
                    comm._force_plan = None
                    err = (y.float() - ref).abs().max().item()
                    bad = comm.errored()
                    cases += 1
                    if bad is not None or not err <= tol * max(1.0, ref.abs().max().item()):
                        fails.append({"dtype": str(dtype), "n": n, "path": path, "plan": plan, "err": err, "latched": str(bad)})
    # non-participant + MAX
    x = torch.full((70_000,), float(rank + 1), device=dev)
    comm.allreduce_(x, scale=1.0, contribute=(rank != 0))
    torch.cuda.synchronize()
    cases += 1
    if not bool((x == sum(r + 1 for r in range(1, world))).all()):
        fails.append({"case": "non_participant"})
    x = torch.full((5000,), float(rank), device=dev)
    comm.allreduce_(x, op=1)
    torch.cuda.synchronize()
    cases += 1
    if not bool((x == world - 1).all()):
        fails.append({"case": "max"})
    # fused fp8 all-reduce of a delta
    for n in (512, 100_000, (2 << 20) + 17):
        a = torch.randn(n, device=dev, generator=gen) * 3
        b = torch.randn(n, device=dev, generator=gen)
        ref = ref_mean(a - b)
        out = torch.empty_like(a)
        comm.q8_allreduce_(out, a, b, scale=1.0 / world)
        torch.cuda.synchronize()
        rel = ((out - ref).abs().mean() / ref.abs().mean()).item()
        cases += 1
        if comm.errored() is not None or not rel <= 0.04:
            fails.append({"case": "q8", "n": n, "rel": rel, "latched": str(comm.errored())})
    # reconfigure to a new epoch with the same members, then once more without the last rank
    comm.configure(dist.PrefixStore("q2", store), rank, world, 2)
    x = torch.ones(1 << 16, device=dev)
    comm.allreduce_(x, scale=1.0)
    torch.cuda.synchronize()
    cases += 1
    if not bool((x == world).all()):
        fails.append({"case": "reconfigure"})
    if world > 2:
        dist.barrier()
        if rank < world - 1:
            comm.configure(dist.PrefixStore("q3", store), rank, world - 1, 3)
            x = torch.ones(1 << 16, device=dev)
            comm.allreduce_(x, scale=1.0)
            torch.cuda.synchronize()
            cases += 1
            if not bool((x == world - 1).all()) or comm.errored() is not None:
                fails.append({"case": "shrink", "latched": str(comm.errored())})
    nf = torch.tensor([len(fails)])
    dist.all_reduce(nf)
    if fails:
        print(f"[rank {rank}] FAIL " + json.dumps(fails[:6]), flush=True)
    if rank == 0:
        print("COMM_OVERSUB " + json.dumps({"world": world, "gpus": ngpu, "cases_rank0": cases, "failures_all_ranks": int(nf.item())}), flush=True)
    dist.barrier()
    comm.shutdown()
    dist.destroy_process_group()
    if int(nf.item()):
        sys.exit(1)


if __name__ == "__main__":
    main()
