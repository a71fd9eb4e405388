"""2-GPU worker (launched by tests/test_ft_gpu.py through torchrun): quantized collectives over every
reconfigurable CUDA backend vs exact results (strategy of the reference's collectives_test.py:45-211),
plus a Baby (subprocess) NCCL all-reduce."""

from __future__ import annotations

import json
import os
import sys
from datetime import timedelta

import torch
import torch.distributed as dist
from torch.distributed import ReduceOp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from torchft_b200.collectives import allocate_reduce_scatter_output, allreduce_quantized, reduce_scatter_quantized  # noqa: E402
from torchft_b200.process_group import ProcessGroupNCCL  # noqa: E402


def main() -> None:
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("gloo", timeout=timedelta(seconds=120))
    store = dist.distributed_c10d._get_default_store()
    store_addr = f"{os.environ['MASTER_ADDR']}:{os.environ['MASTER_PORT']}"
    fails = []

    def exact(t: torch.Tensor, avg: bool) -> torch.Tensor:
        r = t.float().cpu()
        dist.all_reduce(r)
        return (r / world if avg else r).to(dev)

    from torchft_b200.baby import ProcessGroupBabyNCCL
    from torchft_b200.parallel.process_group_b200 import ProcessGroupB200

    backends = [("nccl", ProcessGroupNCCL(timeout=timedelta(seconds=60))),
                ("b200", ProcessGroupB200(timeout=timedelta(seconds=60)))]
    for qi, (name, pg) in enumerate(backends):
        pg.configure(f"{store_addr}/coll/{name}", f"r{rank}", rank, world, quorum_id=qi + 1)
        gen = torch.Generator(device="cuda").manual_seed(100 + rank)
        for dtype in (torch.float32, torch.bfloat16):
            for op in (ReduceOp.SUM, ReduceOp.AVG):
                for n in (256, 1024, 2048):
                    ts = [(torch.rand(n, n // 4 or 1, device=dev, generator=gen) * 9 + 1).to(dtype) for _ in range(3)]
                    refs = [exact(t, op == ReduceOp.AVG) for t in ts]
                    allreduce_quantized(ts, op, pg).wait()
                    torch.cuda.synchronize()
                    for t, r in zip(ts, refs):
                        rel = ((t.float() - r).abs().mean() / r.abs().mean()).item()
                        if not rel <= 0.04:
                            fails.append({"what": "allreduce_quantized", "pg": name, "dtype": str(dtype), "op": str(op), "n": n, "rel": rel})
        if name == "nccl":  # reduce_scatter_quantized rides on alltoall_base
            for op in (ReduceOp.SUM, ReduceOp.AVG):
                ts = [(torch.rand(64 * world, 128, device=dev, generator=gen) + 0.5) for _ in range(2)]
                out, padded = allocate_reduce_scatter_output(ts, world)
                refs = [exact(t, op == ReduceOp.AVG) for t in ts]
                reduce_scatter_quantized(out, ts, op, pg).wait()
                torch.cuda.synchronize()
                off = 0
                for r, ps in zip(refs, padded):
                    rows = ps[0] // world
                    mine = r[rank * rows:(rank + 1) * rows].reshape(-1)
                    got = out[off: off + mine.numel()]
                    off += mine.numel()
                    rel = ((got - mine).abs().mean() / mine.abs().mean()).item()
                    if not rel <= 0.05:
                        fails.append({"what": "reduce_scatter_quantized", "op": str(op), "rel": rel})
        pg.shutdown()

    # subprocess NCCL: tensors travel by CUDA IPC, the communicator lives in a child we could kill
    baby = ProcessGroupBabyNCCL(timeout=timedelta(seconds=60))
    baby.configure(f"{store_addr}/coll/baby", f"r{rank}", rank, world, quorum_id=9)
    x = torch.full((1 << 16,), float(rank + 1), device=dev)
    baby.allreduce([x], ReduceOp.SUM).wait()
    torch.cuda.synchronize()
    if not bool((x == sum(range(1, world + 1))).all()):
        fails.append({"what": "baby_nccl_allreduce", "got": x[:4].tolist()})
    if baby.num_active_work() != 0:
        fails.append({"what": "baby_nccl_active_work", "n": baby.num_active_work()})
    baby.shutdown()

    nf = torch.tensor([len(fails)])
    dist.all_reduce(nf)
    if fails:
        print(f"[rank {rank}] FAIL " + json.dumps(fails[:8]), flush=True)
    if rank == 0:
        print("GPU_COLLECTIVES " + json.dumps({"world": world, "failures": int(nf.item())}), flush=True)
    dist.destroy_process_group()
    sys.exit(1 if int(nf.item()) else 0)


if __name__ == "__main__":
    main()
