"""All-reduce correctness + bandwidth sweep: native NVLink kernels vs NCCL.

BASELINE.md config 5: "allreduce 1 KB-1 GB at 2/4/8 replicas; quorum-reconfigure
latency vs ncclCommAbort+ncclCommInitRank". Launch with

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N \
        --master-addr 127.0.0.1 --master-port 29511 bench/comm_bench.py [--quick]

Every number is device-timed (CUDA events) and reduced with MAX over ranks.
Results: one JSON document on rank 0 (stdout + ``--out`` file).
"""

from __future__ import annotations

import argparse
import json
import os
import sys
import time
from datetime import timedelta

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from torchft_b200.parallel.symm_mem import SymmetricComm  # noqa: E402


def device_time_ms(fn, iters: int, warmup: int, stream=None) -> float:
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / iters
    t = torch.tensor([ms], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--out", default="gpurun_out/comm_bench.json")
    ap.add_argument("--max-mb", type=int, default=1024)
    ap.add_argument("--blocks", type=str, default="")
    args = ap.parse_args()

    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", timeout=timedelta(seconds=120), device_id=torch.device("cuda", local))
    store = dist.distributed_c10d._get_default_store()

    res = {"world": world, "correctness": [], "sweep": [], "reconfigure": {}, "q8": []}
    comm = SymmetricComm(timeout=timedelta(seconds=20))
    max_bytes = (64 if args.quick else args.max_mb) << 20
    symm = comm.alloc("bench", max_bytes)

    # ---- reconfigure latency: remap peer memory vs NCCL communicator re-creation ----
    t0 = time.perf_counter()
    comm.configure(dist.PrefixStore("q1", store), rank, world, 1)
    first = (time.perf_counter() - t0) * 1e3
    times = []
    for q in range(2, 6):
        dist.barrier()
        t0 = time.perf_counter()
        comm.configure(dist.PrefixStore(f"q{q}", store), rank, world, q)
        times.append((time.perf_counter() - t0) * 1e3)
    nccl_times = []
    for q in range(3):
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        g = dist.new_group(list(range(world)), backend="nccl")
        x = torch.ones(1, device="cuda")
        dist.all_reduce(x, group=g)  # forces ncclCommInitRank
        torch.cuda.synchronize()
        nccl_times.append((time.perf_counter() - t0) * 1e3)
        dist.destroy_process_group(g)
    rt = torch.tensor([first, sum(times) / len(times), min(times), sum(nccl_times) / len(nccl_times), min(nccl_times)], device="cuda")
    dist.all_reduce(rt, op=dist.ReduceOp.MAX)
    res["reconfigure"] = dict(zip(["native_first_ms", "native_remap_avg_ms", "native_remap_min_ms", "nccl_new_comm_avg_ms", "nccl_new_comm_min_ms"], [round(v, 3) for v in rt.tolist()]))

    res["symm_mode"] = comm._mode
    res["nvls"] = bool(comm._mc)
    if comm._mc:
        comm._nvls_min = 4096  # make the correctness sweep below cover the NVLS kernel as well
    # ---- correctness vs fp32 reference (bit-level vs NCCL is order dependent) ----
    gen = torch.Generator(device="cuda").manual_seed(1234 + rank)
    for dtype in (torch.float32, torch.bfloat16, torch.float16):
        for n in (1, 7, 1000, 4096, 65536 + 3, 1 << 20, (8 << 20) + 5):
            es = torch.empty(0, dtype=dtype).element_size()
            x = torch.randn(n, device="cuda", generator=gen).to(dtype)
            ref = x.float().clone()
            dist.all_reduce(ref)
            ref = ref / world
            for path in ("staged", "symm"):
                if path == "symm":
                    buf = symm[: n * es].view(dtype)
                    buf.copy_(x)
                    y = buf
                else:
                    y = x.clone()
                comm.allreduce_(y, scale=1.0 / world)
                torch.cuda.synchronize()
                err = (y.float() - ref).abs().max().item()
                tol = 1e-5 if dtype == torch.float32 else (2e-2 if dtype == torch.bfloat16 else 4e-3)
                ok = err <= tol * max(1.0, ref.abs().max().item())
                e = comm.errored()
                res["correctness"].append({"dtype": str(dtype), "n": n, "path": path, "max_err": err, "ok": bool(ok) and e is None, "err": str(e) if e else None})
    if comm._mc:
        # NVLS with a non-participant (zeroes its own copy first) on the symmetric segment
        n = 1 << 18
        buf = symm[: n * 4].view(torch.float32)
        buf.fill_(float(rank + 1))
        comm.allreduce_(buf, scale=0.5, contribute=(rank != 0))
        torch.cuda.synchronize()
        expect = 0.5 * sum(r + 1 for r in range(1, world))
        res["correctness"].append({"case": "nvls_non_participant", "ok": bool((buf == expect).all().item())})
    # non-participant contributes zeros
    x = torch.full((4096,), float(rank + 1), device="cuda")
    comm.allreduce_(x, scale=1.0, contribute=(rank != 0))
    torch.cuda.synchronize()
    expect = sum(r + 1 for r in range(1, world))
    res["correctness"].append({"case": "non_participant_zero", "ok": bool((x == expect).all().item())})
    # max
    x = torch.full((5000,), float(rank), device="cuda")
    comm.allreduce_(x, op=1)
    torch.cuda.synchronize()
    res["correctness"].append({"case": "max", "ok": bool((x == world - 1).all().item())})

    # ---- q8 allreduce accuracy (reference tolerance: mean rel err <= 0.04) ----
    for dtype in (torch.float32, torch.bfloat16):
        for n in (512, 100_000, (4 << 20) + 17):
            a = torch.randn(n, device="cuda", generator=gen).to(dtype) * 3
            b = torch.randn(n, device="cuda", generator=gen).to(dtype)
            ref = (a.float() - b.float()).clone()
            dist.all_reduce(ref)
            ref /= world
            out = torch.empty_like(a)
            comm.q8_allreduce_(out, a, b, scale=1.0 / world)
            torch.cuda.synchronize()
            rel = ((out.float() - ref).abs().mean() / ref.abs().mean()).item()
            e = comm.errored()
            res["q8"].append({"dtype": str(dtype), "n": n, "mean_rel_err": rel, "ok": rel <= 0.04 and e is None, "err": str(e) if e else None})

    # ---- bandwidth sweep ----
    sizes = [1 << k for k in range(10, 31)]
    sizes = [s for s in sizes if s <= max_bytes]
    if args.quick:
        sizes = [s for s in sizes if s in (1 << 10, 1 << 16, 1 << 20, 1 << 24, 1 << 26)]
    block_opts = [int(b) for b in args.blocks.split(",") if b] or [comm._max_blocks]
    for nbytes in sizes:
        n = nbytes // 2
        iters = 50 if nbytes <= (1 << 22) else (20 if nbytes <= (1 << 26) else 8)
        row = {"bytes": nbytes}
        buf = symm[:nbytes].view(torch.bfloat16)
        buf.normal_()
        x = torch.randn(n, device="cuda").to(torch.bfloat16)
        busf = 2 * (world - 1) / world
        nvls_min_saved = comm._nvls_min
        comm._nvls_min = 1 << 62  # P2P kernels only in this column
        for nb in block_opts:
            comm._max_blocks = nb
            ms = device_time_ms(lambda: comm.allreduce_(buf, scale=1.0 / world), iters, 3)
            row[f"native_symm_b{nb}_ms"] = round(ms, 5)
            row[f"native_symm_b{nb}_busbw_gbs"] = round(nbytes * busf / ms / 1e6, 1)
        comm._nvls_min = nvls_min_saved
        comm._max_blocks = block_opts[-1]
        if comm._mc:  # VMM mode with an NVLS multicast object: time the in-switch reduction too
            saved = comm._nvls_min
            comm._nvls_min = 0
            ms = device_time_ms(lambda: comm.allreduce_(buf, scale=1.0 / world), iters, 3)
            row["native_nvls_ms"] = round(ms, 5)
            row["native_nvls_busbw_gbs"] = round(nbytes * busf / ms / 1e6, 1)
            comm._nvls_min = saved
        ms = device_time_ms(lambda: comm.allreduce_(x, scale=1.0 / world), iters, 3)
        row["native_staged_ms"] = round(ms, 5)
        row["native_staged_busbw_gbs"] = round(nbytes * busf / ms / 1e6, 1)

        def nccl_ref():
            dist.all_reduce(x)
            x.div_(world)  # reference: eager tensor /= num_participants (manager.py:477-478)

        ms = device_time_ms(nccl_ref, iters, 3)
        row["nccl_sum_div_ms"] = round(ms, 5)
        row["nccl_sum_div_busbw_gbs"] = round(nbytes * busf / ms / 1e6, 1)
        ms = device_time_ms(lambda: dist.all_reduce(x, op=dist.ReduceOp.AVG), iters, 3)
        row["nccl_avg_ms"] = round(ms, 5)
        row["nccl_avg_busbw_gbs"] = round(nbytes * busf / ms / 1e6, 1)
        if nbytes >= (1 << 16):
            out = torch.empty_like(x)
            ms = device_time_ms(lambda: comm.q8_allreduce_(out, x, None, scale=1.0 / world), iters, 3)
            row["native_q8_ms"] = round(ms, 5)
        e = comm.errored()
        if e is not None:
            row["error"] = str(e)
        res["sweep"].append(row)
        if rank == 0:
            print(json.dumps(row), flush=True)

    res["all_ok"] = all(c.get("ok", False) for c in res["correctness"]) and all(c["ok"] for c in res["q8"])
    res["launches"] = comm.launches
    if rank == 0:
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        with open(args.out, "w") as f:
            json.dump(res, f, indent=1)
        print("COMM_BENCH_RESULT " + json.dumps({k: res[k] for k in ("world", "all_ok", "reconfigure")}), flush=True)
        bad = [c for c in res["correctness"] + res["q8"] if not c.get("ok", False)]
        if bad:
            print("FAILED CASES: " + json.dumps(bad[:10]), flush=True)
    comm.shutdown()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
