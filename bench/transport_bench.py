"""Checkpoint-transport micro-benchmark: HTTP vs PG(NCCL send/recv) vs NVLink P2P heal.

Counterpart of the reference's http_transport_bench.py / pg_transport_bench.py (12 GB state_dict
of 3 MB fp32 tensors; scripts only, no recorded numbers). Two ranks: rank 0 serves, rank 1 heals.

    torchrun --nproc-per-node 2 bench/transport_bench.py --total-gb 12 --out gpurun_out/transport_bench.json
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

from torchft_b200.checkpointing import HTTPTransport, P2PTransport, PGTransport  # noqa: E402
from torchft_b200.process_group import ProcessGroupNCCL  # noqa: E402


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--total-gb", type=float, default=12.0)
    ap.add_argument("--chunk-mb", type=float, default=3.0)
    ap.add_argument("--out", default="gpurun_out/transport_bench.json")
    ap.add_argument("--skip-http", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ["RANK"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo", timeout=timedelta(seconds=600))
    store = dist.distributed_c10d._get_default_store()
    timeout = timedelta(seconds=300)
    n = int(args.total_gb * 1e9 / (args.chunk_mb * 1e6))
    numel = int(args.chunk_mb * 1e6 / 4)
    state = {f"t{i}": torch.full((numel,), float(i if rank == 0 else -1), device="cuda") for i in range(n)}
    nbytes = n * numel * 4
    res = {"bytes": nbytes, "tensors": n}

    def run(name, make):
        tr = make()
        meta = [tr.metadata() if rank == 0 else None]
        dist.broadcast_object_list(meta, src=0)
        dist.barrier()
        t0 = time.perf_counter()
        if rank == 0:
            tr.send_checkpoint([1], 1, state, timeout)
            dist.barrier()
            tr.disallow_checkpoint()
        else:
            got = tr.recv_checkpoint(0, meta[0], 1, timeout)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            ok = all(float(got[f"t{i}"][0]) == float(i) for i in (0, n // 2, n - 1))
            res[name] = {"seconds": round(dt, 3), "gbs": round(nbytes / dt / 1e9, 1), "ok": ok}
            if hasattr(tr, "last_recv_ms") and tr.last_recv_ms:
                res[name]["copy_kernel_ms"] = round(tr.last_recv_ms, 2)
                res[name]["copy_kernel_gbs"] = round(nbytes / tr.last_recv_ms / 1e6, 1)
            del got
            dist.barrier()
        dist.barrier()
        tr.shutdown()
        torch.cuda.empty_cache()

    run("p2p_nvlink_alloc", lambda: P2PTransport(timeout))
    run("p2p_nvlink_inplace", lambda: P2PTransport(timeout, state_dict=(lambda: state) if rank == 1 else None))
    if not args.skip_http:
        run("http", lambda: HTTPTransport(timeout, num_chunks=0))
        run("http_8chunks", lambda: HTTPTransport(timeout, num_chunks=8))
    pg = ProcessGroupNCCL(timeout=timeout)
    pg.configure(f"127.0.0.1:{os.environ['MASTER_PORT']}/tbench/pg", f"r{rank}", rank, 2)
    run("pg_nccl", lambda: PGTransport(pg, timeout, torch.device("cuda")))
    run("pg_nccl_inplace", lambda: PGTransport(pg, timeout, torch.device("cuda"), state_dict=(lambda: state) if rank == 1 else None))
    pg.shutdown()
    out = [res if rank == 1 else None]
    dist.broadcast_object_list(out, src=1)
    if rank == 0:
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        with open(args.out, "w") as f:
            json.dump(out[0], f, indent=1)
        print("TRANSPORT_BENCH " + json.dumps(out[0]))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
