"""GPU integration tests: full fault-tolerant step on one GPU; multi-GPU paths when >= 2 GPUs."""

import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ft_smoke_step_single_gpu():
    from torchft_b200.bench_utils import ft_smoke_step
    from torchft_b200.ops import _native

    before = _native.kernel_launches()
    loss = ft_smoke_step(steps=3)
    assert loss == loss and 0 < loss < 20
    assert _native.kernel_launches() - before > 30  # native kernels actually ran


def test_overlapped_optimizer_is_bit_identical_to_single_launch():
    """Per-stage AdamW on the side stream (next forward gated stage by stage) must produce exactly
    the parameters of the one-launch update."""
    from datetime import timedelta

    from torchft_b200.bench_utils import local_lighthouse, loopback
    from torchft_b200.parallel.trainer import FaultTolerantTrainer

    results = []
    for overlap in (False, True):
        lh = local_lighthouse()
        tr = FaultTolerantTrainer("llama3_debug", loopback(lh.address()), replica_id=f"ovl_{int(overlap)}_0",
                                  timeout=timedelta(seconds=30), bucket_mb=1.0, overlap_optimizer=overlap,
                                  optimizer_blocks=7 if overlap else 0, zero1=False)
        try:
            g = torch.Generator().manual_seed(5)
            losses = []
            for _ in range(4):
                tok = torch.randint(0, tr.cfg.vocab_size, (2, 128), generator=g).pin_memory()
                tgt = torch.randint(0, tr.cfg.vocab_size, (2, 128), generator=g).pin_memory()
                losses.append(tr.step(tok, tgt))
            sd = tr.state_dict()  # joins the optimizer stream
            torch.cuda.synchronize()
            results.append((losses, sd["param"].clone(), sd["m"].clone(), sd["v"].clone()))
            assert len(tr._opt_ranges if overlap else []) == (tr.cfg.n_layers + 2 if overlap else 0)
        finally:
            tr.shutdown()
            lh.shutdown()
    (l0, p0, m0, v0), (l1, p1, m1, v1) = results
    assert l0 == l1
    assert torch.equal(p0, p1) and torch.equal(m0, m1) and torch.equal(v0, v1)


def test_graft_smoke_entry():
    sys.path.insert(0, ROOT)
    import __graft_entry__

    __graft_entry__.smoke()


def test_process_group_b200_world1_collectives():
    from datetime import timedelta

    from torch.distributed import ReduceOp, TCPStore
    from torch.distributed.distributed_c10d import AllreduceOptions, BroadcastOptions

    from torchft_b200.parallel.process_group_b200 import ProcessGroupB200

    store = TCPStore("127.0.0.1", 0, is_master=True, wait_for_workers=False)
    pg = ProcessGroupB200(timeout=timedelta(seconds=10))
    pg.configure(f"127.0.0.1:{store.port}/t/1", "r0", 0, 1, quorum_id=1)
    x = torch.arange(1000, device="cuda", dtype=torch.float32)
    o = AllreduceOptions()
    o.reduceOp = ReduceOp.AVG
    w = pg.allreduce([x], o)
    w.wait()
    torch.cuda.synchronize()
    assert torch.equal(x, torch.arange(1000, device="cuda", dtype=torch.float32))
    y = torch.ones(64, device="cuda", dtype=torch.bfloat16)
    pg.allreduce_native(y, scale=0.5).wait()
    torch.cuda.synchronize()
    assert torch.all(y == 0.5)
    z = torch.ones(64, device="cuda")
    pg.allreduce_native(z, contribute=False).wait()
    torch.cuda.synchronize()
    assert torch.all(z == 0)
    b = BroadcastOptions()
    b.rootRank = 0
    pg.broadcast([x], b).wait()
    pg.barrier().wait()
    assert pg.errored() is None
    # reconfigure keeps working (remap, not re-create)
    pg.configure(f"127.0.0.1:{store.port}/t/2", "r0", 0, 1, quorum_id=2)
    pg.allreduce([x], o).wait()
    assert pg.errored() is None
    pg.shutdown()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_comm_bench_two_gpus(tmp_path):
    out = tmp_path / "comm.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "bench", "comm_bench.py"), "--quick", "--out", str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    res = json.loads(out.read_text())
    assert res["all_ok"], [c for c in res["correctness"] + res["q8"] if not c.get("ok")]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_quantized_collectives_and_baby_nccl_two_gpus():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29537", os.path.join(ROOT, "tests", "_gpu_collectives_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert '"failures": 0' in r.stdout


def test_stream_timeout_fires_only_on_a_stuck_stream():
    import threading
    import time
    from datetime import timedelta

    from torchft_b200.futures import stream_timeout

    late = threading.Event()
    torch.cuda._sleep(int(3e9))  # > 1 s of device time
    stream_timeout(late.set, timedelta(milliseconds=100))
    assert late.wait(3.0)
    torch.cuda.synchronize()
    fine = threading.Event()
    torch.ones(8, device="cuda").sum()
    stream_timeout(fine.set, timedelta(milliseconds=200))
    time.sleep(0.6)
    assert not fine.is_set()
