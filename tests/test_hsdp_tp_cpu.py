"""FSDP2 + tensor parallelism inside a replica group, fault-tolerant averaging across groups, end to end on CPU (gloo):
2 replica groups x 2 TP ranks = 4 processes, one Lighthouse. Reference coverage: fsdp_test.py:64-101 (mocked PG)."""
import json
import os
import subprocess
import sys

import torch
from torch import nn

from torchft_b200.coordination import LighthouseServer

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_replica_groups_with_tensor_parallel_ranks_match_a_single_process_run(tmp_path):
    groups, tp, steps = 2, 2, 3
    lh = LighthouseServer(bind="[::]:0", min_replicas=groups, join_timeout_ms=60000)
    addr = lh.address()
    host = addr.split("//")[1].rsplit(":", 1)[0]
    addr = addr.replace(host, "127.0.0.1")
    out = str(tmp_path / "tp")
    procs = []
    try:
        for g in range(groups):
            env = dict(os.environ, REPLICA_GROUP_ID=str(g), NUM_REPLICA_GROUPS=str(groups), TORCHFT_LIGHTHOUSE=addr, OMP_NUM_THREADS="1")
            for k in ("RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT", "LOCAL_RANK"):
                env.pop(k, None)
            procs.append(subprocess.Popen(
                [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={tp}", "--master-addr=127.0.0.1",
                 f"--master-port={29940 + g}", os.path.join(ROOT, "tests", "_hsdp_tp_worker.py"), str(steps), out],
                env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
        outs = [p.communicate(timeout=420)[0] for p in procs]
        assert all(p.returncode == 0 for p in procs), "\n".join(o[-3000:] for o in outs)
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
        lh.shutdown()

    got = [json.load(open(f"{out}.g{g}")) for g in range(groups)]
    assert all(r["step"] == steps and r["participants"] == groups and r["hook_calls"] == steps for r in got), got

    # single-process reference: same init, loss = mean over groups of the per-group loss, same optimizer
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from _hsdp_tp_worker import build

    model = build()
    opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9)
    gen = torch.Generator().manual_seed(7)
    data = torch.randn(steps, groups, 6, 8, generator=gen)
    for s in range(steps):
        opt.zero_grad()
        torch.stack([model(data[s, g]).pow(2).mean() for g in range(groups)]).mean().backward()
        opt.step()
    for k, v in model.state_dict().items():
        a, b = torch.tensor(got[0]["params"][k]), torch.tensor(got[1]["params"][k])
        assert torch.equal(a, b), f"replica groups diverged on {k}"
        assert torch.allclose(a, v.detach(), rtol=1e-5, atol=1e-6), (k, float((a - v).abs().max()))
