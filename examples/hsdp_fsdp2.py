"""Fault-tolerant HSDP: FSDP2 (``fully_shard``) inside each replica group, torchft_b200 across groups.

    python examples/hsdp_fsdp2.py --groups 2 --shards 2 --steps 4        # needs groups*shards GPUs

Orchestrator + workers in one file: the orchestrator starts a Lighthouse and one ``torchrun`` per
replica group; every worker builds the mesh of ITS group only, shards the model with FSDP2 and
installs ``set_all_reduce_hook`` so each reduce-scattered gradient shard is averaged across replica
groups through ``ManagedProcessGroup`` (-> ``Manager.allreduce`` -> fused P2P kernel). One Manager per
rank, one ManagerServer per group; quorum and ``should_commit`` are barriers over the group's ranks.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
from datetime import timedelta

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def build_model(torch, nn):
    torch.manual_seed(0)
    return nn.Sequential(nn.Linear(64, 256), nn.GELU(), nn.Linear(256, 256), nn.GELU(), nn.Linear(256, 64))


def worker(a: argparse.Namespace) -> None:
    import torch
    import torch.distributed as dist
    from torch import nn
    from torch.distributed.device_mesh import init_device_mesh
    from torch.distributed.fsdp import fully_shard

    from torchft_b200 import ManagedProcessGroup, Manager, Optimizer, ProcessGroupB200
    from torchft_b200.parallel.hsdp import fsdp_local_state, load_fsdp_local_state

    group = int(os.environ["REPLICA_GROUP_ID"])
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", group * world + local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)  # intra-group (FSDP) collectives: plain NCCL, outside FT scope
    mesh = init_device_mesh("cuda", (world,), mesh_dim_names=("dp_shard",))

    model = build_model(torch, nn).to(dev)
    for layer in model:
        if isinstance(layer, nn.Linear):
            fully_shard(layer, mesh=mesh)
    fully_shard(model, mesh=mesh)
    inner = torch.optim.SGD(model.parameters(), lr=0.05)

    pg = ProcessGroupB200(timeout=timedelta(seconds=30))
    manager = Manager(pg=pg, min_replica_size=a.groups, load_state_dict=None, state_dict=None, replica_id=f"hsdp_{group}",
                      timeout=timedelta(seconds=30), quorum_timeout=timedelta(seconds=60), init_sync=False)
    # heal: rank i of a (re)joining group pulls rank i's parameter / optimizer shards of a healthy group, in place
    manager.register_state_dict_fn("hsdp", lambda sd: load_fsdp_local_state(model, inner, sd), lambda: fsdp_local_state(model, inner))
    replicate_pg = ManagedProcessGroup(manager)

    def cross_replica_hook(shard_grad: torch.Tensor) -> None:
        # FSDP2 calls this with the reduce-scattered gradient shard of one param group
        replicate_pg.allreduce([shard_grad], dist.ReduceOp.AVG).wait()

    for m in model.modules():
        if hasattr(m, "set_all_reduce_hook"):
            m.set_all_reduce_hook(cross_replica_hook)
    opt = Optimizer(manager, inner)

    gen = torch.Generator().manual_seed(1)
    data = torch.randn(a.steps, a.groups * world, 8, 64, generator=gen)  # [step, global rank, batch, feat]
    for step in range(a.steps):
        x = data[step, group * world + rank].to(dev)
        opt.zero_grad()
        loss = model(x).pow(2).mean()
        loss.backward()
        opt.step()
    full = {k: v.full_tensor().float().cpu() for k, v in model.state_dict().items()}
    if rank == 0:
        torch.save({"params": full, "step": manager.current_step()}, f"{a.out}.g{group}")
    steps = torch.tensor([manager.current_step()], device=dev)
    gathered = [torch.zeros_like(steps) for _ in range(world)]
    dist.all_gather(gathered, steps)
    if rank == 0:
        with open(f"{a.out}.g{group}.steps", "w") as f:
            json.dump([int(g.item()) for g in gathered], f)
    manager.shutdown(wait=False)
    pg.shutdown()
    dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--role", default="orchestrator")
    ap.add_argument("--groups", type=int, default=2)
    ap.add_argument("--shards", type=int, default=2)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--out", default="gpurun_out/hsdp.json")
    a = ap.parse_args()
    if a.role == "worker":
        worker(a)
        return
    import torch
    from torch import nn

    from torchft_b200.bench_utils import loopback
    from torchft_b200.coordination import LighthouseServer

    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    lh = LighthouseServer(bind="[::]:0", min_replicas=a.groups, join_timeout_ms=60000)
    procs = []
    for g in range(a.groups):
        env = dict(os.environ, REPLICA_GROUP_ID=str(g), NUM_REPLICA_GROUPS=str(a.groups), TORCHFT_LIGHTHOUSE=loopback(lh.address()))
        procs.append(subprocess.Popen(
            [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={a.shards}", "--master-addr=127.0.0.1",
             f"--master-port={29700 + g}", os.path.abspath(__file__), "--role", "worker", "--groups", str(a.groups), "--shards",
             str(a.shards), "--steps", str(a.steps), "--out", a.out], env=env))
    rcs = [p.wait(timeout=900) for p in procs]
    lh.shutdown()
    assert all(rc == 0 for rc in rcs), rcs

    # single-process reference: same model, same data, global batch = all ranks' batches
    model = build_model(torch, nn)
    opt = torch.optim.SGD(model.parameters(), lr=0.05)
    gen = torch.Generator().manual_seed(1)
    data = torch.randn(a.steps, a.groups * a.shards, 8, 64, generator=gen)
    for step in range(a.steps):
        opt.zero_grad()
        # mean over ranks of per-rank mean losses == what HSDP computes (AVG over shards and replicas)
        loss = torch.stack([model(data[step, r]).pow(2).mean() for r in range(a.groups * a.shards)]).mean()
        loss.backward()
        opt.step()
    ref = {k: v.detach().float() for k, v in model.state_dict().items()}
    got = [torch.load(f"{a.out}.g{g}") for g in range(a.groups)]
    steps = [s for g in range(a.groups) for s in json.load(open(f"{a.out}.g{g}.steps"))]
    diff_ref = max(float((got[0]["params"][k] - ref[k]).abs().max()) for k in ref)
    diff_groups = max(float((got[0]["params"][k] - got[g]["params"][k]).abs().max()) for k in ref for g in range(1, a.groups))
    res = {"steps_committed": steps, "max_param_diff_vs_reference": diff_ref, "max_param_diff_between_groups": diff_groups}
    with open(a.out, "w") as f:
        json.dump(res, f)
    print("HSDP " + json.dumps(res))


if __name__ == "__main__":
    main()
