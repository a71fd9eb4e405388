"""One rank of one replica group for tests/test_hsdp_tp_cpu.py: tensor parallelism (Colwise/Rowwise) inside the group,
FSDP2 over the group's (size-1) dp_shard dimension for the gradient hook, torchft_b200 across groups -- on CPU over gloo.
The reference checks this composition with mocks only (fsdp_test.py:64-101); here two real replica groups train and must
match a single-process run."""
import json
import os
import sys
from datetime import timedelta

import torch
import torch.distributed as dist
from torch import nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def build():
    torch.manual_seed(0)
    return nn.Sequential(nn.Linear(8, 16, bias=False), nn.ReLU(), nn.Linear(16, 8, bias=False))


def main() -> None:
    from torch.distributed.device_mesh import init_device_mesh
    from torch.distributed.fsdp import fully_shard
    from torch.distributed.tensor.parallel import ColwiseParallel, RowwiseParallel, parallelize_module

    from torchft_b200 import ManagedProcessGroup, Manager, Optimizer, ProcessGroupGloo
    from torchft_b200.parallel.hsdp import fsdp_local_state, load_fsdp_local_state

    group, groups, steps, out = int(os.environ["REPLICA_GROUP_ID"]), int(os.environ["NUM_REPLICA_GROUPS"]), int(sys.argv[1]), sys.argv[2]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", timeout=timedelta(seconds=120))
    mesh = init_device_mesh("cpu", (1, world), mesh_dim_names=("dp_shard", "tp"))
    model = build()
    parallelize_module(model, mesh["tp"], {"0": ColwiseParallel(), "2": RowwiseParallel()})
    fully_shard(model, mesh=mesh["dp_shard"])
    inner = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9)
    manager = Manager(pg=ProcessGroupGloo(timeout=timedelta(seconds=60)), load_state_dict=None, state_dict=None,
                      min_replica_size=groups, replica_id=f"tp_{group}", timeout=timedelta(seconds=60),
                      quorum_timeout=timedelta(seconds=120), init_sync=False)
    manager.register_state_dict_fn("hsdp", lambda sd: load_fsdp_local_state(model, inner, sd), lambda: fsdp_local_state(model, inner))
    replicate = ManagedProcessGroup(manager)
    calls = []

    def hook(shard_grad: torch.Tensor) -> None:
        calls.append(tuple(shard_grad.shape))
        replicate.allreduce([shard_grad], dist.ReduceOp.AVG).wait()

    model.set_all_reduce_hook(hook)
    opt = Optimizer(manager, inner)
    gen = torch.Generator().manual_seed(7)
    data = torch.randn(steps, groups, 6, 8, generator=gen)
    for s in range(steps):
        opt.zero_grad()
        model(data[s, group]).pow(2).mean().backward()  # both TP ranks of a group see the group's batch
        opt.step()
    full = {k: v.full_tensor().tolist() for k, v in model.state_dict().items()}
    if rank == 0:
        with open(f"{out}.g{group}", "w") as f:
            json.dump({"params": full, "step": manager.current_step(), "hook_calls": len(calls), "participants": manager.num_participants()}, f)
    dist.barrier()
    manager.shutdown(wait=False)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
