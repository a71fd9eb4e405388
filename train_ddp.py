"""Fault-tolerant DDP example: toy CNN on synthetic CIFAR-shaped data.

One process per replica group (run a Lighthouse first):

    python -m torchft_b200.lighthouse --min_replicas 1 --join_timeout_ms 2000 &
    TORCHFT_LIGHTHOUSE=http://127.0.0.1:29510 REPLICA_GROUP_ID=0 NUM_REPLICA_GROUPS=2 python train_ddp.py
    TORCHFT_LIGHTHOUSE=http://127.0.0.1:29510 REPLICA_GROUP_ID=1 NUM_REPLICA_GROUPS=2 python train_ddp.py

Kill either trainer at any time and start it again: it rejoins, heals its weights live
from the survivor, and continues. On a CUDA box the process group is ``ProcessGroupB200``
(NVLink peer-memory collectives + NVLink P2P heal); on CPU it is Gloo + HTTP heal.
Counterpart of the reference's train_ddp.py (same structure: sampler sharded over replica
groups, Manager + DistributedDataParallel + Optimizer wrapper, periodic state logging).
"""

from __future__ import annotations

import json
import logging
import os
import sys
from datetime import timedelta

import torch
import torch.nn.functional as F
from torch import nn, optim
from torch.distributed import TCPStore
from torch.utils.data import DataLoader, TensorDataset

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from torchft_b200 import (  # noqa: E402
    DistributedDataParallel,
    DistributedSampler,
    Manager,
    Optimizer,
    ProcessGroupGloo,
)

logging.basicConfig(level=os.environ.get("LOGLEVEL", "INFO"))


class Net(nn.Module):
    def __init__(self) -> None:
        super().__init__()
        self.cnn = nn.Sequential(
            nn.Conv2d(3, 6, 5), nn.ReLU(), nn.MaxPool2d(2, 2), nn.Conv2d(6, 16, 5), nn.ReLU(), nn.MaxPool2d(2, 2))
        self.classifier = nn.Sequential(nn.Linear(16 * 5 * 5, 120), nn.ReLU(), nn.Linear(120, 84), nn.ReLU(), nn.Linear(84, 10))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.classifier(torch.flatten(self.cnn(x), 1))


def main() -> None:
    replica_group = int(os.environ.get("REPLICA_GROUP_ID", 0))
    num_groups = int(os.environ.get("NUM_REPLICA_GROUPS", 2))
    total_steps = int(os.environ.get("TRAIN_STEPS", 50))
    min_replicas = int(os.environ.get("MIN_REPLICAS", 1))
    out_path = os.environ.get("TRAIN_OUT", "").replace("{group}", str(replica_group))  # one file per replica group
    use_cuda = torch.cuda.is_available() and os.environ.get("USE_CPU", "0") != "1"
    device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", replica_group % max(torch.cuda.device_count(), 1)))) if use_cuda else torch.device("cpu")
    if use_cuda:
        torch.cuda.set_device(device)

    # synthetic CIFAR-10 shaped dataset (no network in this environment)
    g = torch.Generator().manual_seed(0)
    data = TensorDataset(torch.randn(2048, 3, 32, 32, generator=g), torch.randint(0, 10, (2048,), generator=g))
    sampler = DistributedSampler(data, replica_rank=replica_group, num_replica_groups=num_groups, group_rank=0,
                                 num_replicas=1, shuffle=True)
    loader = DataLoader(data, batch_size=64, sampler=sampler)

    # the replica group's own store (world size 1 inside the group => we host it)
    store = TCPStore("127.0.0.1", 0, is_master=True, wait_for_workers=False)

    torch.manual_seed(replica_group)  # deliberately different: init_sync makes them equal
    m = Net().to(device)
    inner = optim.AdamW(m.parameters(), lr=1e-3)

    def load_state_dict(sd):
        m.load_state_dict(sd["model"])
        inner.load_state_dict(sd["optim"])

    def state_dict():
        return {"model": m.state_dict(), "optim": inner.state_dict()}

    if use_cuda:
        from torchft_b200 import ProcessGroupB200

        pg = ProcessGroupB200(timeout=timedelta(seconds=10))
    else:
        pg = ProcessGroupGloo(timeout=timedelta(seconds=10))

    manager = Manager(
        pg=pg, min_replica_size=min_replicas, load_state_dict=load_state_dict, state_dict=state_dict,
        replica_id=f"train_ddp_{replica_group}", store_addr="127.0.0.1", store_port=store.port, rank=0, world_size=1,
        timeout=timedelta(seconds=10), quorum_timeout=timedelta(seconds=30))
    ddp = DistributedDataParallel(manager, m)
    opt = Optimizer(manager, inner)
    crit = nn.CrossEntropyLoss()
    # CKPT_DIR=/shared/dir: durable checkpoints for whole-job restarts (live healing covers partial failures)
    ckpt = None
    if os.environ.get("CKPT_DIR"):
        from torchft_b200.checkpointing import DurableCheckpointer

        ckpt = DurableCheckpointer(manager, state_dict=state_dict, load_state_dict=load_state_dict,
                                   directory=os.environ["CKPT_DIR"], every_n_steps=int(os.environ.get("CKPT_EVERY", 50)))
        resumed = ckpt.restore()
        print(f"[{replica_group}] resumed_from_step={resumed}", flush=True)
    injector = None
    if os.environ.get("TORCHFT_FAILURE_PORT_FILE"):  # chaos testing (examples/orchestrator)
        from torchft_b200.failure import FailureInjector

        injector = FailureInjector(manager).start()
    print(m, f"{sum(p.numel() for p in m.parameters())} params", flush=True)

    # PROFILE_DIR=/some/dir: chrome traces with the torchft::manager::* spans (quorum, configure, allreduce,
    # should_commit, checkpoint send/recv) next to the kernels -- the reference's examples do the same
    prof = None
    if os.environ.get("PROFILE_DIR"):
        from torch.profiler import ProfilerActivity, profile, schedule, tensorboard_trace_handler

        acts = [ProfilerActivity.CPU] + ([ProfilerActivity.CUDA] if use_cuda else [])
        prof = profile(activities=acts, schedule=schedule(wait=5, warmup=2, active=10, repeat=2),
                       on_trace_ready=tensorboard_trace_handler(os.path.join(os.environ["PROFILE_DIR"], f"group{replica_group}")),
                       record_shapes=True)
        prof.start()

    epoch = 0
    while manager.current_step() < total_steps:
        sampler.set_epoch(epoch)
        epoch += 1
        for x, y in loader:
            x, y = x.to(device), y.to(device)
            if injector is not None:
                injector.maybe_stall()
            opt.zero_grad()       # starts the (async) quorum for this step
            loss = crit(ddp(x), y)
            loss.backward()       # gradients all-reduced across the live replica groups
            opt.step()            # only applied if the step committed
            if prof is not None:
                prof.step()
            if ckpt is not None:
                ckpt.maybe_save()
            if manager.current_step() % 10 == 0:
                print(f"[{replica_group}] step={manager.current_step()} batches_committed={manager.batches_committed()} "
                      f"participants={manager.num_participants()} loss={loss.item():.4f}", flush=True)
            if manager.current_step() >= total_steps:
                break

    if prof is not None:
        prof.stop()
    if ckpt is not None:
        ckpt.maybe_save(force=True)
        ckpt.wait()
    if out_path:
        torch.save({"model": {k: v.cpu() for k, v in m.state_dict().items()}, "step": manager.current_step()}, out_path)
    print(json.dumps({"replica_group": replica_group, "final_step": manager.current_step(),
                      "batches_committed": manager.batches_committed()}), flush=True)
    manager.shutdown(wait=False)
    pg.shutdown()


if __name__ == "__main__":
    main()
