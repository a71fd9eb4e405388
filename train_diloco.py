"""Fault-tolerant (Streaming) DiLoCo example.

    python -m torchft_b200.lighthouse --min_replicas 1 --join_timeout_ms 2000 &
    TORCHFT_LIGHTHOUSE=http://127.0.0.1:29510 REPLICA_GROUP_ID=0 NUM_REPLICA_GROUPS=2 python train_diloco.py
    TORCHFT_LIGHTHOUSE=http://127.0.0.1:29510 REPLICA_GROUP_ID=1 NUM_REPLICA_GROUPS=2 python train_diloco.py

Each replica trains with its inner optimizer; every ``SYNC_EVERY`` steps one model *fragment*
synchronises: the averaged pseudo-gradient (last global weights - local weights) is applied by a
Nesterov-SGD outer optimizer. ``USE_STREAMING=1`` splits the model into ``N_FRAGMENTS`` fragments
that take turns (Streaming DiLoCo); ``QUANTIZE=1`` sends fp8 pseudo-gradients (on a CUDA box with
ProcessGroupB200: one fused kernel for delta + quantise + exchange + reduce + dequantise).
Counterpart of the reference's train_diloco.py (which uses torch.distributed.pipelining only to
obtain the fragments; here the layer stack is sliced directly).
"""

from __future__ import annotations

import json
import os
import sys
from datetime import timedelta

import torch
from torch import nn, optim
from torch.distributed import TCPStore

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from torchft_b200 import Manager, ProcessGroupGloo  # noqa: E402
from torchft_b200.local_sgd import DiLoCo  # noqa: E402


class MultiMLP(nn.Module):
    def __init__(self, d: int = 64, hidden: int = 128, n_layers: int = 8) -> None:
        super().__init__()
        self.layers = nn.ModuleList(nn.Sequential(nn.Linear(d, hidden), nn.GELU(), nn.Linear(hidden, d)) for _ in range(n_layers))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        for layer in self.layers:
            x = x + layer(x)
        return x


def main() -> None:
    group = int(os.environ.get("REPLICA_GROUP_ID", 0))
    steps = int(os.environ.get("TRAIN_STEPS", 200))
    sync_every = int(os.environ.get("SYNC_EVERY", 20))
    streaming = os.environ.get("USE_STREAMING", "0") == "1"
    n_frag = int(os.environ.get("N_FRAGMENTS", 2)) if streaming else 1
    quantize = os.environ.get("QUANTIZE", "0") == "1"
    use_cuda = torch.cuda.is_available() and os.environ.get("USE_CPU", "0") != "1"
    device = torch.device("cuda", group % max(torch.cuda.device_count(), 1)) if use_cuda else torch.device("cpu")
    if use_cuda:
        torch.cuda.set_device(device)

    store = TCPStore("127.0.0.1", 0, is_master=True, wait_for_workers=False)
    torch.manual_seed(0)
    model = MultiMLP().to(device)
    per = len(model.layers) // n_frag
    fragments = [nn.Sequential(*model.layers[i * per : (i + 1) * per]) for i in range(n_frag)]
    inner = optim.AdamW(model.parameters(), lr=3e-4, weight_decay=0.1, betas=(0.9, 0.95))
    outers = [optim.SGD(f.parameters(), lr=0.7, momentum=0.9, nesterov=True) for f in fragments]

    if use_cuda:
        from torchft_b200 import ProcessGroupB200

        pg = ProcessGroupB200(timeout=timedelta(seconds=10))
    else:
        pg = ProcessGroupGloo(timeout=timedelta(seconds=10))

    def load_state_dict(sd):
        model.load_state_dict(sd["model"])
        inner.load_state_dict(sd["inner"])

    manager = Manager(
        pg=pg, min_replica_size=int(os.environ.get("MIN_REPLICAS", 1)), use_async_quorum=False,  # DiLoCo needs a sync quorum
        load_state_dict=load_state_dict, state_dict=lambda: {"model": model.state_dict(), "inner": inner.state_dict()},
        replica_id=f"train_diloco_{group}", store_addr="127.0.0.1", store_port=store.port, rank=0, world_size=1,
        timeout=timedelta(seconds=10), quorum_timeout=timedelta(seconds=60))

    gen = torch.Generator(device="cpu").manual_seed(100 + group)
    with DiLoCo(manager, fragments, inner, outers, sync_every=sync_every, backup_device=device if use_cuda else None,
                should_quantize=quantize and use_cuda, use_bucketization=True, bucket_cap_mb=32,
                fragment_sync_delay=int(os.environ.get("SYNC_DELAY", 0)),
                fragment_update_alpha=float(os.environ.get("ALPHA", 0.0))):
        prof = None
        if os.environ.get("PROFILE_DIR"):  # chrome traces incl. the torchft::local_sgd::* spans
            from torch.profiler import ProfilerActivity, profile, schedule, tensorboard_trace_handler

            acts = [ProfilerActivity.CPU] + ([ProfilerActivity.CUDA] if use_cuda else [])
            prof = profile(activities=acts, schedule=schedule(wait=2, warmup=2, active=2 * sync_every, repeat=1),
                           on_trace_ready=tensorboard_trace_handler(os.path.join(os.environ["PROFILE_DIR"], f"group{group}")))
            prof.start()
        for i in range(steps):
            x = torch.randn(32, 64, generator=gen).to(device)
            inner.zero_grad()
            loss = (model(x) - x.flip(-1)).pow(2).mean()
            loss.backward()
            inner.step()  # hooks: prepare/perform the outer sync on schedule
            if prof is not None:
                prof.step()
            if (i + 1) % sync_every == 0:
                print(f"[{group}] inner_step={i + 1} outer_step={manager.current_step()} "
                      f"participants={manager.num_participants()} loss={loss.item():.4f}", flush=True)
        if prof is not None:
            prof.stop()
    print(json.dumps({"replica_group": group, "outer_steps": manager.current_step()}), flush=True)
    manager.shutdown(wait=False)
    pg.shutdown()


if __name__ == "__main__":
    main()
