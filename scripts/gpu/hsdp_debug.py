"""Debug aid: HSDPTrainer on ONE GPU (1 group x 1 shard): exercises FSDP2's hooks against the fused model ops and
reports where non-finite values first appear (gradients per parameter, then parameters after the optimizer step)."""
import os
import sys
from datetime import timedelta

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
os.environ.setdefault("MASTER_PORT", "29655")
os.environ.setdefault("RANK", "0")
os.environ.setdefault("WORLD_SIZE", "1")
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

torch.cuda.set_device(0)
dist.init_process_group("cpu:gloo,cuda:nccl", device_id=torch.device("cuda", 0))
from torchft_b200.bench_utils import local_lighthouse, loopback  # noqa: E402
from torchft_b200.parallel.hsdp import HSDPTrainer  # noqa: E402

model = sys.argv[1] if len(sys.argv) > 1 else "llama3_debug"
backend = sys.argv[2] if len(sys.argv) > 2 else "b200"
lh = local_lighthouse()
tr = HSDPTrainer(model, loopback(lh.address()), shards=1, backend=backend, timeout=timedelta(seconds=30))
cfg = tr.cfg
S = min(cfg.max_seq_len, int(os.environ.get("HSDP_DEBUG_SEQ", "512")))


def mem(tag):
    torch.cuda.synchronize()
    print(f"   [mem] {tag}: allocated {torch.cuda.memory_allocated() / 2**30:.2f} GiB, peak {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB", flush=True)


mem(f"after build ({sum(p.numel() for p in tr.model.parameters()) / 1e9:.2f} B params)")
tok = torch.randint(0, cfg.vocab_size, (2, S), device="cuda")
tgt = torch.randint(0, cfg.vocab_size, (2, S), device="cuda")


def local(t):
    return t.to_local() if hasattr(t, "to_local") else t


losses = []
for i in range(3):
    tr.optim.zero_grad(set_to_none=True)
    if getattr(tr, "_grad_arena", None) is not None:
        tr._grad_arena.begin_step()
    loss = tr.model(tok, tgt)
    mem("after forward")
    loss.backward()
    mem("after backward")
    bad = [(n, str(local(p.grad).dtype)) for n, p in tr.model.named_parameters() if p.grad is not None and not torch.isfinite(local(p.grad)).all()]
    none = [n for n, p in tr.model.named_parameters() if p.grad is None]
    gmax = max(float(local(p.grad).abs().max()) for p in tr.model.parameters() if p.grad is not None)
    print(f"step {i} loss {float(loss.detach()):.4f} non-finite grads: {bad[:8]} none: {none[:4]} gmax {gmax:.3e}", flush=True)
    tr.optim.step()
    mem("after optimizer step")
    badp = [n for n, p in tr.model.named_parameters() if not torch.isfinite(local(p.data)).all()]
    print(f"   committed {tr.manager.current_step()} non-finite params after step: {badp[:8]}", flush=True)
    losses.append(float(loss.detach()))
    assert not bad and not badp, (bad, badp)
assert losses[-1] < losses[0], losses  # same batch every step: the loss must go down
tr.shutdown()
lh.shutdown()
dist.destroy_process_group()
arena = getattr(tr, "_grad_arena", None)
if arena is not None:
    print(f"symmetric gradient arena: {arena._arena.numel() / 2**20:.1f} MiB, used {arena._off / 2**20:.1f} MiB, fallbacks {arena.fallbacks}")
    assert arena.fallbacks == 0 and arena._off > 0
print("HSDP_DEBUG ok")
