"""Debug aid: HSDPTrainer on ONE GPU (1 group x 1 shard): exercises FSDP2's hooks against the fused model ops."""
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
lh = local_lighthouse()
tr = HSDPTrainer(model, loopback(lh.address()), shards=1, backend=sys.argv[2] if len(sys.argv) > 2 else "b200", timeout=timedelta(seconds=30))
cfg = tr.cfg
S = min(cfg.max_seq_len, 512)
tok = torch.randint(0, cfg.vocab_size, (2, S), device="cuda")
tgt = torch.randint(0, cfg.vocab_size, (2, S), device="cuda")
for i in range(3):
    loss = tr.step_device(tok, tgt)
    torch.cuda.synchronize()
    print("step", i, float(loss), "committed", tr.manager.current_step(), flush=True)
tr.shutdown()
lh.shutdown()
dist.destroy_process_group()
print("HSDP_DEBUG ok")
