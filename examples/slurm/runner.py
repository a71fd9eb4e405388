"""Keep N replica groups of a training job alive on a Slurm cluster (or locally).

    python examples/slurm/runner.py --replicas 8 --gpus-per-replica 1 --script bench.py -- --gpus 1

Every ``--check-every`` seconds the runner lists its job steps and re-submits any replica group
that died, which is all the "scheduler integration" the protocol needs: a restarted group gets
a fresh uuid, joins the next quorum and heals live from a peer (reference: examples/slurm/runner.py:
118-149, which does the same for torchtitan llama3_8b replicas). Without ``sbatch`` on PATH it
falls back to local subprocesses through ``torchft_b200.launcher``.
"""

from __future__ import annotations

import argparse
import os
import shutil
import subprocess
import sys
import time
from typing import Dict, List

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def sbatch_cmd(rid: int, a: argparse.Namespace, extra: List[str]) -> List[str]:
    env = (f"REPLICA_GROUP_ID={rid},NUM_REPLICA_GROUPS={a.replicas},TORCHFT_LIGHTHOUSE={a.lighthouse},"
           "TORCH_NCCL_ASYNC_ERROR_HANDLING=1")
    return ["sbatch", "--parsable", f"--job-name={a.job_name}_{rid}", f"--gpus={a.gpus_per_replica}", "--nodes=1",
            f"--export=ALL,{env}", "--wrap",
            " ".join([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc_per_node={a.gpus_per_replica}",
                      "--master_addr=127.0.0.1", f"--master_port={29600 + rid}", a.script, *extra])]


def alive_jobs(job_name: str) -> Dict[int, str]:
    out = subprocess.run(["squeue", "--me", "--noheader", "--format=%j %i"], capture_output=True, text=True).stdout
    res = {}
    for line in out.splitlines():
        name, jid = line.split()
        if name.startswith(job_name + "_"):
            res[int(name.rsplit("_", 1)[1])] = jid
    return res


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--replicas", type=int, default=2)
    ap.add_argument("--gpus-per-replica", type=int, default=1)
    ap.add_argument("--script", default="train_ddp.py")
    ap.add_argument("--lighthouse", default=os.environ.get("TORCHFT_LIGHTHOUSE", "http://127.0.0.1:29510"))
    ap.add_argument("--job-name", default="torchft_b200")
    ap.add_argument("--check-every", type=float, default=10.0)
    ap.add_argument("extra", nargs="*")
    a = ap.parse_args()
    if shutil.which("sbatch") is None:
        from torchft_b200.launcher import hsdp, launch_local

        print("sbatch not found: running replica groups locally with relaunch")
        raise SystemExit(launch_local(hsdp(*a.extra, replicas=a.replicas, workers_per_replica=a.gpus_per_replica,
                                           script=a.script, lighthouse=a.lighthouse), relaunch=True))
    while True:
        alive = alive_jobs(a.job_name)
        for rid in range(a.replicas):
            if rid not in alive:
                jid = subprocess.run(sbatch_cmd(rid, a, a.extra), capture_output=True, text=True).stdout.strip()
                print(f"(re)launched replica group {rid} as job {jid}")
        time.sleep(a.check_every)


if __name__ == "__main__":
    main()
