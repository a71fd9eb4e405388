#!/bin/bash
mkdir -p gpurun_out
echo "=== comm tune 2 gpus (no host reads on wait path)"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench/comm_tune.py --modes 0,2 --out gpurun_out/comm_tune_w2_b.json > gpurun_out/comm_tune_w2_b.log 2>&1; echo "rc=$?"; grep '^{"bytes' gpurun_out/comm_tune_w2_b.log | cut -c1-330
