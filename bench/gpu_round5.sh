#!/bin/bash
# Run 5: barrier/algo/grid tuning sweep (2 GPUs), heal bench with the faster copy kernel.
mkdir -p gpurun_out
echo "=== comm tune 2 gpus"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench/comm_tune.py --out gpurun_out/comm_tune_w2.json > gpurun_out/comm_tune_w2.log 2>&1; echo "rc=$?"; grep '^{"bytes' gpurun_out/comm_tune_w2.log
echo "=== heal bench (kill/rejoin) 8B"; timeout 1200 python bench/heal_bench.py --gpus 2 --model llama3_8b --kill-at 6 --rejoin-at 10 --steps 26 --out gpurun_out/heal_bench.json > gpurun_out/heal_bench.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/heal_bench.log
