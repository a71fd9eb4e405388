#!/bin/bash
# 2-GPU run: checkpoint-transport shoot-out (NVLink P2P heal vs HTTP vs PG/NCCL send-recv).
mkdir -p gpurun_out
echo "=== transport bench"; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 bench/transport_bench.py --total-gb 6 --out gpurun_out/transport_bench.json > gpurun_out/transport_bench.log 2>&1; echo "rc=$?"; grep -E "TRANSPORT_BENCH|Error|Traceback" gpurun_out/transport_bench.log | cut -c1-1500 | head -5; tail -3 gpurun_out/transport_bench.log | cut -c1-300
