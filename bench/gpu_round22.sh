#!/bin/bash
mkdir -p gpurun_out
echo "=== bench 2 gpu --quantize (fused fp8 gradient all-reduce)"; timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29547 bench.py --gpus 2 --steps 5 --warmup 3 --quantize > gpurun_out/bench_n2_q8.log 2>&1; echo "rc=$?"; grep '^{"metric' gpurun_out/bench_n2_q8.log | cut -c1-330; grep -i -E "Traceback|Error" gpurun_out/bench_n2_q8.log | head -5
