#!/bin/bash
# 2-GPU A/B: CTA budget of the overlapped gradient all-reduce (same box, back to back).
mkdir -p gpurun_out
for b in 64 32; do
echo "=== bench 2 gpu AR_BLOCKS=$b"; TORCHFT_B200_AR_BLOCKS=$b timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2955$((b/32)) bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/bench_n2_arb$b.log 2>&1; echo "rc=$?"; grep -o '"value": [0-9.]*, "unit"\|"ms_per_step": [0-9.]*\|"sm_mhz": [0-9]*' gpurun_out/bench_n2_arb$b.log | head -4 | tr '\n' ' '; echo
done
