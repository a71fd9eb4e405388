#!/bin/bash
# Run 3: fixed barrier fences (comm sweep), comm-kernel SM budget A/B in the trainer, kernel rooflines + ncu, kill/rejoin heal.
mkdir -p gpurun_out
echo "=== comm bench 2 gpus"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench/comm_bench.py --max-mb 1024 --out gpurun_out/comm_bench_run3.json > gpurun_out/comm2_run3.log 2>&1; echo "comm rc=$?"; grep -E "COMM_BENCH|FAILED" gpurun_out/comm2_run3.log
python - <<'PY'
import json
r=json.load(open('gpurun_out/comm_bench_run3.json'))
for row in r['sweep']:
    print(row['bytes'], {k:v for k,v in row.items() if k.endswith('_ms')})
PY
echo "=== kernel micro"; timeout 600 python bench/kernel_micro.py --out gpurun_out/kernel_micro.json > gpurun_out/kernel_micro.log 2>&1; echo "rc=$?"; cat gpurun_out/kernel_micro.log | tail -20
echo "=== ncu adamw + xent + rmsnorm_bwd"
for k in adamw xent rmsnorm_bwd swiglu_bwd; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -s 3 -f -o gpurun_out/prof_$k python bench/kernel_micro.py --only ${k%%_*} --iters 1 > gpurun_out/ncu_$k.log 2>&1; echo "ncu $k rc=$?"
done
echo "=== heal bench (kill/rejoin) 8B"; timeout 900 python bench/heal_bench.py --gpus 2 --model llama3_8b --kill-at 5 --rejoin-at 9 --steps 24 --out gpurun_out/heal_bench.json > gpurun_out/heal_bench.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/heal_bench.log; tail -5 gpurun_out/heal_replica0.jsonl.out; tail -5 gpurun_out/heal_replica1.jsonl.out
for b in 16 32; do
echo "=== bench 2 gpu native AR_BLOCKS=$b"; TORCHFT_B200_AR_BLOCKS=$b timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2954$((b/16)) bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2_b$b.log 2>&1; echo "rc=$?"; grep -o '"value": [0-9.]*, "unit"\|"ms_per_step": [0-9.]*' gpurun_out/bench_n2_b$b.log | head -3
done
