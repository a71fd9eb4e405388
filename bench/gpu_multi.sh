#!/bin/bash
# Multi-GPU run: N=$1. Tuning sweep, comm bench vs NCCL, flagship bench (native + nccl-equivalent).
N=${1:-4}
mkdir -p gpurun_out
echo "=== comm tune $N gpus"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench/comm_tune.py --out gpurun_out/comm_tune_w$N.json > gpurun_out/comm_tune_w$N.log 2>&1; echo "rc=$?"; grep '^{"bytes' gpurun_out/comm_tune_w$N.log | cut -c1-260
echo "=== comm bench $N gpus"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench/comm_bench.py --max-mb 1024 --out gpurun_out/comm_bench_w$N.json > gpurun_out/comm_bench_w$N.log 2>&1; echo "comm rc=$?"; grep -E "COMM_BENCH|FAILED" gpurun_out/comm_bench_w$N.log
python - <<PY
import json
r=json.load(open('gpurun_out/comm_bench_w$N.json'))
for row in r['sweep']:
    print(row['bytes'], {k:v for k,v in row.items() if k.endswith('_ms')})
PY
echo "=== bench $N gpu native"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/bench_n$N.log 2>&1; echo "rc=$?"; grep '^{"metric' gpurun_out/bench_n$N.log | cut -c1-300
echo "=== bench $N gpu nccl-equivalent"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 6 --warmup 3 --impl nccl > gpurun_out/bench_n${N}_nccl.log 2>&1; echo "rc=$?"; grep '^{"metric' gpurun_out/bench_n${N}_nccl.log | cut -c1-300
