#!/bin/bash
# Run 2: FT trainer end to end. 1-GPU bench (8B), 2-GPU bench native vs nccl-equivalent, retuned comm sweep.
mkdir -p gpurun_out
echo "=== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu2.log
echo "=== bench 1 gpu"; timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/bench_n1.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/bench_n1.log
echo "=== bench 2 gpu native"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/bench_n2.log
echo "=== bench 2 gpu nccl-equivalent"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 5 --warmup 3 --impl nccl > gpurun_out/bench_n2_nccl.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/bench_n2_nccl.log
echo "=== comm bench 2 gpus"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench/comm_bench.py --max-mb 1024 --out gpurun_out/comm_bench_run2.json > gpurun_out/comm2_run2.log 2>&1; echo "comm rc=$?"; grep -E "COMM_BENCH|FAILED" gpurun_out/comm2_run2.log
python - <<'PY'
import json
r=json.load(open('gpurun_out/comm_bench_run2.json'))
for row in r['sweep']:
    print(row['bytes'], {k:v for k,v in row.items() if k.endswith('_ms')})
PY
