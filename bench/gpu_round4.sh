#!/bin/bash
# Run 4: barrier with <=2 system fences per CTA, hoisted AdamW loads, in-place heal + fail-fast.
mkdir -p gpurun_out
echo "=== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu4.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu4.log
echo "=== comm bench 2 gpus"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench/comm_bench.py --max-mb 1024 --out gpurun_out/comm_bench_run4.json > gpurun_out/comm2_run4.log 2>&1; echo "comm rc=$?"; grep -E "COMM_BENCH|FAILED" gpurun_out/comm2_run4.log
python - <<'PY'
import json
r=json.load(open('gpurun_out/comm_bench_run4.json'))
for row in r['sweep']:
    print(row['bytes'], {k:v for k,v in row.items() if k.endswith('_ms')})
PY
echo "=== kernel micro"; timeout 600 python bench/kernel_micro.py --only adamw --out gpurun_out/kernel_micro_adamw.json 2>&1 | tail -3
echo "=== heal bench (kill/rejoin) 8B"; timeout 1200 python bench/heal_bench.py --gpus 2 --model llama3_8b --kill-at 6 --rejoin-at 10 --steps 26 --out gpurun_out/heal_bench.json > gpurun_out/heal_bench.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/heal_bench.log; grep -c . gpurun_out/heal_replica0.jsonl gpurun_out/heal_replica1.jsonl; grep -i "error\|Traceback" gpurun_out/heal_replica*.jsonl.out | head -10
