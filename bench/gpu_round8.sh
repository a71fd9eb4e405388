#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu8.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu8.log
echo "=== VMM+NVLS comm bench 2 gpus"; TORCHFT_B200_SYMM=vmm timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench/comm_bench.py --max-mb 1024 --out gpurun_out/comm_bench_vmm_w2.json > gpurun_out/comm_bench_vmm_w2.log 2>&1; echo "comm rc=$?"; grep -E "COMM_BENCH|FAILED|Error|error" gpurun_out/comm_bench_vmm_w2.log | head -8
python - <<'PY'
import json
try:
    r=json.load(open('gpurun_out/comm_bench_vmm_w2.json'))
    print('mode', r.get('symm_mode'), 'nvls', r.get('nvls'), 'all_ok', r.get('all_ok'))
    for row in r['sweep']:
        print(row['bytes'], {k:v for k,v in row.items() if k.endswith('_ms')})
except Exception as e: print('no json', e)
PY
echo "=== step profile + sdpa shoot-out"; timeout 600 python bench/step_profile.py --out gpurun_out/step_profile.txt > gpurun_out/step_profile.log 2>&1; echo "rc=$?"; head -60 gpurun_out/step_profile.txt | cut -c1-200
echo "=== bench 1 gpu"; timeout 900 python bench.py --gpus 1 --steps 6 --warmup 3 > gpurun_out/bench_n1_b.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_n1_b.log | cut -c1-400
