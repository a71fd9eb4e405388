#!/bin/bash
# 8-GPU run: flagship bench first (scaling point), then comm sweep in VMM mode (P2P + NVLS + NCCL columns),
# NCCL-equivalent bench, and the W=8 launch-plan tuning sweep if time remains.
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/gpus_topo_w$N.txt 2>&1
echo "=== bench $N gpu native"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/bench_n$N.log 2>&1; echo "rc=$?"; grep '^{"metric' gpurun_out/bench_n$N.log | cut -c1-400
echo "=== comm bench $N gpus (vmm: P2P + NVLS)"; TORCHFT_B200_SYMM=vmm timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench/comm_bench.py --max-mb 1024 --blocks 32,64 --out gpurun_out/comm_bench_vmm_w$N.json > gpurun_out/comm_bench_vmm_w$N.log 2>&1; echo "comm rc=$?"; grep -E "COMM_BENCH|FAILED" gpurun_out/comm_bench_vmm_w$N.log | cut -c1-600
python - <<PY
import json
try:
    r=json.load(open('gpurun_out/comm_bench_vmm_w$N.json'))
    print('mode', r.get('symm_mode'), 'nvls', r.get('nvls'), 'all_ok', r['all_ok'])
    for row in r['sweep']:
        print(row['bytes'], {k.replace('native_','').replace('_ms',''):v for k,v in row.items() if k.endswith('_ms')})
except Exception as e:
    print('no comm json', e)
PY
tail -5 gpurun_out/comm_bench_vmm_w$N.log | cut -c1-300
echo "=== multi-GPU tests (HSDP/FSDP2 4 GPUs, quantized collectives + Baby NCCL 2 GPUs, FT step)"; timeout 600 python -m pytest tests/test_hsdp_gpu.py tests/test_ft_gpu.py tests/test_kernels_gpu.py -m gpu -q --timeout 500 > gpurun_out/pytest_gpu_multi.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu_multi.log | cut -c1-300
echo "=== bench $N gpu nccl-equivalent"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 6 --warmup 3 --impl nccl > gpurun_out/bench_n${N}_nccl.log 2>&1; echo "rc=$?"; grep '^{"metric' gpurun_out/bench_n${N}_nccl.log | cut -c1-400
echo "=== comm tune $N gpus"; timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench/comm_tune.py --out gpurun_out/comm_tune_w$N.json > gpurun_out/comm_tune_w$N.log 2>&1; echo "rc=$?"; grep '^{"bytes' gpurun_out/comm_tune_w$N.log | cut -c1-300
