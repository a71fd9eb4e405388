#!/bin/bash
# First GPU contact: kernel numerics, 8B step probe (1 GPU), 2-GPU comm bench.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt 2>&1
nvidia-smi topo -m >> gpurun_out/gpus.txt 2>&1
echo "=== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log
echo "=== model step 8B"; timeout 600 python bench/model_step.py --model llama3_8b --seq 8192 --batch 1 --steps 3 --warmup 2 --out gpurun_out/model_step_8b.json > gpurun_out/model_step.log 2>&1; echo "model rc=$?"; tail -5 gpurun_out/model_step.log
echo "=== comm bench 2 gpus"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench/comm_bench.py --max-mb 1024 --blocks 16,32,64,128 > gpurun_out/comm2.log 2>&1; echo "comm rc=$?"; tail -40 gpurun_out/comm2.log
