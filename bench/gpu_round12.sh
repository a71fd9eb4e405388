#!/bin/bash
# 4-GPU run: multi-GPU test tier (HSDP/FSDP2 on 4 GPUs, quantized collectives + Baby NCCL on 2, comm correctness),
# then VMM-mode flagship bench on 2 GPUs (validates 16 GB VMM segments + multicast binding under the trainer).
mkdir -p gpurun_out
echo "=== pytest multi-gpu"; timeout 600 python -m pytest tests/test_hsdp_gpu.py tests/test_ft_gpu.py -m gpu -q --timeout 500 > gpurun_out/pytest_gpu12.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu12.log | cut -c1-300
echo "=== bench 2 gpu VMM mode"; TORCHFT_B200_SYMM=vmm timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/bench_n2_vmm.log 2>&1; echo "rc=$?"; grep '^{"metric' gpurun_out/bench_n2_vmm.log | cut -c1-330; grep -i -E "error|Traceback" gpurun_out/bench_n2_vmm.log | head -5
