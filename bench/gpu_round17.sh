#!/bin/bash
# 2-GPU chaos soak: ProcessGroupB200 data plane + in-place NVLink heal under injected failures.
mkdir -p gpurun_out
echo "=== chaos soak 2 gpus"; timeout 420 python bench/chaos_soak.py --steps 6000 --mtbf-secs 8 --failures kill_proc,segfault,comms,kill_group --timeout 380 --out gpurun_out/chaos_soak_2gpu.json > gpurun_out/chaos_soak_2gpu.log 2>&1; echo "rc=$?"; grep CHAOS_SOAK gpurun_out/chaos_soak_2gpu.log | cut -c1-1200; tail -30 gpurun_out/chaos_soak_2gpu.log | cut -c1-250
nvidia-smi --query-gpu=index,memory.used --format=csv | head -4
