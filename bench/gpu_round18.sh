#!/bin/bash
# 2-GPU: full GPU test tier on the final tree + chaos soak with the key-path heal fix.
mkdir -p gpurun_out
echo "=== pytest gpu (2-GPU box)"; timeout 600 python -m pytest tests -x -q -m gpu --timeout 500 > gpurun_out/pytest_gpu18.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu18.log | cut -c1-300
echo "=== chaos soak 2 gpus"; timeout 420 python bench/chaos_soak.py --steps 6000 --mtbf-secs 8 --failures kill_proc,segfault,comms,kill_group --timeout 380 --out gpurun_out/chaos_soak_2gpu.json > gpurun_out/chaos_soak_2gpu.log 2>&1; echo "rc=$?"; grep CHAOS_SOAK gpurun_out/chaos_soak_2gpu.log | cut -c1-1200; grep -c "healing required" gpurun_out/chaos_soak_2gpu.log; tail -12 gpurun_out/chaos_soak_2gpu.log | cut -c1-250
