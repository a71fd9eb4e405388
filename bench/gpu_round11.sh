#!/bin/bash
# 1-GPU run: full GPU test tier (new tests), flagship bench with residual-epilogue fusion, AdamW CTA-cap sweep.
mkdir -p gpurun_out
echo "=== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu11.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu11.log | cut -c1-300
echo "=== bench 1 gpu"; timeout 600 python bench.py --gpus 1 --steps 6 --warmup 3 > gpurun_out/bench_n1_r11.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_n1_r11.log | cut -c1-330
echo "=== kernel micro adamw caps"; timeout 300 python bench/kernel_micro.py --only adamw --out gpurun_out/kernel_micro_adamw_caps.json > gpurun_out/kernel_micro_adamw_caps.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/kernel_micro_adamw_caps.log | cut -c1-200
