#!/bin/bash
# compute-sanitizer over the single-GPU kernel tests (memcheck, then racecheck if time remains).
mkdir -p gpurun_out
SEL="rmsnorm or swiglu or rope or cross_entropy or adamw or q8_roundtrip or heal_copy or fused_block_ops"
echo "=== memcheck"; timeout 140 compute-sanitizer --tool memcheck --error-exitcode 1 --launch-timeout 100 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "$SEL" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|Error" gpurun_out/sanitizer_memcheck.log | head -8
echo "=== racecheck"; timeout 80 compute-sanitizer --tool racecheck --error-exitcode 1 --launch-timeout 100 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "rmsnorm or swiglu or cross_entropy or q8_roundtrip" > gpurun_out/sanitizer_racecheck.log 2>&1; echo "rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard|Error" gpurun_out/sanitizer_racecheck.log | head -8
