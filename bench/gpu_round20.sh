#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest fused block ops"; timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "fused_block_ops or p2p_transport" > gpurun_out/pytest_gpu20.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/pytest_gpu20.log | cut -c1-300
