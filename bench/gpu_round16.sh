#!/bin/bash
# 1-GPU sanity of the round-end sequence (GPU test tier, smoke(), bench) + fresh ncu captures of current kernels.
mkdir -p gpurun_out
echo "=== pytest gpu"; timeout 900 python -m pytest tests -x -q -m gpu --timeout 600 > gpurun_out/pytest_gpu16.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu16.log | cut -c1-300
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > gpurun_out/smoke16.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/smoke16.log | cut -c1-300
echo "=== bench default flags"; timeout 600 python bench.py > gpurun_out/bench_default16.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_default16.log | cut -c1-330
echo "=== ncu rmsnorm_bwd / swiglu_bwd / xent / q8 (current kernels)"
for k in rmsnorm_bwd swiglu_bwd xent; do
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -s 3 -f -o gpurun_out/prof16_$k python bench/kernel_micro.py --only ${k%%_*} --iters 1 > gpurun_out/ncu16_$k.log 2>&1; echo "ncu $k rc=$?"
done
