#!/bin/bash
# Final 1-GPU sanity of the round-end sequence on the final tree + ncu of the heal-copy and fp8 kernels.
mkdir -p gpurun_out
echo "=== pytest gpu"; timeout 600 python -m pytest tests -x -q -m gpu --timeout 500 > gpurun_out/pytest_gpu19.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu19.log | cut -c1-300
echo "=== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > gpurun_out/smoke19.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/smoke19.log | cut -c1-200
echo "=== bench default flags"; timeout 400 python bench.py > gpurun_out/bench_default19.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_default19.log | cut -c1-330
echo "=== ncu heal_copy / q8"
for k in heal_copy q8_quantize q8_dequantize; do
  timeout 150 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -s 2 -f -o gpurun_out/prof19_$k python bench/kernel_micro.py --only ${k%%_*} --iters 1 > gpurun_out/ncu19_$k.log 2>&1; echo "ncu $k rc=$?"
done
