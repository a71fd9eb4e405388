#!/bin/bash
# 1-GPU run: GPU test tier, AdamW/GEMM overlap probe, flagship bench A/B (overlapped vs single-launch AdamW),
# fresh ncu capture of the current AdamW kernel, DiLoCo outer-step bench.
mkdir -p gpurun_out
echo "=== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu10.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu10.log
echo "=== overlap probe"; timeout 300 python bench/overlap_probe.py > gpurun_out/overlap_probe.log 2>&1; echo "rc=$?"; grep OVERLAP_PROBE gpurun_out/overlap_probe.log | cut -c1-1500
echo "=== bench 1 gpu (overlapped AdamW, default)"; timeout 600 python bench.py --gpus 1 --steps 6 --warmup 3 > gpurun_out/bench_n1_ovl.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_n1_ovl.log | cut -c1-330
echo "=== bench 1 gpu (overlapped AdamW, 296 CTAs)"; TORCHFT_B200_OPT_BLOCKS=296 timeout 600 python bench.py --gpus 1 --steps 6 --warmup 3 > gpurun_out/bench_n1_ovl296.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_n1_ovl296.log | cut -c1-330
echo "=== bench 1 gpu (single-launch AdamW)"; TORCHFT_B200_OVERLAP_OPT=0 timeout 600 python bench.py --gpus 1 --steps 6 --warmup 3 > gpurun_out/bench_n1_noovl.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_n1_noovl.log | cut -c1-330
echo "=== W=8 collectives, 8 ranks oversubscribed on this GPU"; timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 bench/comm_oversub.py > gpurun_out/comm_oversub_w8.log 2>&1; echo "rc=$?"; grep -E "COMM_OVERSUB|FAIL|Error" gpurun_out/comm_oversub_w8.log | head -8 | cut -c1-600
echo "=== ncu adamw (current kernel)"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:adamw -c 1 -s 3 -f -o gpurun_out/prof_adamw_v2 python bench/kernel_micro.py --only adamw --iters 1 > gpurun_out/ncu_adamw_v2.log 2>&1; echo "ncu rc=$?"
echo "=== diloco bench"; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29561 bench/diloco_bench.py > gpurun_out/diloco_bench.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/diloco_bench.log | cut -c1-400
