#!/bin/bash
# 8-GPU run: flagship scaling point in IPC (default) and VMM+NVLS mode, NCCL-equivalent arm, DiLoCo sync at 8 replicas.
N=${1:-8}
mkdir -p gpurun_out
run_bench() {  # name, extra env...
  local name=$1; shift
  echo "=== bench $N gpu $name"
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29540 + RANDOM % 50)) bench.py --gpus $N --steps 5 --warmup 3 $BENCH_ARGS > gpurun_out/bench_n${N}_$name.log 2>&1
  echo "rc=$?"; grep '^{"metric' gpurun_out/bench_n${N}_$name.log | cut -c1-360; grep -i -E "Traceback|Error" gpurun_out/bench_n${N}_$name.log | head -3
}
BENCH_ARGS="" run_bench ipc TORCHFT_B200_SYMM=ipc
BENCH_ARGS="" run_bench vmm TORCHFT_B200_SYMM=vmm
BENCH_ARGS="--impl nccl" run_bench nccl TORCHFT_B200_SYMM=ipc
echo "=== diloco $N gpus (llama3_1b, bf16 pseudo-gradients over NVLS)"; TORCHFT_B200_SYMM=vmm timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29591 bench/diloco_bench.py --sync-every 10 --outer-steps 3 --out gpurun_out/diloco_bench_w$N.json > gpurun_out/diloco_bench_w$N.log 2>&1; echo "rc=$?"; grep DILOCO_BENCH gpurun_out/diloco_bench_w$N.log | cut -c1-500
echo "=== diloco $N gpus fused fp8"; TORCHFT_B200_SYMM=vmm timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29592 bench/diloco_bench.py --sync-every 10 --outer-steps 3 --quantize --out gpurun_out/diloco_bench_w${N}_q8.json > gpurun_out/diloco_bench_w${N}_q8.log 2>&1; echo "rc=$?"; grep DILOCO_BENCH gpurun_out/diloco_bench_w${N}_q8.log | cut -c1-500
