mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "=== heal_bench n2 (liveness abort, op timeout 60 s, heartbeat 2 s)"
timeout 420 python bench/heal_bench.py --gpus 2 --model llama3_8b --kill-at 6 --rejoin-at 14 --steps 60 --timeout 60 --heartbeat-timeout-ms 2000 --out gpurun_out/heal_bench_z1_n2.json > gpurun_out/heal_n2.log 2>&1; tail -4 gpurun_out/heal_n2.log
echo "=== nvlink bytes"
timeout 200 $TR --master-port 29533 bench/nvlink_bytes.py --mb 512 --out gpurun_out/nvlink_bytes_n2.json > gpurun_out/nvlink_n2.log 2>&1; grep -v Warning gpurun_out/nvlink_n2.log | tail -9
echo "=== vmm/nvls check"
TORCHFT_B200_SYMM=vmm timeout 200 $TR --master-port 29544 scripts/gpu/zero1_nvls_check.py --out gpurun_out/zero1_nvls_n2.json > gpurun_out/nvls_n2.log 2>&1; grep -v Warning gpurun_out/nvls_n2.log | tail -6
echo "=== diloco 1b n2 quantized"
timeout 300 $TR --master-port 29555 bench/diloco_bench.py --model llama3_1b --sync-every 10 --outer-steps 3 --quantize --out gpurun_out/diloco_1b_n2.json > gpurun_out/diloco_n2.log 2>&1; grep -v Warning gpurun_out/diloco_n2.log | tail -4
echo "=== hsdp 1 group x 2 shards, 8B"
timeout 420 $TR --master-port 29566 bench.py --gpus 2 --shards 2 --steps 4 --warmup 3 --no-baseline-arm > gpurun_out/hsdp_n2.log 2>&1; grep -v Warning gpurun_out/hsdp_n2.log | tail -5
