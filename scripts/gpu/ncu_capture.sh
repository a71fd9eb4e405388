#!/bin/bash
# ONE-GPU ncu captures of the multi-rank kernels (ranks emulated in-process, flags pre-signalled, one kernel at a time:
# "peer" accesses hit local HBM, so these show instruction mix / occupancy / memory-pipe behaviour, not NVLink) plus the
# single-GPU kernels. Run under gpurun; reports land in gpurun_out/ncu/, summaries are made by scripts/ncu_summarize.py.
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/ncu
NCU="ncu --set full --clock-control none --import-source on"
# FT-ZeRO-1 big kernels at W=8 (rank 0's launches: reduce over a 64 MB unit, update of its 2 held slices)
$NCU -k regex:'zero1_(reduce|update)_kernel' -s 6 -c 2 -o gpurun_out/ncu/zero1_w8 -f \
  python bench/zero1_bench.py --virtual 8 --mb 256 --iters 1 --warmup 1 > gpurun_out/ncu/zero1_w8.log 2>&1
# world-1 update (= the flagship's AdamW at N=1), DiLoCo outer step, both heal-copy variants
$NCU -k regex:'zero1_update_kernel|diloco_outer_kernel|heal_copy' -c 12 -o gpurun_out/ncu/micro -f \
  python bench/kernel_micro.py --only zero1_update --iters 1 > gpurun_out/ncu/micro_a.log 2>&1
$NCU -k regex:'diloco_outer_kernel' -s 3 -c 1 -o gpurun_out/ncu/diloco_outer -f \
  python bench/kernel_micro.py --only diloco_outer --iters 1 > gpurun_out/ncu/micro_b.log 2>&1
$NCU -k regex:'heal_copy' -s 3 -c 6 -o gpurun_out/ncu/heal_copy -f \
  python bench/kernel_micro.py --only heal_copy --iters 1 > gpurun_out/ncu/micro_c.log 2>&1
# every collective kernel of the self-check (push exchange is not in it: use the collectives test harness)
$NCU -k regex:'allreduce_|q8_allreduce|zero1_commit|zero1_handshake' -c 20 -o gpurun_out/ncu/selfcheck -f \
  python -c "import torch; torch.cuda.set_device(0); from torchft_b200.bench_utils import collectives_selfcheck as c; print(c(world=8, nelem=1<<22))" > gpurun_out/ncu/selfcheck.log 2>&1
# launch list of smoke() as the driver sees it
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/ncu/smoke_launches.csv \
  python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/ncu/smoke.log 2>&1
ls -la gpurun_out/ncu
