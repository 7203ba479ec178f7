#!/bin/bash
# Run under gpurun on ONE GPU: cfg3 (quantiser->lz4) plain bench first (must exit 0), then the ncu launch list of the same
# command and one `--set full` capture of the histogram kernel. Outputs land in gpurun_out/.
set -u
CMD="python bench.py --workload cfg3 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-per-config"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_bench_cfg3.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_bench_cfg3.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cfg3.csv $CMD > gpurun_out/ncu_launches_cfg3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:histogram_u16_kernel -s 2 -c 1 -f -o gpurun_out/full_hist_cfg3 $CMD > gpurun_out/ncu_full_hist_cfg3.log 2>&1
tail -c 400 gpurun_out/plain_bench_cfg3.log
