#!/bin/bash
# Run under gpurun on ONE GPU: plain bench first (must exit 0), then the ncu launch list of the same command and one
# `--set full` capture of the dominant kernel. Outputs land in gpurun_out/ (copy the summaries into profiles/).
set -u
W=${1:-cfg2}
CMD="python bench.py --workload $W --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-per-config"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_bench_$W.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_bench_$W.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$W.csv $CMD > gpurun_out/ncu_launches_$W.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lz4_encode_kernel -s 3 -c 1 -o gpurun_out/full_enc_$W $CMD > gpurun_out/ncu_full_enc_$W.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lz4_decode_kernel -s 3 -c 1 -o gpurun_out/full_dec_$W $CMD > gpurun_out/ncu_full_dec_$W.log 2>&1
ncu --set full --clock-control none -k regex:bitswap_encode_fast -s 3 -c 1 -o gpurun_out/full_swap_$W $CMD > gpurun_out/ncu_full_swap_$W.log 2>&1
tail -c 600 gpurun_out/plain_bench_$W.log
