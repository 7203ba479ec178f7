# quick one-GPU look at the diff3x3x1 kernels: parity tests, device throughput, per-launch times
mkdir -p gpurun_out/diff
timeout 300 python -m pytest tests/test_gpu_diff.py -x -q > gpurun_out/diff/pytest_diff.log 2>&1; tail -n 2 gpurun_out/diff/pytest_diff.log
timeout 200 python tools/bench_diff.py > gpurun_out/diff/bench_diff.log 2>&1; head -n 1 gpurun_out/diff/bench_diff.log
timeout 120 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:diff_kernel -c 16 --csv --log-file gpurun_out/diff/diff_ncu.csv python tools/bench_diff.py 64x2048x2048 > gpurun_out/diff/diff_ncu.log 2>&1; grep -c diff_kernel gpurun_out/diff/diff_ncu.csv
