# one-GPU check of the diff3x3x1 work, most important first (the GPU budget may end the call early)
mkdir -p gpurun_out/d1
timeout 300 python -m pytest tests/test_gpu_diff.py tests/test_gpu_staging.py -x -q > gpurun_out/d1/pytest_diff.log 2>&1; tail -n 3 gpurun_out/d1/pytest_diff.log
timeout 200 python tools/bench_diff.py > gpurun_out/d1/bench_diff.log 2>&1; tail -n 6 gpurun_out/d1/bench_diff.log
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/d1/pytest.log 2>&1; tail -n 3 gpurun_out/d1/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/d1/smoke.log 2>&1; tail -n 1 gpurun_out/d1/smoke.log
python bench.py > gpurun_out/d1/bench_cfg2_full.log 2> gpurun_out/d1/bench_cfg2_full.err; tail -n 1 gpurun_out/d1/bench_cfg2_full.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['encode_gbs'], d['decode_gbs'], d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline']['value'])"
timeout 120 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:diff_kernel -c 40 --csv --log-file gpurun_out/d1/diff_ncu.csv python tools/bench_diff.py 64x2048x2048 > gpurun_out/d1/diff_ncu.log 2>&1; tail -n 2 gpurun_out/d1/diff_ncu.log | head -c 300
