# last check of a round on ONE GPU: all GPU tests, smoke, the default bench line, the reference arm, the profile captures
mkdir -p gpurun_out/f2
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/f2/pytest.log 2>&1; tail -n 3 gpurun_out/f2/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f2/smoke.log 2>&1; tail -n 1 gpurun_out/f2/smoke.log
python bench.py > gpurun_out/f2/bench_cfg2_full.log 2> gpurun_out/f2/bench_cfg2_full.err; tail -n 1 gpurun_out/f2/bench_cfg2_full.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['encode_gbs'], d['decode_gbs'], d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline']['value'])"
bash tools/profile_bench.sh cfg2 > gpurun_out/f2/profile.log 2>&1; tail -n 1 gpurun_out/f2/profile.log | head -c 200; echo
timeout 200 python tools/bench_diff.py > gpurun_out/f2/bench_diff.log 2>&1; head -n 1 gpurun_out/f2/bench_diff.log
