#!/bin/bash
# round-end run on ONE GPU: all GPU tests, smoke, the default bench line, the reference arm, the profile captures
D=gpurun_out/${1:-fin}
mkdir -p $D
timeout 1500 python -m pytest tests -m gpu -q > $D/pytest.log 2>&1; tail -n 4 $D/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $D/smoke.log 2>&1; tail -n 1 $D/smoke.log
python bench.py > $D/bench_default.log 2> $D/bench_default.err; tail -n 1 $D/bench_default.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['encode_gbs'], d['decode_gbs'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['stage_ms'], d['cpu_baseline']['value'], d['compression_ratio'])"
python bench.py --impl reference --steps 2 --warmup 1 > $D/bench_ref.log 2>&1; tail -n 1 $D/bench_ref.log | head -c 300; echo
bash tools/profile_bench.sh cfg2 > $D/profile.log 2>&1; tail -n 1 $D/profile.log | head -c 200; echo
