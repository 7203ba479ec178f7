#!/bin/bash
# round-2 check on ONE GPU: all GPU tests (no -x so one failure does not hide the rest), smoke, default bench line, cfg1/cfg3 lines
D=gpurun_out/${1:-r2a}
mkdir -p $D
timeout 1500 python -m pytest tests -m gpu -q > $D/pytest.log 2>&1; tail -n 15 $D/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $D/smoke.log 2>&1; tail -n 2 $D/smoke.log
python bench.py > $D/bench_cfg2_full.log 2> $D/bench_cfg2_full.err; tail -n 1 $D/bench_cfg2_full.log | head -c 3000; echo
for w in cfg1 cfg3; do python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > $D/bench_$w.log 2>&1; tail -n 1 $D/bench_$w.log | head -c 1500; echo; done
