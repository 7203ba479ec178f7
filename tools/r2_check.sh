#!/bin/bash
# round-2 check on ONE GPU: all GPU tests (no -x so one failure does not hide the rest), smoke, default bench line, reference arm
D=gpurun_out/${1:-r2a}
mkdir -p $D
timeout 1500 python -m pytest tests -m gpu -q > $D/pytest.log 2>&1; tail -n 15 $D/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $D/smoke.log 2>&1; tail -n 2 $D/smoke.log
( time python bench.py ) > $D/bench_default.log 2> $D/bench_default.err; tail -n 1 $D/bench_default.log | head -c 6000; echo; tail -n 4 $D/bench_default.err
python bench.py --impl reference --steps 2 --warmup 1 > $D/bench_ref.log 2>&1; tail -n 1 $D/bench_ref.log | head -c 900; echo
