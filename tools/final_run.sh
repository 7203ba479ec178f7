# round-end measurements on ONE GPU: tests, every bench workload, profiles, the side benches
mkdir -p gpurun_out/f1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/f1/pytest.log 2>&1; tail -n 4 gpurun_out/f1/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f1/smoke.log 2>&1; tail -n 2 gpurun_out/f1/smoke.log
python bench.py > gpurun_out/f1/bench_cfg2_full.log 2> gpurun_out/f1/bench_cfg2_full.err; tail -n 1 gpurun_out/f1/bench_cfg2_full.log | head -c 400; echo
for w in cfg1 cfg3 cfg5; do python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/f1/bench_$w.log 2>&1; tail -n 1 gpurun_out/f1/bench_$w.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['config']['workload'][:5], d['value'], d['encode_gbs'], d['decode_gbs'], d['compression_ratio'], d['e2e'])"; done
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/f1/bench_ref.log 2>&1; tail -n 1 gpurun_out/f1/bench_ref.log | head -c 600; echo
bash tools/profile_bench.sh cfg2 > gpurun_out/f1/profile.log 2>&1; tail -n 2 gpurun_out/f1/profile.log
timeout 600 python tools/bench_pageable.py > gpurun_out/f1/pageable.log 2>&1; tail -n 5 gpurun_out/f1/pageable.log
timeout 600 python tools/bench_cfg5_batch.py 32 > gpurun_out/f1/cfg5_batch.log 2>&1; tail -n 1 gpurun_out/f1/cfg5_batch.log
timeout 900 python tools/bench_cfg4_batch.py 256x2048x2048 8 > gpurun_out/f1/cfg4_batch.log 2>&1; tail -n 1 gpurun_out/f1/cfg4_batch.log
timeout 900 python tools/bench_cfg4_batch.py 128x2048x2048 8 serial > gpurun_out/f1/cfg4_batch_serial.log 2>&1; tail -n 1 gpurun_out/f1/cfg4_batch_serial.log
timeout 900 python tools/bench_cfg4.py 256x2048x2048 > gpurun_out/f1/cfg4_single.log 2>&1; tail -n 2 gpurun_out/f1/cfg4_single.log
timeout 600 python tools/bench_bitshuffle.py > gpurun_out/f1/bitshuffle.log 2>&1; tail -n 7 gpurun_out/f1/bitshuffle.log
