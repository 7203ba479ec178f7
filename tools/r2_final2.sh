#!/bin/bash
# round-end run on ONE GPU after a change of the LZ4 encoder / histogram only: all GPU tests, smoke, the default bench line,
# the reference arm, the ncu launch list and one full capture of the encoder (decoder and transposes: captures of r2_final.sh stand)
D=gpurun_out/${1:-fin3}
mkdir -p $D
timeout 1200 python -m pytest tests -m gpu -q -x > $D/pytest.log 2>&1; tail -n 3 $D/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $D/smoke.log 2>&1; tail -n 1 $D/smoke.log
python bench.py > $D/bench_default.log 2> $D/bench_default.err; tail -n 1 $D/bench_default.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['encode_gbs'], d['decode_gbs'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['stage_ms'], d['cpu_baseline']['value'], d['compression_ratio']); print({k: (v.get('value'), v.get('encode_gbs'), v.get('decode_gbs')) for k, v in d['per_config'].items()})"
python bench.py --impl reference --steps 1 --warmup 0 > $D/bench_ref.log 2>&1; tail -n 1 $D/bench_ref.log | head -c 200; echo
CMD="python bench.py --workload cfg2 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-per-config"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cfg2.csv $CMD > gpurun_out/ncu_launches_cfg2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lz4_encode_kernel -s 3 -c 1 -f -o gpurun_out/full_enc_cfg2 $CMD > gpurun_out/ncu_full_enc_cfg2.log 2>&1
tail -n 2 gpurun_out/ncu_full_enc_cfg2.log | head -c 300
