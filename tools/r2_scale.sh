#!/bin/bash
# scaling run on N GPUs of one box, launched the way the driver does: bench.py under torchrun, then the sharded tests
N=${1:-8}
D=gpurun_out/${2:-scale$N}
mkdir -p $D
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 ) > $D/bench_n$N.log 2> $D/bench_n$N.err
tail -n 1 $D/bench_n$N.log | head -c 5000; echo; tail -n 5 $D/bench_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --impl reference --gpus $N --steps 1 --warmup 0 > $D/ref_n$N.log 2>&1; tail -n 1 $D/ref_n$N.log | head -c 300; echo
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -q > $D/pytest_sharded.log 2>&1; tail -n 3 $D/pytest_sharded.log
