#!/bin/bash
# quick GPU check of an LZ4 encoder change: LZ4 tests, then A/B of the libraries given as arguments (tools/ab_lz4.py)
D=gpurun_out/${1:-q}; shift
mkdir -p $D
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_determinism.py tests/test_gpu_lz4_decoders.py tests/test_gpu_staging.py -m gpu -q -x > $D/pytest.log 2>&1; tail -n 6 $D/pytest.log
timeout 600 python tools/ab_lz4.py "$@" > $D/ab.log 2>&1; cat $D/ab.log
