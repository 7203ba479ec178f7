#!/bin/bash
# Under gpurun on ONE GPU: plain run of the LZ4 stage on a 64-frame cfg2 slab (must exit 0), then one `ncu --set full`
# capture of lz4_encode_kernel with source counters; the same for 8-bit quantiser codes (noise: the early-store path)
# when the third argument is given. Outputs in gpurun_out/$1.
D=gpurun_out/${1:-enc}
SHAPE=${2:-64x2048x2048}
mkdir -p $D
for MODE in rmest ${3:-}; do
  CMD="python tools/profile_lz4.py $SHAPE 2 $MODE"
  $CMD > $D/plain_$MODE.log 2>&1 || { echo "plain run failed"; tail -5 $D/plain_$MODE.log; exit 1; }
  tail -1 $D/plain_$MODE.log
  ncu --set full --clock-control none --import-source on -k regex:lz4_encode_kernel -s 1 -c 1 -f -o $D/enc_$MODE $CMD > $D/ncu_enc_$MODE.log 2>&1
done
if [ -n "${4:-}" ]; then
  CMD="python tools/profile_lz4.py $SHAPE 2 rmest"
  ncu --set full --clock-control none --import-source on -k regex:lz4_decode_kernel -s 1 -c 1 -f -o $D/dec_rmest $CMD > $D/ncu_dec.log 2>&1
fi
ls -la $D
