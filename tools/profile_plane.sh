#!/bin/bash
# Under gpurun on ONE GPU: ncu --set full of lz4_encode_kernel (and lz4_decode_kernel) on ONE bit plane of a cfg2 slab
# usage: tools/profile_plane.sh outdir plane [shape]
D=gpurun_out/$1; K=$2; SHAPE=${3:-256x2048x2048}
mkdir -p $D
CMD="python tools/profile_lz4.py $SHAPE 2 rmest $K"
$CMD > $D/plain_p$K.log 2>&1 || { echo "plain run failed"; tail -5 $D/plain_p$K.log; exit 1; }
tail -1 $D/plain_p$K.log
ncu --set full --clock-control none --import-source on -k regex:lz4_encode_kernel -s 1 -c 1 -f -o $D/enc_p$K $CMD > $D/ncu_enc_p$K.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lz4_decode_kernel -s 1 -c 1 -f -o $D/dec_p$K $CMD > $D/ncu_dec_p$K.log 2>&1
ls -la $D
