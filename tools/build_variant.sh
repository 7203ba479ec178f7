#!/bin/bash
# development tool: builds build/ab/lib_<name>.so = the library with lz4_encode.cu (or $SRC) compiled with extra -D flags
# usage: tools/build_variant.sh name [-DX ...]
set -e
NAME=$1; shift
SRC=${SRC:-device/lz4_encode}
mkdir -p build/ab build/abobj
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-fvisibility=default,-ffp-contract=off -Xptxas -v \
  "$@" -c sqeazy_b200/csrc/$SRC.cu -o build/abobj/$NAME.o 2> build/abobj/$NAME.ptxas.log
OBJS=$(ls build/obj/*.o build/obj/device/*.o build/obj/host/*.o | grep -v "$SRC.o")
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/ab/lib_$NAME.so build/abobj/$NAME.o $OBJS -cudart static -ldl
grep -A1 "lz4_encode_kernel\|encode_general" build/abobj/$NAME.ptxas.log | grep -i "registers\|spill" | head -4
