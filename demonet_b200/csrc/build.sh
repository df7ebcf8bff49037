#!/bin/bash
# Builds libdemonet_b200.so for sm_100a (nvcc cross-compiles without a GPU).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../lib"
mkdir -p "$OUT"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
SRCS="error.cu postprocess.cu dwconv.cu dwconv_tma.cu dwconv_stream.cu dwconv_stream2.cu se.cu transform.cu stem_tma.cu pwconv_simt.cu pwconv_tc.cu pwdw_fused.cu dwpw_fused.cu engine.cu"
OBJS=""
for s in $SRCS; do
  o="$OUT/${s%.cu}.o"
  if [ ! -f "$o" ] || [ "$HERE/$s" -nt "$o" ] || [ "$HERE/common.cuh" -nt "$o" ] || [ "$HERE/pwconv.cuh" -nt "$o" ] || [ "$HERE/dwconv.cuh" -nt "$o" ] || [ "$HERE/../../include/demonet_b200.h" -nt "$o" ]; then
    $NVCC $FLAGS -c "$HERE/$s" -o "$o" &
  fi
  OBJS="$OBJS $o"
done
wait
$NVCC -Wno-deprecated-gpu-targets -shared -o "$OUT/libdemonet_b200.so" $OBJS -lcudart_static -ldl -lrt -lpthread
echo "built $OUT/libdemonet_b200.so"
