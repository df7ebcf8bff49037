#!/bin/bash
# Builds the C-ABI library for sm_100a (nvcc cross-compiles without a GPU), once per activation storage type:
#   lib/libdemonet_b200_fp16.so  (-DDN_ACT_FP16=1, the default the Python side loads)
#   lib/libdemonet_b200_bf16.so  (-DDN_ACT_FP16=0)
# DN_BUILD_DTYPES="fp16" builds only one of them (faster edit-compile loop).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="${DN_LIB_OUT:-$HERE/../lib}"      # DN_LIB_OUT / DN_EXTRA_FLAGS: variant builds for A/B measurements (DN_LIB_DIR selects one at run time)
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC ${DN_EXTRA_FLAGS:-}"
SRCS="error.cu postprocess.cu dwconv.cu dwconv_tma.cu dwconv_stream.cu dwconv_stream2.cu se.cu transform.cu stem_tma.cu stem_tc.cu pwconv_simt.cu pwconv_tc.cu pwdw_fused.cu dwpw_fused.cu conv3x3_tc.cu vgg_ops.cu ssd_loss.cu engine.cu"
DTYPES=${DN_BUILD_DTYPES:-"fp16 bf16"}
for dt in $DTYPES; do
  if [ "$dt" = "fp16" ]; then DEF="-DDN_ACT_FP16=1"; else DEF="-DDN_ACT_FP16=0"; fi
  mkdir -p "$OUT/$dt"
  for s in $SRCS; do
    o="$OUT/$dt/${s%.cu}.o"
    stale=0
    [ -f "$o" ] || stale=1
    for dep in "$HERE/$s" "$HERE"/*.cuh "$HERE/../../include/demonet_b200.h" "$HERE/build.sh"; do
      [ "$dep" -nt "$o" ] && stale=1
    done
    if [ $stale = 1 ]; then
      $NVCC $FLAGS $DEF -c "$HERE/$s" -o "$o" &
    fi
  done
done
wait
for dt in $DTYPES; do
  OBJS=""
  for s in $SRCS; do OBJS="$OBJS $OUT/$dt/${s%.cu}.o"; done
  $NVCC -Wno-deprecated-gpu-targets -shared -o "$OUT/libdemonet_b200_$dt.so" $OBJS -lcudart_static -ldl -lrt -lpthread
  echo "built $OUT/libdemonet_b200_$dt.so"
done
