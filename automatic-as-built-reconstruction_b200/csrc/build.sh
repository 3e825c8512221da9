#!/bin/bash
# Build libscn_b200.so in-tree for sm_100a (the .so is git-ignored but travels with gpurun).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -O3 ${SCN_NVCC_EXTRA}"
OUT=../libscn_b200.so
SRCS="common.cu metadata.cu conv.cu conv_tc.cu bn.cu io.cu graph.cu roi.cu rpn.cu nms.cu"
mkdir -p build
pids=()
for f in $SRCS; do
  o=build/${f%.cu}.o
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ -n "$(find . -maxdepth 1 -name '*.cuh' -newer "$o")" ] || [ ../../include/scn_b200.h -nt "$o" ]; then
    extra=""
    # nms.cu: no FMA contraction, so the float32 box geometry rounds like the golden values of the reference kernels
    [ "$f" = nms.cu ] && extra="-fmad=false"
    $NVCC $FLAGS $extra -c "$f" -o "$o" &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -shared -o $OUT build/*.o
echo "built $(realpath $OUT)"
