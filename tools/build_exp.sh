#!/bin/bash
# Developer tool: build an experiment variant of libscn_b200.so into tools/_exp/ (git-ignored; travels with gpurun).
#   tools/build_exp.sh STALLS        -> -DSCN_EXPERIMENT_STALLS   (tools/gemm_stalls.py)
#   tools/build_exp.sh NO_A | NO_B | NO_MMA | NO_EPI | XCOMMIT   -> the what-if builds of profiles/experiments/README.md
# Only conv_tc.cu is recompiled; the other objects come from csrc/build (run csrc/build.sh first).
set -e
V=${1:?variant}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
C=$ROOT/automatic-as-built-reconstruction_b200/csrc
mkdir -p "$ROOT/tools/_exp"
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -O3 \
  -DSCN_EXPERIMENT_$V -c "$C/conv_tc.cu" -o "/tmp/conv_tc_$V.o"
/usr/local/cuda/bin/nvcc -shared -o "$ROOT/tools/_exp/libscn_$V.so" "$C/build/common.o" "$C/build/metadata.o" "$C/build/conv.o" \
  "/tmp/conv_tc_$V.o" "$C/build/bn.o" "$C/build/io.o" "$C/build/graph.o" "$C/build/roi.o" "$C/build/rpn.o" "$C/build/nms.o"
echo "built $ROOT/tools/_exp/libscn_$V.so"
