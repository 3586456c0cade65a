#!/bin/bash
# Build exp/libev2h_<name>.so: sa_fused_tc.cu recompiled with the given -D flags, other objects reused.
# usage: tools/build_variant.sh name -DFOO -DBAR ;  run with EV2H_LIB=exp/libev2h_<name>.so
set -e
cd "$(dirname "$0")/.."
NAME=$1; shift
mkdir -p exp
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden "$@" \
    -c ev2hands_b200/csrc/sa_fused_tc.cu -o exp/sa_fused_tc_$NAME.o
OBJS=$(ls ev2hands_b200/build/*.o | grep -v sa_fused_tc.o)
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o exp/libev2h_$NAME.so exp/sa_fused_tc_$NAME.o $OBJS -lcudart
echo exp/libev2h_$NAME.so
