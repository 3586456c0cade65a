#!/bin/bash
# Build exp/libev2h_<name>.so: ONE source of csrc/ recompiled with the given -D flags, the other objects reused.
# usage: tools/build_variant.sh name source.cu [-DFOO ...] ;  run with EV2H_LIB=exp/libev2h_<name>.so
set -e
cd "$(dirname "$0")/.."
NAME=$1; SRC=$2; shift; shift
BASE=$(basename $SRC .cu)
mkdir -p exp
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden "$@" \
    -c ev2hands_b200/csrc/$BASE.cu -o exp/${BASE}_$NAME.o
OBJS=$(ls ev2hands_b200/build/*.o | grep -v "/$BASE.o")
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o exp/libev2h_$NAME.so exp/${BASE}_$NAME.o $OBJS -lcudart
echo exp/libev2h_$NAME.so
