#!/bin/bash
# tools/build_variant.sh <name> <source.cu> <extra nvcc flags...>: builds _lib/libiic_<name>.so = the current objects with ONE
# translation unit recompiled with extra flags (A/B experiments inside a single GPU call via IIC_LIB=...)
set -e
NAME=$1; SRC=$2; shift 2
D=ai-interior-image-classifier_b200
OBJS=""
for o in $D/_lib/*.o; do
  b=$(basename $o .o)
  if [ "$b.cu" == "$SRC" ]; then continue; fi
  case $b in var_*) continue;; esac
  OBJS="$OBJS $o"
done
nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -Xcompiler -fPIC,-ffp-contract=off "$@" -c $D/csrc/$SRC -o $D/_lib/var_$NAME.o
nvcc -shared -gencode arch=compute_100a,code=sm_100a -cudart static -o $D/_lib/libiic_$NAME.so $OBJS $D/_lib/var_$NAME.o
rm -f $D/_lib/var_$NAME.o
echo built $D/_lib/libiic_$NAME.so
