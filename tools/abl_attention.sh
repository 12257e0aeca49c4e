#!/bin/bash
# run tools/bench_attention.py (impl 3) once per variant library built by tools/build_variant.sh
for so in ai-interior-image-classifier_b200/_lib/libiic_b200.so ai-interior-image-classifier_b200/_lib/libiic_abl_*.so "$@"; do
  [ -f "$so" ] || continue
  echo -n "$(basename $so): "
  IIC_LIB=$PWD/$so timeout 120 python tools/bench_attention.py 3 2>&1 | grep "T=197" | head -1
done
