#!/bin/bash
# Kernel experiments: builds luisarender_b200/lib/libb200pt_<tag>.so with extra nvcc flags; select one at run time with
# LRK_DEVICE_LIB=libb200pt_<tag>.so (luisarender_b200/_ffi.py).   usage: tools/build_variants.sh tag1="-DX" tag2="-DY -DZ" ...
set -e
cd "$(dirname "$0")/.."
for spec in "$@"; do
  tag="${spec%%=*}"; flags="${spec#*=}"
  ( nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=${MAD:-false} --shared -Xcompiler -fPIC $flags \
      luisarender_b200/csrc/device/lrk.cu -o luisarender_b200/lib/libb200pt_${tag}.so && echo "built $tag ($flags)" ) &
done
wait
