#!/bin/bash
# Kernel experiments: builds luisarender_b200/lib/libb200pt_<tag>.so with extra nvcc flags (both translation units); select one at
# run time with LRK_DEVICE_LIB=libb200pt_<tag>.so (luisarender_b200/_ffi.py).  LRK_SHADE_STRICT=1 compiles shade.cu with IEEE
# arithmetic like lrk.cu.   usage: [LRK_SHADE_STRICT=1] tools/build_variants.sh tag1="-DX" tag2="-DY -DZ" ...
set -e
cd "$(dirname "$0")/.."
for spec in "$@"; do
  tag="${spec%%=*}"; flags="${spec#*=}"
  ( python -c "
import sys
from luisarender_b200 import build
build.build_device(force=True, name='libb200pt_${tag}.so', extra_flags=tuple('${flags}'.split()))
print('built ${tag} (${flags})')" ) &
done
wait
