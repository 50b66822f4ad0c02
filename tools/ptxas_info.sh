#!/bin/bash
# registers / spills / shared memory of the kernels matching $1 (regex on the demangled name), from a cubin-only compile
set -e
cd "$(dirname "$0")/.."
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -cubin -Xptxas -v $EXTRA_NVCC_FLAGS \
    luisarender_b200/csrc/device/lrk.cu -o /tmp/lrk.cubin 2>&1 | c++filt | grep -A2 -E "Compiling entry function.*(${1:-trace})" | grep -E "Compiling|registers|spill" | sed -e 's/ptxas info    : //'
