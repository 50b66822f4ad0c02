#!/bin/bash
# registers / spills / shared memory of the kernels matching $1 (regex on the demangled name), from cubin-only compiles of the two
# translation units with the flags luisarender_b200/build.py uses (lrk.cu: IEEE arithmetic; shade.cu twice: fast math and strict)
set -e
cd "$(dirname "$0")/.."
for tu in "lrk.cu -fmad=false" "shade.cu --use_fast_math,-DLRK_SHADE_VARIANT=fast" "shade.cu -fmad=false,-DLRK_SHADE_VARIANT=strict"; do
  set -- "${1:-trace|shade}" $tu
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo ${3//,/ } -cubin -Xptxas -v $EXTRA_NVCC_FLAGS \
      luisarender_b200/csrc/device/$2 -o /tmp/${2%.cu}.cubin 2>&1 | c++filt | grep -A3 -E "Compiling entry function.*($1)" | grep -E "Compiling|registers|spill" | sed -e 's/ptxas info    : //'
done
