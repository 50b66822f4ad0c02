#!/bin/bash
# ncu full captures of the first (largest) launches of the three hot kernels.
TAG=${1:-p}
OUT=gpurun_out
mkdir -p $OUT
python __graft_entry__.py > $OUT/build_${TAG}.log 2>&1
for K in trace_closest_kernel trace_shadow_kernel shade_kernel; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 0 -c 2 -f -o $OUT/prof_${K}_${TAG} \
      python bench.py --steps 1 --warmup 1 --no-cpu > $OUT/ncu_${K}_${TAG}.log 2>&1
done
ls -la $OUT | grep ${TAG}
