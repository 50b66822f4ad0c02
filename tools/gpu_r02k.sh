#!/bin/bash
TAG=${1:-r02k}
OUT=gpurun_out
mkdir -p $OUT
: > $OUT/exp_${TAG}.jsonl
for v in "" _fma _fastdiv _fastmath; do
  LRK_DEVICE_LIB=libb200pt$v.so timeout 300 python tools/exp_trace.py >> $OUT/exp_${TAG}.jsonl 2>> $OUT/exp_${TAG}.err
done
cat $OUT/exp_${TAG}.jsonl
for v in _fastdiv _fastmath; do
  LRK_DEVICE_LIB=libb200pt$v.so timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu_${TAG}$v.log 2>&1
  tail -25 $OUT/pytest_gpu_${TAG}$v.log | grep -E "FAILED|passed|failed"
done
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:shade_kernelILj2 -s 10 -c 1 -f -o $OUT/prof_shade2_${TAG} \
    python tools/exp_trace.py --repeat 1 > $OUT/ncu_${TAG}.log 2>&1
tail -3 $OUT/ncu_${TAG}.log
