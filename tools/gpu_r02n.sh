#!/bin/bash
TAG=${1:-r02n}
OUT=gpurun_out
mkdir -p $OUT
: > $OUT/parity_${TAG}.jsonl; : > $OUT/exp_${TAG}.jsonl
for v in "" _strict; do
  LRK_DEVICE_LIB=libb200pt$v.so timeout 300 python tools/exp_parity.py >> $OUT/parity_${TAG}.jsonl 2>> $OUT/exp_${TAG}.err
  LRK_DEVICE_LIB=libb200pt$v.so timeout 300 python tools/exp_trace.py >> $OUT/exp_${TAG}.jsonl 2>> $OUT/exp_${TAG}.err
done
cat $OUT/parity_${TAG}.jsonl; cat $OUT/exp_${TAG}.jsonl; tail -3 $OUT/exp_${TAG}.err
timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_gpu_${TAG}.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu_${TAG}.log
grep -E "^FAILED|passed|failed|rc=" $OUT/pytest_gpu_${TAG}.log | cut -c1-200
