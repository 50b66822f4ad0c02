#!/bin/bash
TAG=${1:-r02m}
OUT=gpurun_out
mkdir -p $OUT
: > $OUT/parity_${TAG}.jsonl; : > $OUT/exp_${TAG}.jsonl
for v in _strict _sfma _sfmadiv _sfmadivsqrt _sftz ""; do
  LRK_DEVICE_LIB=libb200pt$v.so timeout 300 python tools/exp_parity.py >> $OUT/parity_${TAG}.jsonl 2>> $OUT/exp_${TAG}.err
  LRK_DEVICE_LIB=libb200pt$v.so timeout 300 python tools/exp_trace.py >> $OUT/exp_${TAG}.jsonl 2>> $OUT/exp_${TAG}.err
done
for v in _iu2 _iu3; do
  LRK_DEVICE_LIB=libb200pt$v.so timeout 300 python tools/exp_trace.py >> $OUT/exp_${TAG}.jsonl 2>> $OUT/exp_${TAG}.err
done
cat $OUT/parity_${TAG}.jsonl; cat $OUT/exp_${TAG}.jsonl; tail -3 $OUT/exp_${TAG}.err
