#!/bin/bash
TAG=${1:-r02i}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_gpu_${TAG}.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu_${TAG}.log
tail -30 $OUT/pytest_gpu_${TAG}.log
: > $OUT/exp_${TAG}.jsonl
for v in "" _sb128x4 _sb128x5 _sb128x6 _sb256x3 _sb512x1 _sb64x8; do
  LRK_DEVICE_LIB=libb200pt$v.so timeout 300 python tools/exp_trace.py >> $OUT/exp_${TAG}.jsonl 2>> $OUT/exp_${TAG}.err
done
cat $OUT/exp_${TAG}.jsonl
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:shade_kernelILj2 -s 3 -c 2 -f -o $OUT/prof_shade2_${TAG} \
    python tools/exp_trace.py --repeat 1 > $OUT/ncu_${TAG}.log 2>&1
tail -3 $OUT/ncu_${TAG}.log
ls -la $OUT | grep ${TAG}
