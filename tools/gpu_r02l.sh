#!/bin/bash
TAG=${1:-r02l}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_gpu_${TAG}.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu_${TAG}.log
grep -E "^FAILED|passed|failed|rc=" $OUT/pytest_gpu_${TAG}.log | cut -c1-200
: > $OUT/exp_${TAG}.jsonl
for v in "" _strict; do
  LRK_DEVICE_LIB=libb200pt$v.so timeout 300 python tools/exp_trace.py >> $OUT/exp_${TAG}.jsonl 2>> $OUT/exp_${TAG}.err
done
timeout 300 python tools/exp_trace.py --device-bvh --count >> $OUT/exp_${TAG}.jsonl 2>> $OUT/exp_${TAG}.err
timeout 300 python tools/exp_trace.py --scene C2 >> $OUT/exp_${TAG}.jsonl 2>> $OUT/exp_${TAG}.err
timeout 300 python tools/exp_trace.py --scene C4 --spp 16 >> $OUT/exp_${TAG}.jsonl 2>> $OUT/exp_${TAG}.err
LRK_DEVICE_LIB=libb200pt_strict.so timeout 300 python tools/exp_trace.py --scene C4 --spp 16 >> $OUT/exp_${TAG}.jsonl 2>> $OUT/exp_${TAG}.err
cat $OUT/exp_${TAG}.jsonl
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:shade_kernelILj2 -s 10 -c 1 -f -o $OUT/prof_shade2_${TAG} \
    python tools/exp_trace.py --repeat 1 > $OUT/ncu_${TAG}.log 2>&1
tail -2 $OUT/ncu_${TAG}.log
