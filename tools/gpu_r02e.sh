#!/bin/bash
TAG=${1:-r02e}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu_${TAG}.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu_${TAG}.log
tail -8 $OUT/pytest_gpu_${TAG}.log
timeout 300 python tools/exp_trace.py --count > $OUT/exp_${TAG}.jsonl 2> $OUT/exp_${TAG}.err
for v in r01like stk16rec stk0 stk8 stk24; do
  LRK_DEVICE_LIB=libb200pt_${v}.so timeout 300 python tools/exp_trace.py >> $OUT/exp_${TAG}.jsonl 2>> $OUT/exp_${TAG}.err
done
timeout 600 python tools/exp_trace.py --repeat 1 refill_below=12,16,20 inner_min=6,8,10 >> $OUT/exp_${TAG}.jsonl 2>> $OUT/exp_${TAG}.err
cat $OUT/exp_${TAG}.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_closest_kernel -s 10 -c 2 -f -o $OUT/prof_closest_${TAG} \
    python tools/exp_trace.py --repeat 1 > $OUT/ncu_${TAG}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_shadow_kernel -s 10 -c 2 -f -o $OUT/prof_shadow_${TAG} \
    python tools/exp_trace.py --repeat 1 >> $OUT/ncu_${TAG}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:shade_kernel -s 6 -c 3 -f -o $OUT/prof_shade_${TAG} \
    python tools/exp_trace.py --repeat 1 >> $OUT/ncu_${TAG}.log 2>&1
