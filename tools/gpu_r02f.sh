#!/bin/bash
TAG=${1:-r02f}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_gpu_${TAG}.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu_${TAG}.log
tail -12 $OUT/pytest_gpu_${TAG}.log
timeout 300 python tools/exp_trace.py --count > $OUT/exp_${TAG}.jsonl 2> $OUT/exp_${TAG}.err
LRK_DEVICE_LIB=libb200pt_ieee.so timeout 300 python tools/exp_trace.py >> $OUT/exp_${TAG}.jsonl 2>> $OUT/exp_${TAG}.err
timeout 300 python tools/exp_trace.py --scene C2 --spp 64 >> $OUT/exp_${TAG}.jsonl 2>> $OUT/exp_${TAG}.err
cat $OUT/exp_${TAG}.jsonl
timeout 900 python bench.py --steps 8 --warmup 3 > $OUT/bench_${TAG}.json 2> $OUT/bench_${TAG}.err; echo "bench rc=$?"
cat $OUT/bench_${TAG}.json
tail -5 $OUT/bench_${TAG}.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > $OUT/bench_ref_${TAG}.json 2>> $OUT/bench_${TAG}.err; echo "ref rc=$?"
cat $OUT/bench_ref_${TAG}.json
