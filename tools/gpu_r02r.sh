#!/bin/bash
TAG=${1:-r02r}
OUT=gpurun_out
mkdir -p $OUT
timeout 1700 python -m pytest tests -m gpu -q > $OUT/pytest_gpu_${TAG}.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu_${TAG}.log
grep -E "^FAILED|passed|failed|rc=" $OUT/pytest_gpu_${TAG}.log | cut -c1-220
timeout 300 python tools/exp_trace.py > $OUT/exp_${TAG}.jsonl 2> $OUT/exp_${TAG}.err; cat $OUT/exp_${TAG}.jsonl | cut -c1-400
