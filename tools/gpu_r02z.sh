#!/bin/bash
# round 2, last run: the GPU suite and one default bench line on the final tree
TAG=${1:-r02z}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q > $OUT/pytest_gpu_${TAG}.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu_${TAG}.log
grep -E "^FAILED|^ERROR|passed|failed|rc=" $OUT/pytest_gpu_${TAG}.log | cut -c1-220
timeout 300 python bench.py --steps 8 --warmup 3 --no-configs > $OUT/bench_${TAG}.json 2> $OUT/bench_${TAG}.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('$OUT/bench_${TAG}.json').read().strip().split('\n')[-1]); print(d['value'], d['e2e']['value'], d['roofline']['traffic'], d['roofline'].get('dram_frac'))"
