#!/bin/bash
# round 2, run v: the GPU suite with the thin Disney bucket + one default bench line (no reference arm, no experiments)
TAG=${1:-r02v}
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest_gpu_${TAG}.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu_${TAG}.log
grep -E "^FAILED|^ERROR|passed|failed|rc=" $OUT/pytest_gpu_${TAG}.log | cut -c1-220
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_${TAG}.log 2>&1; tail -2 $OUT/smoke_${TAG}.log
timeout 600 python bench.py --steps 8 --warmup 3 > $OUT/bench_${TAG}.json 2> $OUT/bench_${TAG}.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('$OUT/bench_${TAG}.json').read().strip().split('\n')[-1]); print(d['value'], d['e2e']['value'], d['configs'] and {k:v['msamples_per_s'] for k,v in d['configs'].items()}, d['arithmetic']['strict_math']['value'])"
