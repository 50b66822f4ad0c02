#!/bin/bash
TAG=${1:-r02u}
OUT=gpurun_out
mkdir -p $OUT
timeout 1700 python -m pytest tests -m gpu -q > $OUT/pytest_gpu_${TAG}.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu_${TAG}.log
grep -E "^FAILED|passed|failed|rc=" $OUT/pytest_gpu_${TAG}.log | cut -c1-220
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_${TAG}.log 2>&1; tail -2 $OUT/smoke_${TAG}.log
timeout 900 python bench.py --steps 8 --warmup 3 > $OUT/bench_${TAG}.json 2> $OUT/bench_${TAG}.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('$OUT/bench_${TAG}.json').read().strip().split('\n')[-1]); print(d['value'], d['e2e']['value'], d['configs'] and {k:v['msamples_per_s'] for k,v in d['configs'].items()}, d['arithmetic']['strict_math']['value'])"
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > $OUT/bench_ref_${TAG}.json 2>> $OUT/bench_${TAG}.err; echo "ref rc=$?"
python -c "
import json
d=json.loads(open('$OUT/bench_ref_${TAG}.json').read().strip().split('\n')[-1]); print(d['value'], d['cpu_baseline']['cores'])"
timeout 300 python tools/exp_trace.py --scene F3 > $OUT/exp_${TAG}.jsonl 2> $OUT/exp_${TAG}.err; cat $OUT/exp_${TAG}.jsonl | cut -c1-300
