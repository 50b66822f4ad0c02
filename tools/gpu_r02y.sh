#!/bin/bash
# round 2, final single-GPU evidence with the last code (thin Disney, image-emission lights, streamed ray records, 5-block any-hit
# kernel): tests in both arithmetic modes, smoke, per-kernel breakdown, bench, launch list, per-launch traversal metrics
TAG=${1:-r02y}
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest_gpu_${TAG}.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu_${TAG}.log
grep -E "^FAILED|^ERROR|passed|failed|rc=" $OUT/pytest_gpu_${TAG}.log | cut -c1-220
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_${TAG}.log 2>&1; tail -2 $OUT/smoke_${TAG}.log
: > $OUT/exp_${TAG}.jsonl
timeout 200 python tools/exp_trace.py --count >> $OUT/exp_${TAG}.jsonl 2>> $OUT/exp_${TAG}.err
cut -c1-400 $OUT/exp_${TAG}.jsonl
timeout 600 python bench.py --steps 8 --warmup 3 > $OUT/bench_${TAG}.json 2> $OUT/bench_${TAG}.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('$OUT/bench_${TAG}.json').read().strip().split('\n')[-1]); print(d['value'], d['e2e']['value'], d['configs'] and {k:v['msamples_per_s'] for k,v in d['configs'].items()}, d['arithmetic']['strict_math']['value'], d['roofline'])"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 1 --no-configs --no-cpu > $OUT/bench_under_ncu_${TAG}.log 2>&1
timeout 600 ncu --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct \
    -k regex:"trace_(closest|shadow)_kernel" -s 4 -c 40 --csv --log-file $OUT/traversal_metrics_${TAG}.csv python tools/exp_trace.py --repeat 1 > $OUT/ncu_${TAG}.log 2>&1
ls -la $OUT | grep ${TAG}
