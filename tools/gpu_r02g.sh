#!/bin/bash
TAG=${1:-r02g}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_gpu_${TAG}.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu_${TAG}.log
tail -12 $OUT/pytest_gpu_${TAG}.log
timeout 300 python tools/exp_trace.py --count > $OUT/exp_${TAG}.jsonl 2> $OUT/exp_${TAG}.err
timeout 300 python tools/exp_trace.py --scene C2 --spp 64 >> $OUT/exp_${TAG}.jsonl 2>> $OUT/exp_${TAG}.err
cat $OUT/exp_${TAG}.jsonl
timeout 900 python bench.py --steps 8 --warmup 3 > $OUT/bench_${TAG}.json 2> $OUT/bench_${TAG}.err; echo "bench rc=$?"
cat $OUT/bench_${TAG}.json
tail -5 $OUT/bench_${TAG}.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > $OUT/bench_ref_${TAG}.json 2>> $OUT/bench_${TAG}.err; echo "ref rc=$?"
cat $OUT/bench_ref_${TAG}.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 1 --no-configs > $OUT/bench_under_ncu_${TAG}.log 2>&1
for k in trace_closest_kernel trace_shadow_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 10 -c 1 -f -o $OUT/prof_${k}_${TAG} \
      python tools/exp_trace.py --repeat 1 >> $OUT/ncu_${TAG}.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:shade_kernel -s 6 -c 3 -f -o $OUT/prof_shade_${TAG} \
    python tools/exp_trace.py --repeat 1 >> $OUT/ncu_${TAG}.log 2>&1
ls -la $OUT | tail -12
