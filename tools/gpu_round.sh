#!/bin/bash
# One GPU session: parity tests, bench (both arms), ncu launch list + one full capture of the traversal kernel.
# Usage (under gpurun): bash tools/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu_${TAG}.txt
python __graft_entry__.py > $OUT/build_${TAG}.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_${TAG}.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu_${TAG}.log
tail -5 $OUT/pytest_gpu_${TAG}.log
timeout 600 python bench.py --steps 20 --warmup 3 > $OUT/bench_${TAG}.json 2> $OUT/bench_${TAG}.err; echo "bench rc=$?"
cat $OUT/bench_${TAG}.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_ref_${TAG}.json 2>> $OUT/bench_${TAG}.err
cat $OUT/bench_ref_${TAG}.json
if [ -z "$NO_NCU" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > $OUT/ncu_bench_${TAG}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:trace_closest_kernel -s 4 -c 3 -f -o $OUT/prof_trace_${TAG} \
    python bench.py --steps 1 --warmup 1 --no-cpu > $OUT/ncu_full_${TAG}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:shade_kernel -s 4 -c 2 -f -o $OUT/prof_shade_${TAG} \
    python bench.py --steps 1 --warmup 1 --no-cpu >> $OUT/ncu_full_${TAG}.log 2>&1
ls -la $OUT | tail -20
fi
