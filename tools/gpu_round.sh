#!/bin/bash
# One GPU session: parity tests, bench (both arms), ncu launch list, DRAM traffic of the traversal launches of one pass, and
# full captures of the first launches of the hot kernels.  Usage (under gpurun): bash tools/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu_${TAG}.txt
python __graft_entry__.py > $OUT/build_${TAG}.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_${TAG}.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu_${TAG}.log
tail -3 $OUT/pytest_gpu_${TAG}.log
timeout 600 python bench.py --steps 8 --warmup 3 > $OUT/bench_${TAG}.json 2> $OUT/bench_${TAG}.err; echo "bench rc=$?"
cat $OUT/bench_${TAG}.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_ref_${TAG}.json 2>> $OUT/bench_${TAG}.err
cat $OUT/bench_ref_${TAG}.json
if [ -z "$NO_NCU" ]; then
# (1) launch list: per-launch device time of every kernel of the first steps
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_${TAG}.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu > $OUT/ncu_bench_${TAG}.log 2>&1
# (2) DRAM bytes + time of the closest-hit traversal launches of one pass (10 bounces): the `traffic` figure of the roofline
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:trace_closest_kernel -c 10 --csv --log-file $OUT/traffic_${TAG}.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu >> $OUT/ncu_bench_${TAG}.log 2>&1
# (3) full captures, source-correlated (skipped with NO_FULL=1)
if [ -z "$NO_FULL" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:trace_closest_kernel -s 0 -c 2 -f -o $OUT/prof_trace_${TAG} \
    python bench.py --steps 1 --warmup 1 --no-cpu > $OUT/ncu_full_${TAG}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:shade_kernel -s 0 -c 3 -f -o $OUT/prof_shade_${TAG} \
    python bench.py --steps 1 --warmup 1 --no-cpu >> $OUT/ncu_full_${TAG}.log 2>&1
fi
ls -la $OUT | grep ${TAG}
fi
