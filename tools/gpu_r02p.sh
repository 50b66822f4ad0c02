#!/bin/bash
# final single-GPU evidence of round 2: tests in both arithmetic modes, smoke, both bench arms, per-kernel breakdown, launch list,
# per-launch traversal metrics, full ncu reports of the three hot kernels
TAG=${1:-r02p}
OUT=gpurun_out
mkdir -p $OUT
timeout 1700 python -m pytest tests -m gpu -q > $OUT/pytest_gpu_${TAG}.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu_${TAG}.log
grep -E "^FAILED|passed|failed|rc=" $OUT/pytest_gpu_${TAG}.log | cut -c1-220
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_${TAG}.log 2>&1; tail -2 $OUT/smoke_${TAG}.log
: > $OUT/exp_${TAG}.jsonl
timeout 300 python tools/exp_trace.py --count >> $OUT/exp_${TAG}.jsonl 2>> $OUT/exp_${TAG}.err
for v in _sb256x3 _sb128x5; do
  LRK_DEVICE_LIB=libb200pt$v.so timeout 300 python tools/exp_trace.py >> $OUT/exp_${TAG}.jsonl 2>> $OUT/exp_${TAG}.err
done
timeout 300 python tools/exp_trace.py --scene C2 >> $OUT/exp_${TAG}.jsonl 2>> $OUT/exp_${TAG}.err
timeout 300 python tools/exp_trace.py --scene C4 --spp 16 >> $OUT/exp_${TAG}.jsonl 2>> $OUT/exp_${TAG}.err
timeout 300 python tools/exp_trace.py --scene F3 >> $OUT/exp_${TAG}.jsonl 2>> $OUT/exp_${TAG}.err
cat $OUT/exp_${TAG}.jsonl
timeout 900 python bench.py --steps 8 --warmup 3 > $OUT/bench_${TAG}.json 2> $OUT/bench_${TAG}.err; echo "bench rc=$?"
cat $OUT/bench_${TAG}.json; tail -3 $OUT/bench_${TAG}.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > $OUT/bench_ref_${TAG}.json 2>> $OUT/bench_${TAG}.err; echo "ref rc=$?"
cat $OUT/bench_ref_${TAG}.json | cut -c1-600
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 1 --no-configs --no-cpu > $OUT/bench_under_ncu_${TAG}.log 2>&1
timeout 900 ncu --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct \
    -k regex:"trace_(closest|shadow)_kernel" -s 4 -c 40 --csv --log-file $OUT/traversal_metrics_${TAG}.csv python tools/exp_trace.py --repeat 1 > $OUT/ncu_${TAG}.log 2>&1
for k in trace_closest_kernel trace_shadow_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 10 -c 1 -f -o $OUT/prof_${k}_${TAG} \
      python tools/exp_trace.py --repeat 1 >> $OUT/ncu_${TAG}.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:shade_kernelILj2 -s 10 -c 1 -f -o $OUT/prof_shade2_${TAG} \
    python tools/exp_trace.py --repeat 1 >> $OUT/ncu_${TAG}.log 2>&1
ls -la $OUT | grep ${TAG}
