#!/bin/bash
TAG=${1:-r02h}
OUT=gpurun_out
mkdir -p $OUT
: > $OUT/exp_${TAG}.jsonl
for v in "" _fma _sb128x5 _sb256x3 _sb128x4 _fma_sb128x5; do
  LRK_DEVICE_LIB=libb200pt$v.so timeout 300 python tools/exp_trace.py >> $OUT/exp_${TAG}.jsonl 2>> $OUT/exp_${TAG}.err
done
cat $OUT/exp_${TAG}.jsonl
# every traversal launch of one 64-spp pass: DRAM bytes, lanes per instruction, issue utilisation
timeout 900 ncu --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct \
    -k regex:"trace_(closest|shadow)_kernel" -s 4 -c 40 --csv --log-file $OUT/traversal_metrics_${TAG}.csv python tools/exp_trace.py --repeat 1 > $OUT/ncu_${TAG}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"shade_kernel<2" -s 3 -c 2 -f -o $OUT/prof_shade2_${TAG} \
    python tools/exp_trace.py --repeat 1 >> $OUT/ncu_${TAG}.log 2>&1
ls -la $OUT | tail -6
