#!/bin/bash
# round 2, run x: traversal occupancy / cache-policy variants (tools/build_variants.sh) on config C3, one 64-spp pass each,
# and the GPU test of the thin Disney surfaces inside a medium
TAG=${1:-r02x}
OUT=gpurun_out
mkdir -p $OUT
: > $OUT/exp_${TAG}.jsonl
for v in "" _mb5 _mb6 _mb5s8 _b128mb8 _b128mb10 _stream _nodeel _streammb5; do
  LRK_DEVICE_LIB=libb200pt$v.so timeout 200 python tools/exp_trace.py >> $OUT/exp_${TAG}.jsonl 2>> $OUT/exp_${TAG}.err
done
cut -c1-230 $OUT/exp_${TAG}.jsonl; tail -3 $OUT/exp_${TAG}.err
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "thin" > $OUT/pytest_thin_${TAG}.log 2>&1; tail -3 $OUT/pytest_thin_${TAG}.log
