#!/bin/bash
# Times bench.py with each prebuilt lib/variant_*.so in place of libb200pt.so (occupancy experiments).
cd luisarender_b200/lib
cp libb200pt.so /tmp/orig.so
for v in variant_*.so; do
  cp $v libb200pt.so
  (cd ../.. && python bench.py --steps 6 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', d['value'], d['ms_per_step'], d['roofline']['kernel_ms_total'], d['roofline']['other_kernels_ms'])")
done
cp /tmp/orig.so libb200pt.so
