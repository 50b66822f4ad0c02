#!/bin/bash
# Strong-scaling sweep on one multi-GPU box: bench.py at N = 1, 2, 4, 8 (as many as the box has), same launch line as the driver's.
OUT=gpurun_out
mkdir -p $OUT
NG=$(nvidia-smi -L | wc -l)
for N in 1 2 4 8; do
  [ $N -le $NG ] || continue
  if [ $N -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu > $OUT/scale_${N}.json 2> $OUT/scale_${N}.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) \
        bench.py --gpus $N --steps 10 --warmup 3 --no-cpu > $OUT/scale_${N}.json 2> $OUT/scale_${N}.err
  fi
  tail -1 $OUT/scale_${N}.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print($N, d['value'], d['ms_per_step'], d.get('e2e',{}).get('value'))"
done
