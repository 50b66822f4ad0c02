#!/bin/bash
# strong-scaling table on ONE 8-GPU box: bench.py at N = 8, 4, 2, 1 back to back (gpurun --gpus 8)
TAG=${1:-r02q}
OUT=gpurun_out
mkdir -p $OUT
for N in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2956$N bench.py --gpus $N --steps 8 --warmup 3 \
      > $OUT/scale_${N}gpu_${TAG}.json 2> $OUT/scale_${N}gpu_${TAG}.err; echo "N=$N rc=$?"
  python - <<PY
import json
d=json.loads(open("$OUT/scale_${N}gpu_${TAG}.json").read().strip().split("\n")[-1])
print($N, d["value"], d["e2e"]["value"], d["per_rank"], d["film_check"], d["configs"])
PY
done
timeout 600 python bench.py --gpus 1 --steps 8 --warmup 3 --no-cpu > $OUT/scale_1gpu_${TAG}.json 2> $OUT/scale_1gpu_${TAG}.err; echo "N=1 rc=$?"
python -c "
import json
d=json.loads(open('$OUT/scale_1gpu_${TAG}.json').read().strip().split('\n')[-1]); print(1, d['value'], d['e2e']['value'])"
timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -q 2>&1 | tail -2
