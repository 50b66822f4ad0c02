#!/bin/bash
TAG=${1:-r02s}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_ref_render.py -m gpu -q -k "multi or two_rank or cli_on or layered" > $OUT/pytest_${TAG}.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_${TAG}.log
grep -E "^FAILED|passed|failed|rc=" $OUT/pytest_${TAG}.log | cut -c1-220; grep -n "^E  " $OUT/pytest_${TAG}.log | head -8 | cut -c1-250
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 8 --warmup 3 \
    > $OUT/bench_2gpu_${TAG}.json 2> $OUT/bench_2gpu_${TAG}.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("$OUT/bench_2gpu_${TAG}.json").read().strip().split("\n")[-1])
print(d["value"], d["e2e"]["value"], d["per_rank"], d["film_check"], d["configs"])
PY
tail -3 $OUT/bench_2gpu_${TAG}.err
