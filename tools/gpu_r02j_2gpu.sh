#!/bin/bash
TAG=${1:-r02j}
N=${2:-2}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_gpu_parity.py::test_volume_with_shape_media_matches_oracle -m gpu -q > $OUT/pytest_multigpu_${TAG}.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_multigpu_${TAG}.log
tail -15 $OUT/pytest_multigpu_${TAG}.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 8 --warmup 3 \
    > $OUT/bench_${N}gpu_${TAG}.json 2> $OUT/bench_${N}gpu_${TAG}.err; echo "bench rc=$?"
cat $OUT/bench_${N}gpu_${TAG}.json
tail -5 $OUT/bench_${N}gpu_${TAG}.err
