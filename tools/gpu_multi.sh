#!/bin/bash
# Multi-GPU pass (under gpurun --gpus N): the sharding tests must RUN (TTIRT_EXPECT_GPUS), then the bench line at N ranks.
#   usage: bash tools/gpu_multi.sh <tag> <N> [steps]
set -u
TAG=${1:-r02m}; N=${2:-2}; STEPS=${3:-3}
OUT=gpurun_out
mkdir -p $OUT
{ nvidia-smi -L; nproc; free -g; nvidia-smi topo -m; } > $OUT/${TAG}_env.log 2>&1
TTIRT_EXPECT_GPUS=$N timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "multi_device or virtual_devices" -rs > $OUT/${TAG}_pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -4 $OUT/${TAG}_pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps $STEPS --warmup 3 > $OUT/${TAG}_bench_n$N.json 2> $OUT/${TAG}_bench_n$N.err; echo "bench N=$N rc=$?"
tail -c 6000 $OUT/${TAG}_bench_n$N.json; tail -5 $OUT/${TAG}_bench_n$N.err
