#!/bin/bash
# One GPU-box pass for a round: parity tests, the bench line, the ncu launch list and one full capture of the
# dominant kernel.  usage (from the repo root, under gpurun):  bash tools/gpu_round.sh <tag> [tests|notests]
# Everything lands in gpurun_out/<tag>_*; numbers printed by the runs under ncu are never bench values.
set -u
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
if [ "${2:-tests}" = "tests" ]; then
  python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/${TAG}_pytest_gpu.log
  python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/${TAG}_smoke.log
fi
python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; cat $OUT/${TAG}_bench.json
python bench.py --impl reference --steps 1 --warmup 1 > $OUT/${TAG}_bench_reference.json 2>> $OUT/${TAG}_bench.err; echo "ref rc=$?"; cat $OUT/${TAG}_bench_reference.json
# launch list of the same command at a short setting (cold-cache, serialised: shares only)
python bench.py --log2m 20 --steps 2 --warmup 1 --no-e2e --no-cpu --no-next-rows > $OUT/${TAG}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --log2m 20 --steps 2 --warmup 1 --no-e2e --no-cpu --no-next-rows > $OUT/${TAG}_ncu_list.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:transition_kernel -s 40 -c 1 -f -o $OUT/${TAG}_transition \
    python bench.py --log2m 20 --steps 1 --warmup 1 --no-e2e --no-cpu --no-next-rows > $OUT/${TAG}_ncu_full.log 2>&1
echo "ncu full rc=$?"
