#!/bin/bash
# r02 first GPU pass: sanitizer on a small case, parity tests, smoke, bench line, launch list + full ncu capture.
set -u
TAG=${1:-r02a}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L > $OUT/${TAG}_env.log; nproc >> $OUT/${TAG}_env.log; free -g >> $OUT/${TAG}_env.log
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python tests/devtools/sanity_small.py > $OUT/${TAG}_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 $OUT/${TAG}_memcheck.log
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/${TAG}_smoke.log
timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; cat $OUT/${TAG}_bench.json | cut -c1-3000
TTIRT_DEVICE_STREAMS=1 timeout 300 python bench.py --no-e2e --no-cpu --no-next-rows --no-other-configs > $OUT/${TAG}_bench_1stream.json 2>> $OUT/${TAG}_bench.err; echo "bench 1stream rc=$?"; cut -c1-400 $OUT/${TAG}_bench_1stream.json
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > $OUT/${TAG}_bench_reference.json 2>> $OUT/${TAG}_bench.err; echo "ref rc=$?"; cut -c1-600 $OUT/${TAG}_bench_reference.json
python bench.py --log2m 20 --steps 2 --warmup 1 --no-e2e --no-cpu --no-next-rows --no-other-configs > $OUT/${TAG}_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --log2m 20 --steps 2 --warmup 1 --no-e2e --no-cpu --no-next-rows --no-other-configs > $OUT/${TAG}_ncu_list.log 2>&1
echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:transition_kernel -s 40 -c 1 -f -o $OUT/${TAG}_transition \
    python bench.py --log2m 20 --steps 1 --warmup 1 --no-e2e --no-cpu --no-next-rows --no-other-configs > $OUT/${TAG}_ncu_full.log 2>&1
echo "ncu full rc=$?"
