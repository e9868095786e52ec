#!/bin/bash
set -u
TAG=${1:-r02e}; OUT=gpurun_out; mkdir -p $OUT
timeout 300 python tools/e2e_trace.py > $OUT/${TAG}_e2e_plain.log 2>&1; cat $OUT/${TAG}_e2e_plain.log
TTIRT_TRACE=1 timeout 300 python tools/e2e_trace.py > $OUT/${TAG}_e2e_trace.log 2>&1; grep -v "^$" $OUT/${TAG}_e2e_trace.log | tail -40
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "allocation or virtual or edge" > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/${TAG}_pytest.log
for ch in 1048576 2097152 4194304; do TTIRT_CHUNK=$ch timeout 200 python bench.py --shape 40,33,32 --log2m 22 --steps 5 --warmup 3 --no-cpu --no-e2e --no-next-rows > $OUT/${TAG}_c3_chunk$ch.json 2>/dev/null; python -c "
import json; j=json.load(open('$OUT/${TAG}_c3_chunk$ch.json')); print('config3 chunk $ch: %.1f M/s' % (j['value']/1e6))"; done
