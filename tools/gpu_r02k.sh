#!/bin/bash
set -u
TAG=${1:-r02k}; OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "walk or virtual or ragged" > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/${TAG}_pytest.log
for t in 4 8 12; do TTIRT_COPY_THREADS=$t timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu --no-next-rows --no-other-configs > $OUT/${TAG}_ct$t.json 2>/dev/null; python -c "
import json; j=json.load(open('$OUT/${TAG}_ct$t.json')); print('copy threads $t: pinned e2e %.2f M/s, pageable %.2f M/s' % (j['e2e']['value']/1e6, j['e2e']['pageable_numpy_value']/1e6))"; done
