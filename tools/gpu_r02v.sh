#!/bin/bash
set -u
TAG=${1:-r02v}; OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "ramp or golden or full_size or wrapper or repeated or stale" 2>&1 | tail -2
for ns in 0 1 0 1; do if [ $ns = 1 ]; then export TTIRT_NO_STAGING=1; else unset TTIRT_NO_STAGING; fi
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-next-rows --no-other-configs > $OUT/${TAG}_ns$ns.json 2>/dev/null; python -c "
import json; j=json.load(open('$OUT/${TAG}_ns$ns.json')); print('no_staging=$ns: pinned e2e %.2f M/s (%.1f ms), pageable %.2f M/s' % (j['e2e']['value']/1e6, j['e2e']['ms_per_step'], j['e2e']['pageable_numpy_value']/1e6))"; done
unset TTIRT_NO_STAGING
TTIRT_TRACE=1 timeout 300 python tools/e2e_trace.py 2>&1 | grep -E "model_load|shape" | tail -8
