#!/bin/bash
set -u
TAG=${1:-r02r}; OUT=gpurun_out; mkdir -p $OUT
for r in 0 1 0 1; do TTIRT_RAMP=$r timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-next-rows --no-other-configs > $OUT/${TAG}_ramp$r.json 2>/dev/null; python -c "
import json; j=json.load(open('$OUT/${TAG}_ramp$r.json')); print('ramp $r: pinned e2e %.2f M/s (%.1f ms), pageable %.2f M/s' % (j['e2e']['value']/1e6, j['e2e']['ms_per_step'], j['e2e']['pageable_numpy_value']/1e6))"; done
TTIRT_RAMP=1 timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "full_size or leading or repeated" 2>&1 | tail -2
