#!/bin/bash
set -u
TAG=${1:-r02s}; OUT=gpurun_out; mkdir -p $OUT
for r in 2 3 0; do TTIRT_RAMP=$r timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-next-rows > $OUT/${TAG}_ramp$r.json 2>/dev/null; python -c "
import json; j=json.load(open('$OUT/${TAG}_ramp$r.json')); print('ramp $r: pinned e2e %.2f M/s (%.1f ms), pageable %.2f M/s' % (j['e2e']['value']/1e6, j['e2e']['ms_per_step'], j['e2e']['pageable_numpy_value']/1e6), {k[:10]:round(v['e2e_ms_per_call'],3) for k,v in j['other_configs'].items()})"; done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
