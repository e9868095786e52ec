#!/bin/bash
set -u
TAG=${1:-r02t}; OUT=gpurun_out; mkdir -p $OUT
for cfg in "1048576 2" "2097152 2" "2097152 3" "4194304 3"; do set -- $cfg
TTIRT_CHUNK=$1 TTIRT_RAMP=$2 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-next-rows --no-other-configs > $OUT/${TAG}_c$1_r$2.json 2>/dev/null; python -c "
import json; j=json.load(open('$OUT/${TAG}_c$1_r$2.json')); print('chunk $1 ramp $2: value %.2f | pinned e2e %.2f M/s (%.1f ms), pageable %.2f M/s' % (j['value']/1e6, j['e2e']['value']/1e6, j['e2e']['ms_per_step'], j['e2e']['pageable_numpy_value']/1e6))"; done
