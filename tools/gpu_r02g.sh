#!/bin/bash
set -u
TAG=${1:-r02g}; OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q -rs > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/${TAG}_pytest_gpu.log
for lib in default gd1; do
  for sh in "40,33,32 22" "11,17,16 20" "24,33,32 22" "16,17,16 22"; do set -- $sh
    if [ $lib = gd1 ]; then export TTIRT_LIBRARY=$PWD/tools/exp_gd1.so; else unset TTIRT_LIBRARY; fi
    TTIRT_WALK=0 timeout 200 python bench.py --shape $1 --log2m $2 --steps 5 --warmup 3 --no-cpu --no-e2e --no-next-rows > $OUT/${TAG}_${lib}_$1.json 2>/dev/null
    python -c "
import json; j=json.load(open('$OUT/${TAG}_${lib}_$1.json')); print('$lib shape $1 2^$2 (walk off): %.1f M/s' % (j['value']/1e6))"
  done
done
unset TTIRT_LIBRARY
timeout 600 python bench.py --no-cpu --no-next-rows > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
python - <<'P'
import json
j=json.load(open('gpurun_out/r02g_bench.json')); r=j['roofline']
print('value %.2f M/s' % (j['value']/1e6), 'kernel frac %.4f' % r['frac'], 'whole %.4f' % r['whole_step_frac'], 'e2e %.2f' % (j['e2e']['value']/1e6), 'pageable %.2f' % (j['e2e']['pageable_numpy_value']/1e6))
for k,v in (j.get('other_configs') or {}).items(): print('   ', k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a in ('value','ms_per_step','launches_per_step','e2e_value','e2e_ms_per_call','e2e_ms_per_call_same_tt_again','frac_of_fp64_peak','e2e_matches_device_resident_bit_for_bit','unavailable')})
P
