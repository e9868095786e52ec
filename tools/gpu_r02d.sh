#!/bin/bash
set -u
TAG=${1:-r02d}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q -rs > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 $OUT/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/${TAG}_smoke.log
timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
python - $TAG <<'P'
import json,sys
TAG=sys.argv[1]
j=json.load(open('gpurun_out/'+TAG+'_bench.json')); r=j['roofline']
print('value %.2f M/s' % (j['value']/1e6), 'kernel frac %.4f' % r['frac'], 'avg ms %.4f' % r['kernel_avg_ms'], 'whole %.4f' % r['whole_step_frac'], 'e2e %.2f' % (j['e2e']['value']/1e6), 'pageable %.2f' % (j['e2e']['pageable_numpy_value']/1e6))
for k,v in (j.get('other_configs') or {}).items(): print('   ', k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a in ('value','ms_per_step','launches_per_step','e2e_value','e2e_ms_per_call','e2e_ms_per_call_same_tt_again','frac_of_fp64_peak','e2e_matches_device_resident_bit_for_bit','unavailable')})
print(j['cpu_baseline'])
P
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > $OUT/${TAG}_bench_reference.json 2>> $OUT/${TAG}_bench.err; echo "ref rc=$?"
# launch list + full capture at the new chunk size (2^22 rows per launch)
python bench.py --log2m 22 --steps 1 --warmup 1 --no-e2e --no-cpu --no-next-rows --no-other-configs > $OUT/${TAG}_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --log2m 22 --steps 1 --warmup 1 --no-e2e --no-cpu --no-next-rows --no-other-configs > $OUT/${TAG}_ncu_list.log 2>&1
echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:transition_kernel -s 40 -c 1 -f -o $OUT/${TAG}_transition \
    python bench.py --log2m 22 --steps 1 --warmup 1 --no-e2e --no-cpu --no-next-rows --no-other-configs > $OUT/${TAG}_ncu_full.log 2>&1
echo "ncu full rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:walk_kernel -s 4 -c 1 -f -o $OUT/${TAG}_walk8 python bench.py --shape 8,17,8 --log2m 20 --steps 2 --warmup 2 --no-e2e --no-cpu --no-next-rows > $OUT/${TAG}_ncu_walk8.log 2>&1; echo "ncu walk8 rc=$?"
