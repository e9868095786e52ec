#!/bin/bash
set -u
TAG=${1:-r02c}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q -rs > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 $OUT/${TAG}_pytest_gpu.log
B="python bench.py --no-cpu --no-e2e --no-next-rows"
timeout 400 $B > $OUT/${TAG}_bench_main.json 2> $OUT/${TAG}_bench.err; echo "main rc=$?"
TTIRT_LIBRARY=$PWD/tools/exp_swpipe.so timeout 300 $B --no-other-configs > $OUT/${TAG}_bench_swpipe.json 2>> $OUT/${TAG}_bench.err; echo "swpipe rc=$?"
TTIRT_CHUNK=2097152 timeout 300 $B --no-other-configs > $OUT/${TAG}_bench_chunk21.json 2>> $OUT/${TAG}_bench.err; echo "chunk21 rc=$?"
TTIRT_CHUNK=4194304 timeout 300 $B --no-other-configs > $OUT/${TAG}_bench_chunk22.json 2>> $OUT/${TAG}_bench.err; echo "chunk22 rc=$?"
TTIRT_CHUNK=524288 timeout 300 $B --no-other-configs > $OUT/${TAG}_bench_chunk19.json 2>> $OUT/${TAG}_bench.err; echo "chunk19 rc=$?"
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02c_bench_*.json')):
    try:
        j=json.load(open(f)); r=j['roofline']
        print(f.split('bench_')[1], '%.2f M/s' % (j['value']/1e6), 'kernel frac %.4f' % r['frac'], 'avg ms %.4f' % r['kernel_avg_ms'], 'whole %.4f' % r['whole_step_frac'])
        for k,v in (j.get('other_configs') or {}).items(): print('   ', k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a in ('value','ms_per_step','launches_per_step','e2e_value','e2e_ms_per_call','frac_of_fp64_peak','e2e_matches_device_resident_bit_for_bit','unavailable')})
    except Exception as e: print(f, 'ERR', e)
P
# walk kernel on/off at the two small BASELINE shapes, device-resident
for sh in "11,17,16 20" "8,17,8 14" "8,17,8 20" "11,17,16 22"; do set -- $sh
  for w in 1 0; do TTIRT_WALK=$w timeout 200 python bench.py --shape $1 --log2m $2 --steps 10 --warmup 3 --no-cpu --no-next-rows > $OUT/${TAG}_walk${w}_$1_$2.json 2>> $OUT/${TAG}_bench.err
    python -c "
import json; j=json.load(open('$OUT/${TAG}_walk${w}_$1_$2.json')); print('walk=$w shape $1 2^$2: %.1f M/s device, e2e %.1f M/s, launches %d' % (j['value']/1e6, j['e2e']['value']/1e6, j['gpu_launches']))"
  done
done
# ncu: walk kernel and the r<=32 transition kernel
timeout 300 ncu --set full --clock-control none --import-source on -k regex:walk_kernel -s 4 -c 1 -f -o $OUT/${TAG}_walk16 python bench.py --shape 11,17,16 --log2m 20 --steps 2 --warmup 2 --no-e2e --no-cpu --no-next-rows > $OUT/${TAG}_ncu_walk.log 2>&1; echo "ncu walk rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:transition_kernel -s 60 -c 1 -f -o $OUT/${TAG}_trans32 python bench.py --shape 40,33,32 --log2m 20 --steps 1 --warmup 1 --no-e2e --no-cpu --no-next-rows > $OUT/${TAG}_ncu_t32.log 2>&1; echo "ncu t32 rc=$?"
