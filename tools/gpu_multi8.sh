#!/bin/bash
# 8-GPU pass: sharding tests must run, bench at N=8 and N=4 (torchrun), in-process probe with two chunk sizes.
set -u
TAG=${1:-r02m8}; OUT=gpurun_out; mkdir -p $OUT
{ nvidia-smi -L; nproc; free -g; nvidia-smi topo -m; } > $OUT/${TAG}_env.log 2>&1
TTIRT_EXPECT_GPUS=8 timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_sqr_gpu.py -m gpu -x -q -k "multi_device or virtual_devices or sharding" -rs > $OUT/${TAG}_pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -4 $OUT/${TAG}_pytest_multi.log
for N in 8; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 3 --warmup 3 > $OUT/${TAG}_bench_n$N.json 2> $OUT/${TAG}_bench_n$N.err; echo "bench N=$N rc=$?"
python - <<P
import json
j=json.loads(open('$OUT/${TAG}_bench_n$N.json').read().strip().splitlines()[-1]); e=j['e2e']
print('N=$N value %.1f M/s | e2e one call %.1f M/s (%.1f ms) | no_upload %.1f M/s | ceiling' % (j['value']/1e6, e['value']/1e6, e['ms_per_step'], e['no_upload']['value']/1e6), {k:(round(v,1) if isinstance(v,float) else v) for k,v in e['host_copy_ceiling'].items() if k!='how'}, 'frac', e.get('frac_of_host_copy_ceiling'), 'same bits', e.get('same_bits_as_one_device_on_last_rows'))
P
done
PROBE_DEVICES=8 TTIRT_TRACE=1 timeout 300 python tools/multi_device_probe.py 26 > $OUT/${TAG}_probe_chunk20.json 2> $OUT/${TAG}_probe_chunk20.err; cat $OUT/${TAG}_probe_chunk20.json; tail -12 $OUT/${TAG}_probe_chunk20.err
PROBE_DEVICES=8 TTIRT_CHUNK=524288 timeout 300 python tools/multi_device_probe.py 26 > $OUT/${TAG}_probe_chunk19.json 2>/dev/null; cat $OUT/${TAG}_probe_chunk19.json
PROBE_DEVICES=8 TTIRT_BALANCE=static timeout 300 python tools/multi_device_probe.py 26 > $OUT/${TAG}_probe_static.json 2>/dev/null; cat $OUT/${TAG}_probe_static.json
