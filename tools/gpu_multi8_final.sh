#!/bin/bash
# final 8-GPU pass of round 2: sharding tests must run, then the bench line at N=8 as the driver launches it
set -u
TAG=${1:-r02m8c}; OUT=gpurun_out; mkdir -p $OUT
TTIRT_EXPECT_GPUS=8 timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_sqr_gpu.py -m gpu -x -q -k "multi_device or virtual_devices or sharding" -rs > $OUT/${TAG}_pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -3 $OUT/${TAG}_pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 5 --warmup 3 > $OUT/${TAG}_bench_n8.json 2> $OUT/${TAG}_bench_n8.err; echo "bench N=8 rc=$?"
python - <<P
import json
j=json.loads(open('$OUT/${TAG}_bench_n8.json').read().strip().splitlines()[-1]); e=j['e2e']
print('N=8 value %.1f M/s | e2e one call %.1f M/s (%.1f ms) | no_upload %.1f M/s | ceiling' % (j['value']/1e6, e['value']/1e6, e['ms_per_step'], e['no_upload']['value']/1e6), {k:(round(v,1) if isinstance(v,float) else v) for k,v in e['host_copy_ceiling'].items() if k!='how'}, 'frac', e.get('frac_of_host_copy_ceiling'), 'same bits', e.get('same_bits_as_one_device_on_last_rows'))
P
