#!/bin/bash
set -u
TAG=${1:-r02q}; OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_parity_2p14.py -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/${TAG}_pytest.log
for lib in default t104; do
  if [ $lib = default ]; then unset TTIRT_LIBRARY; else export TTIRT_LIBRARY=$PWD/tools/exp_$lib.so; fi
  for sh in "40,33,32 22" "11,17,16 20" "24,33,32 22" "16,24,16 22"; do set -- $sh
    TTIRT_WALK=0 timeout 300 python bench.py --shape $1 --log2m $2 --steps 6 --warmup 3 --no-cpu --no-e2e --no-next-rows --no-other-configs > $OUT/${TAG}_${lib}_$1.json 2>/dev/null
    python -c "
import json; j=json.load(open('$OUT/${TAG}_${lib}_$1.json')); print('$lib shape $1 2^$2 (walk off): %.2f M/s  kernel frac %.4f' % (j['value']/1e6, j['roofline']['frac']))"
  done
done
