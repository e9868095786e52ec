#!/bin/bash
# sorted walk kernel (TTIRT_WALK_SORT=1) against the plain one: parity tests + device-resident timing
TAG=${1:-r02ws}
mkdir -p gpurun_out
for SORT in 0 1; do
  echo "== TTIRT_WALK_SORT=$SORT"
  for S in "11 17 16 20" "11 17 16 22" "8 17 8 20" "8 17 8 14"; do
    TTIRT_WALK_SORT=$SORT timeout 120 python tests/devtools/wide_one.py $S 5 | tee -a gpurun_out/${TAG}_timing.log
  done
done
TTIRT_WALK_SORT=1 timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_parity_2p14.py -x -q -m gpu -k "walk or strict_is_bitexact or 2p14 or golden or kat" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${TAG}_pytest.log
