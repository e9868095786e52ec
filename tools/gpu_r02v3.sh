#!/bin/bash
# wide GEMM build variants (slice length, CTAs per SM), rebuilt on the box: usage bash tools/gpu_r02v3.sh <tag>
TAG=${1:-r02ae}
mkdir -p gpurun_out
cd tt-irt_b200
for V in "32 2" "16 2"; do
  set -- $V
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -DTTIRT_WIDE_KS=$1 -DTTIRT_WIDE_CTAS=$2 -c csrc/ttirt_wide.cu -o build/ttirt_wide.o && make > /dev/null 2>&1
  echo "== KS=$1 CTAS=$2 (build rc=$?)"
  for S in "8 129 128 17" "8 257 256 16" "6 65 96 18"; do
    (cd .. && timeout 120 python tests/devtools/wide_one.py $S 3) | tee -a ../gpurun_out/${TAG}_variants.log
  done
  if [ "$1" = "32" ]; then (cd .. && timeout 200 python tests/devtools/wide_check.py | python -c "
import sys, json
for l in sys.stdin:
    j = json.loads(l); print(j['ranks'], 'fails', j['fast']['fails'], 'flips', j['fast']['idx_flips'], 'z/tol %.3f' % j['fast']['z_max_over_tol'])"); fi
done
