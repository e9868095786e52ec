#!/bin/bash
# wide GEMM build variants (CTAs per SM, ring stages), rebuilt on the box: usage bash tools/gpu_r02v2.sh <tag>
TAG=${1:-r02ab}
mkdir -p gpurun_out
cd tt-irt_b200
for V in "3 2" "4 2" "3 3"; do
  set -- $V
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -DTTIRT_WIDE_CTAS=$1 -DTTIRT_WIDE_STAGES=$2 -c csrc/ttirt_wide.cu -o build/ttirt_wide.o && make > /dev/null 2>&1
  echo "== CTAS=$1 STAGES=$2 (build rc=$?)"
  for S in "8 129 128 17" "8 257 256 16" "6 65 96 18" "6 129 72 18"; do
    (cd .. && timeout 120 python tests/devtools/wide_one.py $S 3) | tee -a ../gpurun_out/${TAG}_variants.log
  done
done
