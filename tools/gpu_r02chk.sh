#!/bin/bash
# wide path with every global address of the GEMMs checked against the operand extents (-DTTIRT_WIDE_CHECK, rebuilt on the box)
mkdir -p gpurun_out
cd tt-irt_b200
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -DTTIRT_WIDE_CHECK -c csrc/ttirt_wide.cu -o build/ttirt_wide.o && make > /dev/null 2>&1
echo "checked build rc=$?"
cd ..
timeout 300 python tests/devtools/wide_sanitize.py > gpurun_out/r02_wide_bounds_check.log 2>&1; echo "sanitize-shapes rc=$?"; tail -5 gpurun_out/r02_wide_bounds_check.log
timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "wide or random_ragged or strict_is_bitexact" >> gpurun_out/r02_wide_bounds_check.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_wide_bounds_check.log
grep -c "bounds violation" gpurun_out/r02_wide_bounds_check.log
