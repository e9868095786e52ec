#!/bin/bash
# r02w: first GPU pass of the wide path (csrc/ttirt_wide.cu): parity + perf probe, then the GPU tests that touch it
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02w_env.log 2>&1
timeout 300 python tests/devtools/wide_check.py --perf > gpurun_out/r02w_wide_check.log 2> gpurun_out/r02w_wide_check.err
echo "wide_check rc=$?"
tail -c 3000 gpurun_out/r02w_wide_check.log
tail -c 1500 gpurun_out/r02w_wide_check.err
timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "wide or random_ragged or strict_is_bitexact" > gpurun_out/r02w_pytest.log 2>&1
echo "pytest rc=$?"
tail -15 gpurun_out/r02w_pytest.log
