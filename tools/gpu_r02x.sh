#!/bin/bash
# r02x: launch list and one full ncu capture of the wide path's update GEMM
mkdir -p gpurun_out
timeout 120 python tests/devtools/wide_one.py 8 129 128 17 3 > gpurun_out/r02x_plain.log 2>&1; echo "plain rc=$?"; cat gpurun_out/r02x_plain.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02x_launches.csv \
    python tests/devtools/wide_one.py 8 129 128 17 1 > gpurun_out/r02x_ncu_list.log 2>&1; echo "list rc=$?"
python tools/launch_summary.py gpurun_out/r02x_launches.csv
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wide_gemm_kernel -s 6 -c 2 -f -o gpurun_out/r02x_wide_gemm \
    python tests/devtools/wide_one.py 8 129 128 17 1 > gpurun_out/r02x_ncu_full.log 2>&1; echo "full rc=$?"
