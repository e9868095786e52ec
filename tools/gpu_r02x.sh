#!/bin/bash
# wide path: perf probe, launch list, one full ncu capture of both GEMMs.  usage: bash tools/gpu_r02x.sh <tag>
TAG=${1:-r02x}
mkdir -p gpurun_out
timeout 120 python tests/devtools/wide_one.py 8 129 128 17 3 > gpurun_out/${TAG}_plain.log 2>&1; echo "plain rc=$?"; cat gpurun_out/${TAG}_plain.log
timeout 120 python tests/devtools/wide_one.py 8 257 256 16 3; timeout 120 python tests/devtools/wide_one.py 6 65 96 18 3
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python tests/devtools/wide_one.py 8 129 128 17 1 > gpurun_out/${TAG}_ncu_list.log 2>&1; echo "list rc=$?"
python tools/launch_summary.py gpurun_out/${TAG}_launches.csv | grep -i "wide\|scatter\|segment"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wide_gemm_kernel -s 6 -c 2 -f -o gpurun_out/${TAG}_wide_gemm \
    python tests/devtools/wide_one.py 8 129 128 17 1 > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "full rc=$?"
