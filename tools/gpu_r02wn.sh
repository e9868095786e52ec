#!/bin/bash
# ncu full capture of the sorted walk kernel at d=11 n=17 r=16 M=2^20
mkdir -p gpurun_out
TTIRT_WALK_SORT=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:walk_sorted_kernel -s 1 -c 1 -f -o gpurun_out/r02_walk_sorted16 \
    python tests/devtools/wide_one.py 11 17 16 20 1 > gpurun_out/r02_walk_sorted16_ncu.log 2>&1; echo "full rc=$?"
