#!/bin/bash
# compute-sanitizer (memcheck, racecheck) over the wide path's kernels
mkdir -p gpurun_out
for TOOL in memcheck racecheck; do
  TTIRT_GRAPHS=0 timeout 400 compute-sanitizer --tool $TOOL --kernel-regex kns=wide python tests/devtools/wide_sanitize.py > gpurun_out/r02_sanitize_$TOOL.log 2>&1
  echo "$TOOL rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|finite|Error|hazard" gpurun_out/r02_sanitize_$TOOL.log | head -12
done
