#!/bin/bash
# r02y: wide path after a change: parity check + perf probe + launch list
TAG=${1:-r02y}
mkdir -p gpurun_out
timeout 300 python tests/devtools/wide_check.py --perf > gpurun_out/${TAG}_wide_check.log 2> gpurun_out/${TAG}_wide_check.err
echo "wide_check rc=$?"
python - <<PY
import json
for l in open("gpurun_out/${TAG}_wide_check.log"):
    j = json.loads(l)
    if "perf" in j:
        print(j["perf"], "fast %.3f ms %.2f TF (%.3f)  strict %.1f ms" % (j["fast"]["ms"], j["fast"]["tflops"], j["fast"]["frac_of_dmma_peak_37.1"], j["strict"]["ms"]))
    else:
        print(j["ranks"], j["ns"], "fails", j["fast"]["fails"], "flips", j["fast"]["idx_flips"], "z/tol %.3f" % j["fast"]["z_max_over_tol"], "strict ok", j["strict_bitexact"])
PY
tail -c 600 gpurun_out/${TAG}_wide_check.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python tests/devtools/wide_one.py 8 129 128 17 1 > gpurun_out/${TAG}_ncu_list.log 2>&1; echo "list rc=$?"
python tools/launch_summary.py gpurun_out/${TAG}_launches.csv | grep -i "wide\|scatter\|segment"
