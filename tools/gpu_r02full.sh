#!/bin/bash
# full GPU pass: all GPU tests, smoke, the default bench line.  usage: bash tools/gpu_r02full.sh <tag>
TAG=${1:-r02full}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
j = json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print("value %.2f M/s frac %.4f whole %.4f e2e %.2f launches %d clocks %s" % (j["value"] / 1e6, j["roofline"]["frac"], j["roofline"].get("whole_step_frac", 0), j["e2e"]["value"] / 1e6, j["gpu_launches"], j["clocks"]))
for k, v in j.get("other_configs", {}).items():
    print(" ", k, {kk: (round(vv, 4) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk in ("value", "e2e_value", "frac_of_fp64_peak", "launches_per_step", "unavailable", "e2e_ms_per_call")})
PY
tail -5 gpurun_out/${TAG}_bench.err
