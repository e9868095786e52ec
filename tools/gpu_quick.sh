#!/bin/bash
# Quick GPU check of a kernel change, every step under its own timeout:
#   parity on all dev shapes, the GPU test suite, a short bench line, the phase-timing build if present.
TAG=${1:-q}
timeout 120 python tests/devtools/dev_check.py 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: j=json.loads(l)
    except Exception: print(l[:300]); continue
    print(j['shape'], 'strict', j['strict']['Z_bitexact'], 'fast fails', j['fast']['fails'], 'flips', j['fast']['idx_flips'], 'z/tol %.3f' % j['fast']['z_max_over_tol'], 'lpz %.2e' % j['fast']['lpz_max_rel'])
" | tee gpurun_out/${TAG}_dev_check.log
timeout 180 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 120 python bench.py --no-cpu --no-e2e --steps 2 > gpurun_out/${TAG}_bench.json 2>gpurun_out/${TAG}_bench.err
python -c "
import json
j=json.load(open('gpurun_out/${TAG}_bench.json'))
print('bench %.2f M/s frac %.3f kernel_ms %.4f' % (j['value']/1e6, j['roofline']['frac'], j['roofline']['kernel_avg_ms']))" || tail -3 gpurun_out/${TAG}_bench.err
[ -f tools/pt_w8.so ] && timeout 60 python tools/phase_timing.py tools/pt_w8.so 20 2>&1 | tee gpurun_out/${TAG}_pt.log
