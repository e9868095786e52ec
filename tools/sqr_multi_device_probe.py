"""tt_irt_sqr through the C symbol with TTIRT_DEVICES = 1, 2, 4, ... on one M (rows sharded over the devices inside one
process, one host thread per device, cores replicated, no collective): samples/s end to end and bit-identity of the results.
usage: python tools/sqr_multi_device_probe.py [log2M] [d,n,r]  -> one JSON line"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tt-irt_b200"))
from tt_irt_py import synth, tt_irt, tt_irt_sqr  # noqa: E402

log2m = int(sys.argv[1]) if len(sys.argv) > 1 else 22
d, n, r = [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "32,65,64").split(",")]
M = 1 << log2m
ns, xs, rk, c = synth.make_tt(d, n, r, seed=5)
f = tt_irt.TTTensor(ns, rk, c)
q = synth.make_q(M, d, seed=3)
ndev = tt_irt.device_count()
out = {"shape": [d, n, r], "log2M": log2m, "devices_visible": ndev, "runs": []}
ref = None
k = 1
while k <= ndev:
    os.environ["TTIRT_DEVICES"] = str(k)
    tt_irt_sqr.tt_irt_sqr(xs, f, q[:1 << 16])                     # warm every device (context, modules, pool)
    best = None
    for _ in range(2):
        t0 = time.perf_counter()
        Z, l = tt_irt_sqr.tt_irt_sqr(xs, f, q)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    if ref is None:
        ref = (Z, l)
    out["runs"].append({"devices": k, "seconds": best, "samples_per_s": M / best,
                        "identical_to_one_device": bool(np.array_equal(Z, ref[0]) and np.array_equal(l, ref[1]))})
    k *= 2
print(json.dumps(out))
