"""Small shapes of all three fast-path classes through the host and device entry points (for compute-sanitizer)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tt-irt_b200"))
import numpy as np
import oracle
from oracle import parity
from tt_irt_py import synth, tt_irt
for (d, n, r, M) in [(4, 17, 8, 700), (3, 33, 32, 400), (3, 65, 64, 300), (3, 12, 40, 200), (4, 9, 5, 300)]:
    ns, xs, rk, c = synth.make_tt(d, n, r, seed=3)
    q = synth.make_q(M, d, seed=4)
    Zo, lo, io, kap, gap, cond, lsens = oracle.oracle_run(ns, xs, rk, c, q, extras=True)
    md = tt_irt.Model(ns, xs, rk, c)
    Z, l, idx = md.sample(q, want_idx=True)
    md.close()
    stats, fails = parity.compare(Z, l, idx, Zo, lo, io, cond, gap, lsens=lsens)
    print(d, n, r, M, "fails", fails, "flips", stats["idx_flips"], "z/tol %.3f" % stats["z_max_over_tol"])
    assert not fails
print("sanity ok")
