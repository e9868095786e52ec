"""Diagnose lPz parity outliers: per-dimension Z error, condition and density slope of the worst samples."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tt-irt_b200")); sys.path.insert(0, ROOT)
from tt_irt_py import synth, tt_irt
import oracle
d, n = 8, 17
ns, xs, rk, c = synth.make_tt(d, n, 8, seed=5, lo=0.0, hi=1.0)
M = 2 ** 14
q = np.random.default_rng(1).random([M, d]); q = np.reshape(q, [M, d], order="F")
Zo, lo, io, kap, gap, cond, lsens = oracle.oracle_run(ns, xs, rk, c, q, extras=True)
md = tt_irt.Model(ns, xs, rk, c)
for mode, nm in ((tt_irt.MODE_FAST, "fast"), (tt_irt.MODE_STRICT, "strict")):
    Z, l, ix = md.sample(q, mode=mode, want_idx=True)
    rel = np.abs(l - lo) / np.maximum(1.0, np.abs(lo))
    worst = np.argsort(-rel)[:4]
    print(nm, "max rel", rel.max(), "count>1e-12", int((rel > 1e-12).sum()), "count>5e-13", int((rel > 5e-13).sum()))
    for m in worst:
        print("  sample", m, "rel", rel[m], "lPz", lo[m], "dZ", np.abs(Z[m] - Zo[m]).max(), "cumcond*eps", (np.cumsum(cond[m]) * 2.2e-16).max(),
              "flip", bool((ix[m] != io[m]).any()))
if oracle.have_ref(32, "openblas"):
    Zr, lr = oracle.ref_run(ns, xs, rk, c, q, width=32, blas="openblas")
    rel = np.abs(lr - lo) / np.maximum(1.0, np.abs(lo))
    print("reference openblas vs oracle: max rel", rel.max(), "count>5e-13", int((rel > 5e-13).sum()))
    for m in np.argsort(-rel)[:3]:
        print("  sample", m, "rel", rel[m], "dZ", np.abs(Zr[m] - Zo[m]).max())
