"""Development helper: per-dimension differences of one shape against the numpy oracle.
python tests/devtools/sqr_one.py d n r M [d n r M ...]"""
import os
import sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import sqr_check
from oracle.tt_irt_sqr_oracle import tt_irt_sqr_oracle
from tt_irt_py import synth, tt_irt_sqr
a = [int(v) for v in sys.argv[1:]]
for i in range(0, len(a), 4):
    d, n, r, M = a[i:i + 4]
    ns, xs, rk, c = sqr_check.make(d, n, r, "uniform", "uniform", False, 300 + d + n + r)
    q = synth.make_q(M, d, seed=11)
    Zo, lo, io, cond, gap, ls = tt_irt_sqr_oracle(ns, xs, rk, c, q, extras=True)
    md = tt_irt_sqr.SqrModel(ns, xs, rk, c)
    Z, lF, idx = md.sample(q, want_idx=True)
    md.close()
    print((d, n, r, M), "per-dim max |dZ|:", ["%.1e" % v for v in np.abs(Z - Zo).max(axis=0)], "rows bad:", int((np.abs(Z - Zo).max(axis=1) > 1e-9).sum()),
          "first bad rows:", np.nonzero(np.abs(Z - Zo).max(axis=1) > 1e-9)[0][:12].tolist(), flush=True)
