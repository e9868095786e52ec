"""Two small wide shapes through the fast path once (for compute-sanitizer): python tests/devtools/wide_sanitize.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tt-irt_b200")); sys.path.insert(0, ROOT)
from tt_irt_py import synth, tt_irt
for ranks, ns, M in [((1, 70, 130, 9, 1), (12, 90, 75, 5), 700), ((1, 65, 67, 1), (73, 2, 74), 500), ((1, 96, 96, 1), (129, 129, 129), 900)]:
    rk = np.array(ranks, dtype=np.int64); ns = np.array(ns, dtype=np.int64); d = ns.size
    rng = np.random.default_rng(5)
    xs = np.concatenate([np.sort(rng.uniform(-1.5, 2.5, size=n)) for n in ns])
    c = rng.random(int((rk[:-1] * ns * rk[1:]).sum()))
    q = synth.make_q(M, d, seed=5)
    md = tt_irt.Model(ns, xs, rk, c)
    Z, l, ix = md.sample(q, mode=tt_irt.MODE_FAST, want_idx=True)
    print(ranks, "finite", bool(np.isfinite(l).all()), "launches", tt_irt.kernel_launches())
    md.close()
