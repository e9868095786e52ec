"""Development check on a GPU box: sweep, strict and fast paths against the oracle on several shapes."""
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tt-irt_b200")); sys.path.insert(0, ROOT)
from tt_irt_py import synth, tt_irt
import oracle
from oracle import parity

shapes = [(4, 9, 4, 1000, "uniform", "uniform"), (8, 17, 8, 5000, "uniform", "uniform"), (8, 17, 16, 4096, "uniform", "chebyshev"),
          (11, 17, 16, 4096, "uniform", "uniform"), (40, 33, 32, 2048, "uniform", "uniform"), (32, 65, 64, 2048, "uniform", "uniform"),
          (6, 65, 8, 3000, "uniform", "uniform"), (5, 12, 40, 1500, "uniform", "uniform"), (6, 20, 24, 777, "normal", "uniform")]
if len(sys.argv) > 1:
    shapes = shapes[:int(sys.argv[1])]
for (d, n, r, M, cores, grid) in shapes:
    ns, xs, rk, c = synth.make_tt(d, n, r, seed=11, cores=cores, grid=grid)
    q = synth.make_q(M, d, seed=12)
    Zo, lo, io, kap, gap, cond, lsens = oracle.oracle_run(ns, xs, rk, c, q, extras=True)
    md = tt_irt.Model(ns, xs, rk, c)
    P, Mg = md.sweep(); Po, Mgo = oracle.oracle_sweep(ns, xs, rk, c)
    sweep_ok = all(np.array_equal(a, b) for a, b in zip(P, Po)) and all(np.array_equal(a, b) for a, b in zip(Mg[:-1], Mgo[:-1]))
    t = time.time(); Zs, ls, isx = md.sample(q, mode=tt_irt.MODE_STRICT, want_idx=True); ts = time.time() - t
    t = time.time(); Zf, lf, ifx = md.sample(q, mode=tt_irt.MODE_FAST, want_idx=True); tf = time.time() - t
    st_s, f_s = parity.compare(Zs, ls, isx, Zo, lo, io, cond, gap, lsens=lsens)
    st_f, f_f = parity.compare(Zf, lf, ifx, Zo, lo, io, cond, gap, lsens=lsens)
    print(json.dumps({"shape": [d, n, r, M, cores, grid], "sweep_bitexact": bool(sweep_ok),
                      "strict": {"Z_bitexact": bool(np.array_equal(Zs, Zo)), "idx_equal": bool(np.array_equal(isx, io)), **st_s, "fails": f_s, "sec": round(ts, 3)},
                      "fast": {**st_f, "fails": f_f, "sec": round(tf, 3)}}))
    md.close()
