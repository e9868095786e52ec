"""Development checker of the squared-density path (GPU): sweep operands, samples and log-densities against the numpy
oracle, shape by shape, one JSON line each; optional timing of a large shape.  Test infrastructure (uses oracle/).

    python tests/devtools/sqr_check.py [--time d,n,r,log2M]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "tt-irt_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")

from oracle import parity                      # noqa: E402
from oracle.tt_irt_sqr_oracle import sqr_sweep, tt_irt_sqr_oracle   # noqa: E402
from tt_irt_py import synth, tt_irt_sqr        # noqa: E402

SHAPES = [
    # d, n, r, M, cores, grid, boundary-less cores, D
    (1, 9, 1, 500, "uniform", "uniform", False, None),
    (3, 5, 3, 400, "uniform", "uniform", False, None),
    (4, 9, 4, 1000, "uniform", "uniform", False, None),
    (8, 17, 8, 3000, "uniform", "uniform", False, None),
    (6, 17, 16, 3000, "uniform", "chebyshev", False, None),
    (5, 15, 6, 1500, "uniform", "uniform", True, None),       # grid carries boundary points the cores lack
    (6, 33, 32, 1500, "uniform", "uniform", False, 4),        # marginal of the first 4 variables
    (5, 12, 40, 1200, "uniform", "uniform", False, None),     # rank not a multiple of 8
    (4, 65, 64, 1024, "uniform", "uniform", False, None),
    (5, 20, 12, 2000, "normal", "uniform", False, None),      # signed cores
    (3, 72, 64, 600, "uniform", "chebyshev", False, None),
    (4, 6, 30, 800, "uniform", "uniform", False, None),       # n s < r: rank-deficient factor
]


def make(d, n, r, cores, grid, ext, seed):
    ns, xs, rk, c = synth.make_tt(d, n, r, seed=seed, cores=cores, grid=grid)
    if ext:   # the same grids plus two boundary points per dimension
        xs2 = []
        for k in range(d):
            x = xs[k * n:(k + 1) * n]
            xs2.append(np.concatenate([[x[0] - 0.7 * (x[1] - x[0])], x, [x[-1] + 0.4 * (x[-1] - x[-2])]]))
        xs = np.concatenate(xs2)
    return ns, xs, rk, c


def check(d, n, r, M, cores, grid, ext, D):
    ns, xs, rk, c = make(d, n, r, cores, grid, ext, 300 + d + n + r)
    D = d if D is None else D
    q = synth.make_q(M, D, seed=11)
    out = {"shape": [d, n, r, M, cores, grid, ext, D]}
    sw = sqr_sweep(ns, xs, rk, c)
    md = tt_irt_sqr.SqrModel(ns, xs, rk, c)
    try:
        gerr, rerr = 0.0, 0.0
        for k in range(d):
            G, RR = md.sweep(k)
            Go = sw["P"][k]
            gerr = max(gerr, float(np.abs(G - Go).max() / np.abs(Go).max()))
            if k > 0:
                RRo = sw["R"][k] @ sw["R"][k].T
                rerr = max(rerr, float(np.abs(RR - RRo).max() / np.abs(RRo).max()))
        out["gram_rel"] = gerr
        out["rr_rel"] = rerr
        Zo, lo, io, cond, gap, lsens = tt_irt_sqr_oracle(ns, xs, rk, c, q, extras=True)
        Z, lF, idx = md.sample(q, want_idx=True)
        st, fails = parity.compare(Z, lF, idx, Zo, lo, io, cond, gap, lsens)
        out.update(st)
        out["fails"] = fails
        out["finite"] = bool(np.isfinite(Z).all())
    finally:
        md.close()
    return out


def timing(d, n, r, log2m):
    import ctypes
    ns, xs, rk, c = synth.make_tt(d, n, r, seed=5)
    M = 1 << log2m
    q = synth.make_q(M, d, seed=3)
    t0 = time.time()
    md = tt_irt_sqr.SqrModel(ns, xs, rk, c)
    t_model = time.time() - t0
    try:
        md.sample(q[:4096])
        md.profile_enable(True)
        t0 = time.time()
        Z, lF = md.sample(q)
        t_host = time.time() - t0
        ms, nl, fl = md.profile_read()
        md.profile_enable(False)
    finally:
        md.close()
    W = tt_irt_sqr.flops_per_sample(ns, rk)
    return {"time_shape": [d, n, r, log2m], "model_create_s": t_model, "host_call_s": t_host, "samples_per_s_host": M / t_host,
            "pdf_kernel_ms": ms, "pdf_launches": nl, "pdf_tflops": fl / ms / 1e9 if ms > 0 else None,
            "flops_per_sample": W, "algorithmic_tflops_host": W * M / t_host / 1e12, "finite": bool(np.isfinite(Z).all() and np.isfinite(lF).all())}


if __name__ == "__main__":
    args = sys.argv[1:]
    if args and args[0] == "--time":
        for spec in args[1:]:
            d, n, r, l = [int(v) for v in spec.split(",")]
            print(json.dumps(timing(d, n, r, l)), flush=True)
        sys.exit(0)
    for s in SHAPES:
        try:
            print(json.dumps(check(*s)), flush=True)
        except Exception as e:   # keep going: one GPU call should report every shape
            print(json.dumps({"shape": list(s), "error": repr(e)}), flush=True)
