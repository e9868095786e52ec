"""Generates tests/golden/matlab_*.npz: outputs of the reference's own MATLAB sources, executed here from their source files.

This image has neither Matlab nor Octave, so the reference's Matlab-only routines on either side of `tt_irt1` could not be run
and their numpy restatements under oracle/ were "parity unpinned".  This script closes that gap as far as it can be closed
without Matlab:

  * oracle/mlite.py, a small interpreter for the subset of the Matlab language these files are written in, executes the
    UNMODIFIED text of
        matlab/samplers/qmcnodes.m, randref.m                      (seeds before the path,        SURVEY section 8(f) rank 3)
        matlab/samplers/essinv.m, hellinger.m, iw_prune.m, mcmc_prune.m   (consumers after the path,      rank 2)
        matlab/samplers/tt_irt_sqr.m, tt_rt_sqr.m                  (squared-density transforms,    rank 4)
        matlab/samplers/tt_dirt_sample.m, tt_dirt_inverse.m        (their callers, the DIRT sample / inverse loops)
        matlab/samplers/tt_irt_lin.m                               (the Matlab implementation of tt_irt1's own transform: the
                                                                    interpreter's cross-check against the compiled reference C)
        matlab/samplers/tt_irt_debias.m                            (the top-level driver: seeds -> sampler -> exact density ->
                                                                    MH / IW correction; oracle/matlab_driver.py)
    read from /root/reference at generation time (never copied into this repo);
  * the MEX function those transforms call, matlab/utils/tracemult.c, is the reference's own C, compiled unmodified against a
    stand-in mex.h (oracle/mexstub/, oracle/Makefile -> oracle/_ref/libref_tracemult.so) and called through oracle/mex_host.py.

What stands in for Matlab itself: numpy / LAPACK for elementary functions, sums, matrix products and qr (so a real Matlab run can
differ in the last bits; the tests that use these fixtures carry tolerances), a recorded stream for rand, and a synthetic
generating-vector table for qmcnodes' load() (the reference does not ship its lattice file).

Run in the build container (needs /root/reference and `make -C oracle`):   python tests/golden/make_golden_matlab.py
tests/test_matlab_pins.py checks the oracles against these fixtures and, when the reference is present, re-runs this generation
and demands the committed bits.
"""
import os
import sys

os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")   # as tests/conftest.py: one BLAS thread, reproducible bits

import numpy as np  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tt-irt_b200"))
REF_M = "/root/reference/matlab/samplers"
LATTICE_FILE = "lattice-39102-1024-1048576.3600.txt"          # qmcnodes.m:4 (not vendored by the reference)

SQR_CASES = [
    # name, d, n (nodes of the cores), r, M, seed, cores, grid, boundary nodes in the cores?, columns of q (None: d)
    ("tiny_d3_n5_r3", 3, 5, 3, 200, 1, "uniform", "uniform", True, None),
    ("d4_n9_r4_cheb", 4, 9, 4, 500, 2, "uniform", "chebyshev", True, None),
    ("shock_d6_n17_r8", 6, 17, 8, 1000, 3, "uniform", "uniform", True, None),
    ("noboundary_d4_n7_r5", 4, 7, 5, 400, 4, "uniform", "uniform", False, None),
    ("marginal_d5_n9_r4", 5, 9, 4, 300, 5, "uniform", "uniform", True, 3),
    ("signed_d4_n9_r4", 4, 9, 4, 300, 6, "normal", "uniform", True, None),
]


def available():
    from oracle import mex_host
    return os.path.isdir(REF_M) and mex_host.available()


def helper_inputs():
    """Seeded inputs of the helper fixtures (regenerated identically by the tests)."""
    rng = np.random.default_rng(20260119)
    M = 2000
    lFex = rng.normal(size=(M, 1)) * 0.7 - 3.0
    qoi = rng.normal(size=(M, 1))
    lFapp = lFex + rng.normal(size=(M, 1)) * 0.3 + 0.2
    d, l = 5, 10
    table = np.column_stack([np.arange(1, 40), rng.integers(1, 2 ** 20, size=39) | 1]).astype(np.float64)   # (index, generating vector)
    shift = rng.random((d, 1))
    u = rng.random((300, 4))
    Mm = 3000
    us = rng.random(Mm)
    y = rng.normal(size=(Mm, 3))
    lfe = np.hstack([rng.normal(size=(Mm, 1)), rng.normal(size=(Mm, 1))])
    lfa = lfe[:, :1] + rng.normal(size=(Mm, 1)) * 0.8
    return {"lFex": np.hstack([lFex, qoi]), "lFapp": lFapp, "d": d, "l": l, "table": table, "shift": shift, "u": u,
            "mh_u": us, "mh_y": y, "mh_lFex": lfe, "mh_lFapp": lfa}


def run_helpers():
    from oracle import mlite
    inp = helper_inputs()
    out = {}
    ip = mlite.Interp()
    for f in ("essinv", "hellinger", "iw_prune"):
        ip.load_file(os.path.join(REF_M, f + ".m"))
    out["essinv"] = ip.call("essinv", [inp["lFex"][:, :1], inp["lFapp"]])[0]
    out["hellinger"] = ip.call("hellinger", [inp["lFex"][:, :1], inp["lFapp"]])[0]
    lw, isstd, mx, err1 = ip.call("iw_prune", [inp["lFex"], inp["lFapp"]], 4)
    out.update(iw_lFex=lw, iw_isstd=isstd, iw_max_ratio=mx, iw_err1=err1)
    ipq = mlite.Interp(rand_stream=lambda shape: inp["shift"].reshape(shape, order="F"), files={LATTICE_FILE: inp["table"]})
    ipq.load_file(os.path.join(REF_M, "qmcnodes.m"))
    out["qmc_Y"] = ipq.call("qmcnodes", [inp["d"], inp["l"]])[0]
    ipr = mlite.Interp()
    ipr.load_file(os.path.join(REF_M, "randref.m"))
    for tag, ref in (("normal", "Normal"), ("normal3", "normal 3"), ("n2p5", "n2.5"), ("uni", "UNI")):
        out["randref_" + tag] = ipr.call("randref", [mlite.MStr(ref), inp["u"]])[0]
    it = iter(inp["mh_u"])
    ipm = mlite.Interp(rand_stream=lambda shape: np.array([[next(it)]]))
    ipm.load_file(os.path.join(REF_M, "mcmc_prune.m"))
    y, lfe, lfa, nrej, dist = ipm.call("mcmc_prune", [inp["mh_y"], inp["mh_lFex"], inp["mh_lFapp"]], 5)
    out.update(mh_y=y, mh_lFex=lfe, mh_lFapp=lfa, mh_num_of_rejects=nrej, mh_rej_distribution=dist)
    return {k: np.asarray(v, dtype=np.float64) for k, v in out.items()}


def sqr_inputs(case):
    from tt_irt_py import synth
    name, d, n, r, M, seed, cores, grid, boundary, D = case
    ns, xs, rk, c = synth.make_tt(d, n, r, seed=seed, cores=cores, grid=grid)
    if not boundary:
        # the cores hold the interior nodes only; the grid keeps its two boundary points per dimension (tt_irt_sqr.m:33-36)
        xs = np.concatenate([np.concatenate([[-1.25], xs[k * n:(k + 1) * n], [1.25]]) for k in range(d)])
    q = synth.make_q(M, d if D is None else D, seed=seed + 100)
    q[0, :] = 0.0
    q[1, :] = 1.0
    return ns, xs, rk, c, q


def _cell_of_cores(mlite, ns, rk, c):
    d = len(ns)
    f = mlite.MCell((d, 1))
    off = 0
    for k in range(d):
        sz = int(rk[k] * ns[k] * rk[k + 1])
        f.a[k, 0] = np.asfortranarray(c[off:off + sz].reshape((int(rk[k]), int(ns[k]), int(rk[k + 1])), order="F"))
        off += sz
    return f


def run_sqr(case):
    """[xq, lFapp] = tt_irt_sqr(xsf, f, q) and [q2, lF2] = tt_rt_sqr(xsf, f, xq) from the reference sources."""
    from oracle import mex_host, mlite
    ns, xs, rk, c, q = sqr_inputs(case)
    ip = mlite.Interp(externals={"tracemult": lambda ip_, args, nargout: mex_host.ref_tracemult(*args)})
    ip.load_file(os.path.join(REF_M, "tt_irt_sqr.m"))
    ip.load_file(os.path.join(REF_M, "tt_rt_sqr.m"))
    xsf = np.asarray(xs, dtype=np.float64).reshape(-1, 1)
    xq, lF = ip.call("tt_irt_sqr", [xsf, _cell_of_cores(mlite, ns, rk, c), np.asfortranarray(q)], 2)
    q2, lF2 = ip.call("tt_rt_sqr", [xsf, _cell_of_cores(mlite, ns, rk, c), np.asfortranarray(xq)], 2)
    return {"xq": np.asarray(xq), "lFapp": np.asarray(lF).reshape(-1), "rt_q": np.asarray(q2), "rt_lFapp": np.asarray(lF2).reshape(-1)}


DIRT_CASES = [("uni", "uni", 0.0, 1.0), ("normal3", "Normal 3", -3.0, 3.0)]


def dirt_inputs(case):
    """A synthetic two-level DIRT (levels 1.. share the grid IRTstruct.x, level 0 has its own x0) and its seeds."""
    from tt_irt_py import synth
    tag, reference, lo, hi = case
    d, n, r, nl, M = 3, 9, 4, 2, 300
    levels = [synth.make_tt(d, n, r, seed=11 + j, lo=lo, hi=hi) for j in range(nl + 1)]
    levels = [levels[0]] + [(l[0], levels[1][1], l[2], l[3]) for l in levels[1:]]
    levels[0] = synth.make_tt(d, n, r, seed=5, lo=-2.0, hi=2.5)
    rng = np.random.default_rng(3)
    q = rng.random((M, d)) if reference[0].lower() == "u" else np.clip(rng.normal(size=(M, d)), -2.9, 2.9)
    return levels, np.asfortranarray(q), reference


def run_dirt(case):
    """[z, lFapp] = tt_dirt_sample(IRTstruct, q) and [q2, lF2] = tt_dirt_inverse(IRTstruct, z) from the reference sources."""
    from oracle import mex_host, mlite
    levels, q, reference = dirt_inputs(case)
    ip = mlite.Interp(externals={"tracemult": lambda ip_, args, nargout: mex_host.ref_tracemult(*args)})
    for f in ("tt_irt_sqr", "tt_rt_sqr", "tt_dirt_sample", "tt_dirt_inverse"):
        ip.load_file(os.path.join(REF_M, f + ".m"))
    nl = len(levels) - 1
    F = mlite.MCell((nl, 1))
    for j in range(1, nl + 1):
        F.a[j - 1, 0] = _cell_of_cores(mlite, levels[j][0], levels[j][2], levels[j][3])
    st = {"beta": np.arange(nl + 1, dtype=np.float64).reshape(1, -1), "reference": mlite.MStr(reference),
          "crossmethod": mlite.MStr("amen_cross_s"), "interpolation": mlite.MStr("spline"),
          "x": np.asarray(levels[1][1], dtype=np.float64).reshape(-1, 1), "x0": np.asarray(levels[0][1], dtype=np.float64).reshape(-1, 1),
          "F": F, "F0": _cell_of_cores(mlite, levels[0][0], levels[0][2], levels[0][3])}
    z, lF = ip.call("tt_dirt_sample", [st, q], 2)
    q2, lF2 = ip.call("tt_dirt_inverse", [st, z], 2)
    return {"z": np.asarray(z), "lFapp": np.asarray(lF).reshape(-1), "inv_q": np.asarray(q2), "inv_lFapp": np.asarray(lF2).reshape(-1)}


LIN_CASES = [("d4_n9_r4", 4, 9, 4, 500, 2, "uniform"), ("shock_d8_n17_r8", 8, 17, 8, 1500, 3, "uniform"), ("d5_n17_r16_cheb", 5, 17, 16, 800, 4, "chebyshev")]


def lin_inputs(case):
    from tt_irt_py import synth
    name, d, n, r, M, seed, grid = case
    ns, xs, rk, c = synth.make_tt(d, n, r, seed=seed, grid=grid)
    return ns, xs, rk, c, synth.make_q(M, d, seed=seed + 50)


def run_lin(case):
    """[xq, lFapp] = tt_irt_lin(xsf, f, q): the reference's MATLAB implementation of the transform its C routine tt_irt1 computes."""
    from oracle import mex_host, mlite
    ns, xs, rk, c, q = lin_inputs(case)
    ip = mlite.Interp(externals={"tracemult": lambda ip_, args, nargout: mex_host.ref_tracemult(*args)})
    ip.load_file(os.path.join(REF_M, "tt_irt_lin.m"))
    xq, lF = ip.call("tt_irt_lin", [np.asarray(xs, dtype=np.float64).reshape(-1, 1), _cell_of_cores(mlite, ns, rk, c), np.asfortranarray(q)], 2)
    return {"xq": np.asarray(xq), "lFapp": np.asarray(lF).reshape(-1)}


def run_debias():
    """The reference's top-level driver tt_irt_debias.m end to end (oracle/matlab_driver.py), its sampler call patched to the MEX
    gateway as install.m:160-169 does, served by the reference's own gateway + C: both corrections."""
    from oracle import matlab_driver
    compiled = matlab_driver.compile_reference()
    out = {}
    for corr in ("mcmc", "iw"):
        r = matlab_driver.run_debias("reference", corr, compiled)
        out.update({corr + "_" + k: v for k, v in r.items()})
    return out


def tracemult_inputs():
    rng = np.random.default_rng(77)
    A = rng.normal(size=(3, 4, 50))
    B = rng.normal(size=(4, 5, 7))
    j = rng.integers(1, 8, size=(50, 1)).astype(np.float64)
    A2 = rng.normal(size=(60, 9))
    j2 = rng.integers(1, 10, size=(60, 1)).astype(np.float64)
    return A, B, j, A2, j2


def run_tracemult():
    from oracle import mex_host
    A, B, j, A2, j2 = tracemult_inputs()
    return {"C": mex_host.ref_tracemult(A, j, B), "C2": mex_host.ref_tracemult(A2, j2)}


def main():
    if not available():
        raise SystemExit("needs /root/reference and oracle/_ref/libref_tracemult.so (make -C oracle)")
    np.savez_compressed(os.path.join(HERE, "matlab_helpers.npz"), **run_helpers())
    np.savez_compressed(os.path.join(HERE, "matlab_tracemult.npz"), **run_tracemult())
    for case in SQR_CASES:
        np.savez_compressed(os.path.join(HERE, "matlab_sqr_%s.npz" % case[0]), **run_sqr(case))
        print("wrote", case[0])
    for case in DIRT_CASES:
        np.savez_compressed(os.path.join(HERE, "matlab_dirt_%s.npz" % case[0]), **run_dirt(case))
        print("wrote dirt", case[0])
    for case in LIN_CASES:
        np.savez_compressed(os.path.join(HERE, "matlab_lin_%s.npz" % case[0]), **run_lin(case))
        print("wrote lin", case[0])
    np.savez_compressed(os.path.join(HERE, "matlab_debias.npz"), **run_debias())
    print("wrote debias")


if __name__ == "__main__":
    main()
