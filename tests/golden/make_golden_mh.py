"""Generates tests/golden/shock_mh_D2.npz: the reference's own Python driver for BASELINE configs[0]
(/root/reference/python/test_shock_absorber_tt.py) run here on the D=2 shock-absorber posterior, with the pieces that need
the absent `ttpy` package replaced as SURVEY.md section 8(d) proposes:

  * the posterior density is evaluated on the full 17^4 grid with the REFERENCE's own functions (logf_prior,
    logL_weibull, crossfun -- executed from the reference source at generation time, never copied into this repo) and
    compressed by an exact-to-eps TT-SVD instead of `rect_cross.cross` (ttpy absent);
  * `tt_irt1` is the CPU oracle (oracle/tt_irt1_oracle.c, pinned bit for bit to the reference C);
  * the independence Metropolis-Hastings loop is the REFERENCE's own loop (source lines 164-171), executed from the
    reference file with a seeded np.random; the row index rides along as an extra column of Z so that the loop's
    in-place copies reveal which sample ends up where.

Run in the build container (needs /root/reference):   python tests/golden/make_golden_mh.py
The fixture pins oracle/samplers_oracle.py::mcmc_prune and oracle/shock_absorber_oracle.py against the reference's
Python, and gives the GPU path an end-to-end check: seeds -> tt_irt1 -> exact density -> MH prune -> quantile of interest.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tt-irt_b200"))
REF = "/root/reference/python/test_shock_absorber_tt.py"


def ref_lines(lo, hi, must_contain):
    """Source lines lo..hi (1-based, inclusive) of the reference driver, checked against a marker so that a changed
    reference file is noticed instead of silently executing something else."""
    with open(REF) as f:
        src = f.read().splitlines()
    import textwrap
    seg = textwrap.dedent("\n".join(src[lo - 1:hi]) + "\n")
    for m in must_contain:
        if m not in seg:
            raise RuntimeError("reference %s:%d-%d does not contain %r any more" % (REF, lo, hi, m))
    return seg


def tt_svd(A, eps):
    """TT-SVD (Oseledets 2011, Alg. 1) of a full tensor with relative Frobenius accuracy eps; returns the cores as a list
    of (r_k, n_k, r_{k+1}) arrays."""
    d = A.ndim
    ns = A.shape
    delta = eps / np.sqrt(d - 1) * np.linalg.norm(A)
    cores, r = [], 1
    C = A.reshape(ns[0], -1)
    for k in range(d - 1):
        C = C.reshape(r * ns[k], -1)
        U, s, Vt = np.linalg.svd(C, full_matrices=False)
        tail = np.sqrt(np.cumsum(s[::-1] ** 2))[::-1]
        keep = max(1, int((tail > delta).sum()))
        cores.append(U[:, :keep].reshape(r, ns[k], keep))
        C = s[:keep, None] * Vt[:keep]
        r = keep
    cores.append(C.reshape(r, ns[-1], 1))
    return cores


def main():
    import oracle
    ns_ref = {"np": np}
    # the reference's density functions and cross index function (pure numpy)
    exec(ref_lines(14, 87, ["def logf_prior", "def logL_weibull", "def crossfun"]), ns_ref)
    # the data (failure times and censoring flags) and the prior / grid construction of the reference's MAIN
    ns_ref.update({"d": 2, "n": 17})     # D = 2 covariates (the reference default is 6: 17^8 points cannot be tabulated)
    exec(ref_lines(98, 102, ["y = [6700", "censind"]), ns_ref)
    exec(ref_lines(104, 124, ["beta_mean", "theta_f = np.hstack(theta_f)"]), ns_ref)
    np.random.seed(20261019)
    exec(ref_lines(127, 128, ["np.random.randn"]), ns_ref)            # artificial covariates x
    d, n = ns_ref["d"], ns_ref["n"]
    D = d + 2
    theta, x, y, censind = ns_ref["theta"], ns_ref["x"], ns_ref["y"], ns_ref["censind"]
    beta_mean, beta_var = ns_ref["beta_mean"], ns_ref["beta_var"]
    # posterior on the full grid through the reference's crossfun (column-major multi-index, first index fastest)
    grids = np.meshgrid(*[np.arange(n)] * D, indexing="ij")
    ind = np.stack([g.ravel(order="F") for g in grids], axis=1)
    F = ns_ref["crossfun"](ind, theta, x, y, censind, beta_mean, beta_var)
    A = np.reshape(F, [n] * D, order="F")
    cores = tt_svd(A, 1e-7)
    ranks = np.array([c.shape[0] for c in cores] + [1], dtype=np.int64)
    flat = np.concatenate([c.ravel(order="F") for c in cores])
    nvec = np.full(D, n, dtype=np.int64)
    xs = np.asarray(ns_ref["theta_f"], dtype=np.float64).ravel(order="F")
    # relative accuracy of the TT on the grid
    full = cores[0]
    for c in cores[1:]:
        full = np.tensordot(full, c, axes=([-1], [0]))
    tt_err = np.linalg.norm(full.reshape([n] * D) - A) / np.linalg.norm(A)

    # seeds exactly as the reference draws them (test_shock_absorber_tt.py:146-148)
    M = 2 ** 14
    np.random.seed(7)
    q = np.random.random([M, D])
    q = np.reshape(q, [M, D], order="F")
    Z, lPz = oracle.oracle_run(nvec, xs, ranks, flat, q)
    Zr = np.reshape(Z, [M, D], order="F")
    lPz_col = np.reshape(lPz, [M, 1], order="F")
    lPex = ns_ref["logL_weibull"](Zr, x, y, censind) + ns_ref["logf_prior"](Zr, beta_mean, beta_var)   # :160
    # the reference's MH loop, executed from its source, on copies; the row index rides in an extra column of Z
    loop = {"np": np, "M": M, "Z": np.hstack([Zr, np.arange(M, dtype=np.float64)[:, None]]), "lPex": lPex.copy(), "lPz": lPz_col.copy()}
    np.random.seed(11)
    u = np.array([np.random.random(1)[0] for _ in range(M - 1)])     # the uniforms the loop will draw, in order
    np.random.seed(11)
    exec(ref_lines(164, 171, ["num_of_rejects = 0", "alpha = np.exp(lPex[i+1] - lPex[i] + lPz[i] - lPz[i+1])", "np.random.random(1)"]), loop)
    src = loop["Z"][:, D].astype(np.int32)
    Zp = loop["Z"][:, :D]
    # the quantile of interest, reference :175-180, on the pruned chain
    qoi = {"np": np, "Z": Zp, "d": d}
    exec(ref_lines(176, 179, ["q_post = theta1*((-np.log(q))**(1.0/theta2))"]), qoi)
    import hashlib
    # q and u are regenerated by the tests from the legacy np.random seeds (7 and 11); their hashes guard against drift
    out = {"n": nvec, "ranks": ranks, "xs": xs, "cores": flat, "M": M, "q_seed": 7, "u_seed": 11,
           "q_sha256": hashlib.sha256(np.ascontiguousarray(q).tobytes()).hexdigest(), "u_sha256": hashlib.sha256(u.tobytes()).hexdigest(),
           "x": np.asarray(x, dtype=np.float64), "y": np.asarray(y, dtype=np.float64), "censind": np.asarray(censind, dtype=np.int64),
           "beta_mean": beta_mean, "beta_var": beta_var, "d_cov": d,
           "Z_oracle_sha256": hashlib.sha256(np.asfortranarray(Z).tobytes(order="F")).hexdigest(), "lPz_oracle": lPz, "lPex_ref": lPex.ravel(),
           "src_ref": src, "num_of_rejects_ref": int(loop["num_of_rejects"]), "q_post_mean_ref": float(np.mean(qoi["q_post"])),
           "tt_rel_err": tt_err}
    np.savez_compressed(os.path.join(HERE, "shock_mh_D2.npz"), **out)
    print("ranks", ranks.tolist(), "TT rel err %.2e" % tt_err, "rejects", loop["num_of_rejects"], "of", M,
          "q_post mean %.6f" % out["q_post_mean_ref"], "min TT value %.3e" % full.min())


if __name__ == "__main__":
    main()
