"""TEST INFRASTRUCTURE ONLY: numpy restatement of the squared-density inverse Rosenblatt transform.

Follows /root/reference/matlab/samplers/tt_irt_sqr.m:1-208 statement by statement (the `tracemult` MEX it calls is
matlab/utils/tracemult.c:103-112 for the batched product and :131-136 for the column pick).  Nothing under
tt-irt_b200/ imports this file; only tests/, __graft_entry__.smoke() and the bench scripts' checker legs do.

PARITY PINS.  The routine is Matlab-only and the reference ships no golden vectors for it; this image has
neither Matlab nor Octave.  (a) The reference's own source files tt_irt_sqr.m and tt_rt_sqr.m are executed here, unmodified,
by oracle/mlite.py (an interpreter for the Matlab subset they use; numpy / LAPACK numerics stand in for Matlab's), with the MEX
function they call, tracemult.c, compiled unmodified against a stand-in mex.h (oracle/mexstub/): six seeded TTs -- cores with and
without boundary nodes, a marginal, signed cores, seeds 0 and 1 -- agree with this file to 1e-12 (tests/golden/make_golden_matlab.py
-> tests/golden/matlab_sqr_*.npz, tests/test_matlab_pins.py).  (b) Independently of any interpreter it is pinned against the UNMODIFIED reference C routine
tt_irt1 (oracle/_ref) wherever the two transforms are the same map (tests/test_sqr_oracle.py): on separable (rank-1)
sqrt-densities every output (Z and the log-density), and for general ranks the first coordinate, which exercises the whole
backward sweep (core x R, weighted QR, Cartesian square) against tt_irt1's marginalisation of the Kronecker-squared TT.
What the reference C cannot pin -- the squared LINEAR interpolant of the interface between grid nodes for ranks > 1 -- is
pinned by closed forms: CDF(x_k) = q_k for every sample and dimension with the conditionals rebuilt independently, the
first semi-marginal against the explicitly marginalised squared TT, marginal sampling = prefix of the full transform,
boundary extension = explicitly extended cores, the zero-mass fallback.
Third-party arithmetic: `qr` (LAPACK dgeqrf behind Matlab, tt_irt_sqr.m:69) is numpy.linalg.qr here, i.e. the
wheel-bundled OpenBLAS 0.3.30 dgeqrf; only R'R enters the result, so the factor's row signs do not matter.
"""
import numpy as np

EPS = np.finfo(np.float64).eps


def split_cores(n, ranks, cores):
    """TT2.0 flat storage -> list of (r_k, n_k, r_{k+1}) arrays (column-major, as core2cell gives them, :22-23)."""
    n = np.asarray(n, dtype=np.int64)
    ranks = np.asarray(ranks, dtype=np.int64)
    cores = np.asarray(cores, dtype=np.float64).ravel()
    out, off = [], 0
    for k in range(n.size):
        sz = int(ranks[k] * n[k] * ranks[k + 1])
        out.append(cores[off:off + sz].reshape((int(ranks[k]), int(n[k]), int(ranks[k + 1])), order="F").copy())
        off += sz
    assert off == cores.size
    return out


def sqr_sweep(n, xs, ranks, cores):
    """tt_irt_sqr.m:25-82.  Returns dict with
         n     mode sizes after boundary extension (:34-36)
         f     list of cores (r_k, n_k, r_{k+1}), extrapolated to the boundary where the TT lacks it (:53-60)
         h     list of (n_k,) interval vectors with h[0] = 0 (:47-48)
         P     list of (r_k^2, n_k): the Cartesian square of core_k x R_{k+1}, summed over the right index (:75-80)
         R     list of the R' factors used to the right of core k (R[d] = [[1]]) (:62, 66-72)
    """
    f = split_cores(n, ranks, cores)
    d = len(f)
    n = np.array([c.shape[1] for c in f], dtype=np.int64)
    rf = np.asarray(ranks, dtype=np.int64)
    xs = np.asarray(xs, dtype=np.float64).ravel()
    if xs.size == int((n + 2).sum()):
        n = n + 2                                   # :34-36 f doesn't contain boundary points
    if xs.size != int(n.sum()):
        raise ValueError("number of grid points (with or without boundaries) in xsf should be sum of mode sizes in f")
    pos = np.concatenate([[0], np.cumsum(n)])
    P = [None] * d
    R = [None] * (d + 1)
    R[d] = np.ones((1, 1))
    Rprev = R[d]
    h = [None] * d
    for k in range(d - 1, -1, -1):
        nk = int(n[k])
        x = xs[pos[k]:pos[k] + nk]
        hk = np.zeros(nk)
        hk[1:] = x[1:] - x[:-1]                     # :47-48
        h[k] = hk
        w = np.concatenate([[hk[1]], hk[1:nk - 1] + hk[2:nk], [hk[nk - 1]]])   # :50
        w = np.sqrt(w * 0.5)                        # :51
        if f[k].shape[1] == nk - 2:                 # :53-60 linear extrapolation to the boundary
            fk = np.zeros((int(rf[k]), nk, int(rf[k + 1])))
            fk[:, 1:nk - 1, :] = f[k]
            fk[:, 0, :] = fk[:, 1, :] - (fk[:, 2, :] - fk[:, 1, :]) * hk[1] / hk[2]
            fk[:, nk - 1, :] = fk[:, nk - 2, :] + (fk[:, nk - 2, :] - fk[:, nk - 3, :]) * (hk[nk - 1] + hk[nk - 2]) / hk[nk - 2]
            f[k] = fk
        Pk = f[k].reshape((int(rf[k]) * nk, int(rf[k + 1])), order="F") @ Rprev      # :62-63
        Pk = Pk.reshape((int(rf[k]), nk, -1), order="F")                             # :64
        if k > 0:
            A = Pk * w[None, :, None]                                                 # :68
            A = A.reshape((int(rf[k]), -1), order="F")                                # :69
            Rq = np.linalg.qr(A.T, mode="r")                                          # :70 economy qr
            Rprev = Rq.T.copy()                                                       # :71
            R[k] = Rprev
        # :74-80  fk(a, b, j) = sum_s Pk(a, j, s) Pk(b, j, s)
        G = np.einsum("ajs,bjs->abj", Pk, Pk)
        P[k] = G.reshape((int(rf[k]) ** 2, nk), order="F")
    return {"n": n, "f": f, "h": h, "P": P, "R": R, "pos": pos, "xs": xs, "rf": rf}


def tt_irt_sqr_oracle(n, xs, ranks, cores, q, extras=False, block=2 ** 11):
    """[xq, lFapp] = tt_irt_sqr(xsf, f, q), tt_irt_sqr.m:1.  q is (M, D), 0 < D <= d.
    Returns xq (M, D) F-order, lFapp (M,); with extras=True also idx (M, D) int32 (0-based i0), cond (M, D) the
    first-order sensitivity of x_k to O(eps) relative perturbations of the conditional (the root's own 1/p plus the
    cancellation factor of the formula at :147-149), gap (M, D) the distance of q from the nearest CDF node, and
    lsens (M, D) = |d log p / d x_k| of the interpolated conditional."""
    sw = sqr_sweep(n, xs, ranks, cores)
    nn, f, h, P, pos, xsv, rf = sw["n"], sw["f"], sw["h"], sw["P"], sw["pos"], sw["xs"], sw["rf"]
    d = len(f)
    q = np.asarray(q, dtype=np.float64)
    if q.ndim == 1:
        q = q[:, None]
    M, D = q.shape
    D = min(d, D)
    xq = np.zeros((M, D), order="F")
    lF = np.zeros(M)
    idx = np.zeros((M, D), dtype=np.int32, order="F")
    cond = np.zeros((M, D), order="F")
    gap = np.zeros((M, D), order="F")
    lsens = np.zeros((M, D), order="F")
    for start in range(0, M, block):                # :94-103 blocking
        Mb = min(block, M - start)
        fkm1 = np.ones((1, Mb))                     # :103
        for k in range(D):
            nk = int(nn[k])
            r0 = int(rf[k])
            # :107-110  square of the conditioned left interface, (r0^2, Mb), index a + r0*b
            fk = (fkm1[:, None, :] * fkm1[None, :, :]).reshape((r0 * r0, Mb), order="F")
            fk = fk.T @ P[k]                        # :112  Mb x n_k
            Ck = np.zeros_like(fk)
            Ck[:, 1:] = 0.5 * fk[:, :-1] + 0.5 * fk[:, 1:]     # :114-115 (S has 0.5 on the diagonal and superdiagonal)
            Ck[:, 0] = 0.5 * fk[:, 0]
            Ck = Ck * h[k][None, :]                 # :116
            Ck = np.cumsum(Ck, axis=1)              # :117
            Cmax = Ck[:, nk - 1].copy()             # :120
            iz = np.nonzero(Cmax <= 0)[0]           # :121
            if iz.size:                             # :123-127
                fk[iz, :] = h[k][None, :]
                Ck[iz, :] = np.cumsum(h[k])[None, :]
                Cmax[iz] = Ck[iz, nk - 1]
            Ck = Ck / Cmax[:, None]                 # :129
            fk = fk / Cmax[:, None]                 # :130
            qk = q[start:start + Mb, k]             # :134
            i0 = np.zeros(Mb, dtype=np.int64)       # :135-136 (0-based)
            i2 = np.full(Mb, nk - 1, dtype=np.int64)
            ar = np.arange(Mb)
            while np.any(i2 - i0 > 1):              # :137-143
                i1 = (i0 + i2) // 2
                C1 = Ck[ar, i1]
                left = qk > C1
                i0 = np.where(left, i1, i0)
                i2 = np.where(~left, i1, i2)
            C1 = Ck[ar, i0]                         # :146-149
            C2 = Ck[ar, i0 + 1]
            f1 = fk[ar, i0]
            f2 = fk[ar, i0 + 1]
            x1 = xsv[pos[k] + i0]                   # :157-159
            x2 = xsv[pos[k] + i0 + 1]
            h3 = x2 - x1
            with np.errstate(divide="ignore", invalid="ignore"):
                Aq = 0.5 * (f2 - f1) / h3           # :161
                Dq = f1 ** 2 + 4 * Aq * (qk - C1)   # :162
                xk = x1 + (-f1 + np.sqrt(np.abs(Dq))) / (2 * Aq)     # :163
                lin = Aq == 0                       # :164-165
                xk = np.where(lin, x1 + (qk - C1) / f1, xk)
                xk = np.where((f1 == 0) & lin, x1, xk)               # :166-170
            xk = np.where(xk > x2, x2, xk)          # :173-177
            xk = np.where(xk < x1, x1, xk)          # :178-182
            xq[start:start + Mb, k] = xk            # :184
            wA = (x2 - xk) / h3                     # :187-188
            wB = (xk - x1) / h3
            dens = f1 * wA + f2 * wB                # :193
            with np.errstate(divide="ignore", invalid="ignore"):
                lF[start:start + Mb] += np.log(dens)  # :194
            idx[start:start + Mb, k] = i0
            if extras:
                with np.errstate(divide="ignore", invalid="ignore"):
                    root = np.sqrt(np.abs(Dq))
                    canc = np.where(lin, 0.0, np.maximum(np.abs(f1), root) / (2 * np.abs(Aq)))
                    canc = np.minimum(canc, h3 / EPS)        # the clamp at :173-182 bounds the damage by the cell
                    own = 1.0 / np.maximum(np.abs(dens), 1e-300)
                    cond[start:start + Mb, k] = np.where(np.isfinite(canc), canc, 0.0) + np.minimum(own, h3 / EPS)
                    inner = Ck[:, 1:nk - 1]
                    gap[start:start + Mb, k] = np.abs(inner - qk[:, None]).min(axis=1) if nk > 2 else 1.0
                    lsens[start:start + Mb, k] = np.abs((f2 - f1) / h3) / np.maximum(np.abs(dens), 1e-300)
            if k < d - 1:                           # :197-208  on-the-fly interpolated product
                core = f[k]                         # (r0, n, r1)
                S0 = core[:, i0, :]                 # (r0, Mb, r1)
                S1 = core[:, i0 + 1, :]
                t0 = np.einsum("am,amb->bm", fkm1, S0)
                t1 = np.einsum("am,amb->bm", fkm1, S1)
                fkm1 = t0 * wA[None, :] + t1 * wB[None, :]
    if extras:
        return xq, lF, idx, cond, gap, lsens
    return xq, lF


def tt_rt_sqr_oracle(n, xs, ranks, cores, x):
    """[q, lFapp] = tt_rt_sqr(xsf, f, x): the forward (Rosenblatt) transform, /root/reference/matlab/samplers/tt_rt_sqr.m:1-178.
    Its sweep (:41-78) and conditional (:100-127) are those of tt_irt_sqr.m; the per-dimension tail differs (:129-166):
    the cell is found on the grid, the CDF is the quadratic spline evaluated at x_k, nothing is clamped.  x is (M, D)."""
    sw = sqr_sweep(n, xs, ranks, cores)
    nn, f, h, P, pos, xsv, rf = sw["n"], sw["f"], sw["h"], sw["P"], sw["pos"], sw["xs"], sw["rf"]
    d = len(f)
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 1:
        x = x[:, None]
    M, D = x.shape
    D = min(d, D)
    q = np.zeros((M, D), order="F")
    lF = np.zeros(M)
    block = 2 ** 11
    for start in range(0, M, block):
        Mb = min(block, M - start)
        fkm1 = np.ones((1, Mb))
        for k in range(D):
            nk = int(nn[k])
            r0 = int(rf[k])
            fk = (fkm1[:, None, :] * fkm1[None, :, :]).reshape((r0 * r0, Mb), order="F")      # :103-105
            fk = fk.T @ P[k]                                                                   # :107
            Ck = np.zeros_like(fk)
            Ck[:, 1:] = 0.5 * fk[:, :-1] + 0.5 * fk[:, 1:]                                     # :109-110
            Ck[:, 0] = 0.5 * fk[:, 0]
            Ck = np.cumsum(Ck * h[k][None, :], axis=1)                                         # :111-112
            Cmax = Ck[:, nk - 1].copy()                                                        # :115
            iz = np.nonzero(Cmax <= 0)[0]
            if iz.size:                                                                        # :118-122
                fk[iz, :] = h[k][None, :]
                Ck[iz, :] = np.cumsum(h[k])[None, :]
                Cmax[iz] = Ck[iz, nk - 1]
            Ck = Ck / Cmax[:, None]                                                            # :124-125
            fk = fk / Cmax[:, None]
            xk = x[start:start + Mb, k]                                                        # :129
            grid = xsv[pos[k]:pos[k] + nk]
            i0 = np.zeros(Mb, dtype=np.int64)
            i2 = np.full(Mb, nk - 1, dtype=np.int64)
            while np.any(i2 - i0 > 1):                                                         # :132-138
                i1 = (i0 + i2) // 2
                left = xk > grid[i1]
                i0 = np.where(left, i1, i0)
                i2 = np.where(~left, i1, i2)
            ar = np.arange(Mb)
            C1 = Ck[ar, i0]                                                                    # :141-143
            f1 = fk[ar, i0]
            f2 = fk[ar, i0 + 1]
            x1 = grid[i0]                                                                      # :145-147
            x2 = grid[i0 + 1]
            h3 = x2 - x1
            Aq = 0.5 * (f2 - f1) / h3                                                          # :150
            q[start:start + Mb, k] = Aq * (xk - x1) ** 2 + f1 * (xk - x1) + C1                 # :151
            wA = (x2 - xk) / h3                                                                # :156-157
            wB = (xk - x1) / h3
            with np.errstate(divide="ignore", invalid="ignore"):
                lF[start:start + Mb] += np.log(f1 * wA + f2 * wB)                              # :162-163
            if k < d - 1:                                                                      # :166-177
                core = f[k]
                t0 = np.einsum("am,amb->bm", fkm1, core[:, i0, :])
                t1 = np.einsum("am,amb->bm", fkm1, core[:, i0 + 1, :])
                fkm1 = t0 * wA[None, :] + t1 * wB[None, :]
    return q, lF


def tracemult_oracle(A, j, B=None):
    """matlab/utils/tracemult.c, real case: C(:,:,i) = A(:,:,i) * B(:,:,j(i)) (:103-112) or C(i) = A(i, j(i)) (:131-136);
    j one-based as the MEX takes it (:106-107)."""
    j0 = np.asarray(j).astype(np.int64).ravel() - 1
    A = np.asarray(A, dtype=np.float64)
    if B is None:
        return A[np.arange(j0.size), j0].copy()
    B = np.asarray(B, dtype=np.float64)
    A = A.reshape(A.shape + (1,) * (3 - A.ndim), order="F")
    B = B.reshape(B.shape + (1,) * (3 - B.ndim), order="F")
    C = np.zeros((A.shape[0], B.shape[1], j0.size), order="F")
    for i in range(j0.size):
        C[:, :, i] = A[:, :, i] @ B[:, :, j0[i]]
    return C
