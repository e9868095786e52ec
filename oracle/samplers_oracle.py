"""TEST INFRASTRUCTURE ONLY: CPU restatements (numpy) of the steps either side of tt_irt1 that the B200 library also
offers on the device.  Each function cites the reference lines it follows.

Parity pins: the reference ships no golden vectors for these helpers and this image has neither Matlab nor Octave.  They are
pinned (a) against the reference's own source files executed here -- qmcnodes.m, randref.m, essinv.m, hellinger.m, iw_prune.m
and mcmc_prune.m run unmodified under oracle/mlite.py, a small interpreter for the Matlab subset they use (numpy numerics
stand in for Matlab's; tests/golden/make_golden_matlab.py -> tests/golden/matlab_helpers.npz, tests/test_matlab_pins.py),
mcmc_prune also against the reference's Python loop (tests/golden/make_golden_mh.py) -- and (b) by closed forms
(tests/test_samplers_oracle.py).  The Philox4x32-10 generator, which the reference does not have, is pinned by the published
known-answer vectors of the Random123 distribution.
"""
import numpy as np
from math import erf, sqrt

try:  # scipy is present in this image; erfinv has no numpy equivalent
    from scipy.special import erfinv as _erfinv
except Exception:  # pragma: no cover
    _erfinv = None


def qmc_lattice(d, l, genvec, shift, m0=0, M=None):
    """matlab/samplers/qmcnodes.m:6-13.  Y = (0:2^l-1)/2^l; Y = z(1:d)*Y; Y = Y + Delta; Y = Y - floor(Y).
    Returns the slice [m0, m0+M) as an (M, d) F-ordered array (the reference returns d x N; tt_irt1 takes M x d)."""
    N = 2 ** l
    M = N - m0 if M is None else M
    y = np.arange(m0, m0 + M, dtype=np.float64) / float(N)                     # :6-7
    z = np.asarray(genvec[:d], dtype=np.float64)
    Y = z[:, None] * y[None, :]                                                # :8
    Y = Y + np.asarray(shift, dtype=np.float64)[:d, None]                      # :10-12
    Y = Y - np.floor(Y)                                                        # :13
    return np.asfortranarray(Y.T)


def philox4x32_10(counter, key):
    """Philox4x32-10 (Salmon et al. 2011).  counter: (..., 4) uint32, key: (2,) uint32 -> (..., 4) uint32."""
    c = np.array(counter, dtype=np.uint64).reshape(-1, 4).copy()
    k0, k1 = np.uint64(int(key[0]) & 0xFFFFFFFF), np.uint64(int(key[1]) & 0xFFFFFFFF)
    M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = M0 * c[:, 0]
        p1 = M1 * c[:, 2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & mask
        hi1, lo1 = p1 >> np.uint64(32), p1 & mask
        n0 = hi1 ^ c[:, 1] ^ k0
        n2 = hi0 ^ c[:, 3] ^ k1
        c = np.stack([n0, lo1, n2, lo0], axis=1)
        k0 = (k0 + np.uint64(0x9E3779B9)) & mask
        k1 = (k1 + np.uint64(0xBB67AE85)) & mask
    return c.astype(np.uint32).reshape(np.shape(counter))


def uniform_philox(d, M, seed, m0=0):
    """The library's reproducible stand-in for rand / np.random.random (python/test_shock_absorber_tt.py:147):
    counter = (index lo, index hi, dimension, 0), key = (seed lo, seed hi); u = (first 64 bits >> 11) * 2^-53."""
    idx = np.arange(m0, m0 + M, dtype=np.uint64)
    out = np.empty((M, d), dtype=np.float64, order="F")
    key = (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    for k in range(d):
        ctr = np.stack([idx & np.uint64(0xFFFFFFFF), idx >> np.uint64(32), np.full(M, k, dtype=np.uint64),
                        np.zeros(M, dtype=np.uint64)], axis=1)
        r = philox4x32_10(ctr, key).astype(np.uint64)
        bits = (r[:, 1] << np.uint64(32)) | r[:, 0]
        out[:, k] = (bits >> np.uint64(11)).astype(np.float64) * 2.0 ** -53
    return out


def truncnormal_map(u, sigma=4.0):
    """matlab/samplers/randref.m:31-33.  cdf_ifactor = erf(sigma/sqrt(2))/0.5; y = erfinv((u-0.5)*cdf_ifactor)*sqrt(2)."""
    cdf_ifactor = erf(sigma / sqrt(2.0)) / 0.5
    return _erfinv((np.asarray(u, dtype=np.float64) - 0.5) * cdf_ifactor) * sqrt(2.0)


def iw_prune(lFex, lFapp):
    """matlab/samplers/iw_prune.m:19-29 for a single exact-density column.  Returns (weights, isstd, max_ratio, err1,
    log_renorm); the reference's output lFex_ is lFex .* weights."""
    lFex = np.asarray(lFex, dtype=np.float64); lFapp = np.asarray(lFapp, dtype=np.float64)
    w = np.exp(lFex - lFapp)                                                   # :19
    renorm = w.mean()                                                          # :20
    w = w / renorm                                                             # :21
    max_ratio = w.max()                                                        # :24
    lren = np.log(renorm)                                                      # :25
    err1 = np.mean(np.abs(np.exp(lFex - lren) - np.exp(lFapp)) / np.exp(lFapp))  # :26
    isstd = np.sqrt(np.mean((w - 1.0) ** 2))                                   # :29
    return w, isstd, max_ratio, err1, lren


def essinv(lFex, lFapp):
    """matlab/samplers/essinv.m:12-14."""
    dF = np.asarray(lFex, dtype=np.float64) - np.asarray(lFapp, dtype=np.float64)
    dF = dF - dF.max()
    return dF.size * np.sum(np.exp(dF * 2)) / np.sum(np.exp(dF)) ** 2


def hellinger(lFex, lFapp):
    """matlab/samplers/hellinger.m:12-16."""
    dF = np.asarray(lFex, dtype=np.float64) - np.asarray(lFapp, dtype=np.float64)
    dF = dF - dF.max()
    lZex = np.log(np.mean(np.exp(dF)))
    H = np.mean((np.exp(0.5 * (dF - lZex)) - 1.0) ** 2)
    return np.sqrt(H / 2)


def mcmc_prune(lFex, lFapp, u):
    """matlab/samplers/mcmc_prune.m:24-43 with the uniforms pre-drawn (u[i] is the i-th call of rand).
    Returns (src, num_of_rejects, rej_distribution): src[i] is the index of the sample that occupies position i
    after the in-place copies of :29-31 (so y_pruned = y[src], lFex_pruned = lFex[src], ...)."""
    lFex = np.asarray(lFex, dtype=np.float64); lFapp = np.asarray(lFapp, dtype=np.float64)
    M = lFapp.size
    src = np.arange(M, dtype=np.int32)
    fe = lFex.copy(); fa = lFapp.copy()
    rej_distribution = []
    num_of_rejects = 0
    rej_seq = 0
    for i in range(M - 1):
        alpha = ((fe[i + 1] - fe[i]) - fa[i + 1]) + fa[i]                       # :25 (left-to-right)
        alpha = np.exp(alpha)                                                  # :26
        if alpha < u[i]:                                                       # :27
            src[i + 1] = src[i]; fa[i + 1] = fa[i]; fe[i + 1] = fe[i]          # :29-31
            num_of_rejects += 1; rej_seq += 1                                  # :32-33
        elif rej_seq > 0:                                                      # :34
            while len(rej_distribution) < rej_seq:
                rej_distribution.append(0)
            rej_distribution[rej_seq - 1] += 1                                 # :36-40
            rej_seq = 0                                                        # :41
    return src, num_of_rejects, np.array(rej_distribution, dtype=np.int64)
