"""Host-side mirror of the reference's sampler helpers that the B200 library runs on the device
(include/tt_irt1.h, csrc/ttirt_aux.cu).  Same names and argument meaning as the reference's functions:

    qmcnodes(d, l, genvec, shift)      matlab/samplers/qmcnodes.m      (the reference loads genvec from a file and draws shift)
    randref(reference, u)              matlab/samplers/randref.m       ('uniform' passes through, 'normal S' truncates at S sigmas)
    iw_prune(lFex, lFapp)              matlab/samplers/iw_prune.m
    essinv(lFex, lFapp)                matlab/samplers/essinv.m
    hellinger(lFex, lFapp)             matlab/samplers/hellinger.m
    mcmc_prune(y, lFex, lFapp, u)      matlab/samplers/mcmc_prune.m, python/test_shock_absorber_tt.py:165-171

There is no CPU fallback: every function raises when the CUDA library reports an error.
"""
import re
from ctypes import POINTER, c_double, c_int, c_longlong, c_ulonglong, c_void_p

import numpy as np

from . import tt_irt as _tt

_dp, _lp = POINTER(c_double), POINTER(c_longlong)


def _lib():
    lib = _tt.load_library()
    if not getattr(lib, "_samplers_bound", False):
        lib.ttirt_seeds_lattice_host.argtypes = [c_longlong, c_longlong, c_longlong, c_longlong, _lp, _dp, _dp, c_longlong]
        lib.ttirt_seeds_uniform_host.argtypes = [c_longlong, c_longlong, c_longlong, c_ulonglong, _dp, c_longlong]
        lib.ttirt_truncnormal_map_host.argtypes = [c_longlong, c_double, _dp, _dp]
        lib.ttirt_iw_stats_host.argtypes = [c_longlong, _dp, _dp, _dp, _dp]
        lib.ttirt_mcmc_prune_host.argtypes = [c_longlong, _dp, _dp, _dp, c_void_p, _lp, _lp, c_longlong]
        for f in (lib.ttirt_seeds_lattice_host, lib.ttirt_seeds_uniform_host, lib.ttirt_truncnormal_map_host,
                  lib.ttirt_iw_stats_host, lib.ttirt_mcmc_prune_host):
            f.restype = c_int
        lib._samplers_bound = True
    return lib


def _check(lib, rc, what):
    if rc != 0:
        _tt._raise_last(lib, what)


def qmcnodes(d, l, genvec, shift, m0=0, M=None):
    """2^l shifted rank-1 lattice nodes in d dimensions (qmcnodes.m:6-13), rows [m0, m0+M), as an (M, d) F-ordered
    array ready for tt_irt1.  genvec: integer generating vector (the second column of the reference's lattice file),
    shift: the random shift Delta (d doubles in [0, 1))."""
    lib = _lib()
    N = 1 << int(l)
    M = N - m0 if M is None else int(M)
    z = np.ascontiguousarray(np.asarray(genvec)[:d], dtype=np.int64)
    s = np.ascontiguousarray(np.asarray(shift, dtype=np.float64).ravel()[:d])
    q = np.zeros((M, d), dtype=np.float64, order="F")
    _check(lib, lib.ttirt_seeds_lattice_host(d, M, int(m0), N, z.ctypes.data_as(_lp), s.ctypes.data_as(_dp), q.ctypes.data_as(_dp), M),
           "ttirt_seeds_lattice_host")
    return q


def rand_uniform(M, d, seed, m0=0):
    """(M, d) reproducible uniforms in [0, 1) (Philox4x32-10 on the device), the stand-in for np.random.random([M, d])."""
    lib = _lib()
    q = np.zeros((M, d), dtype=np.float64, order="F")
    _check(lib, lib.ttirt_seeds_uniform_host(d, M, int(m0), int(seed) & 0xFFFFFFFFFFFFFFFF, q.ctypes.data_as(_dp), M), "ttirt_seeds_uniform_host")
    return q


def randref(reference, u):
    """randref.m with an array of numbers in [0, 1] as its second argument (:17-20): 'uniform' returns them
    unchanged, 'normal' / 'normal S' maps them to the normal truncated at S sigmas (default 4, :22-34)."""
    u = np.asarray(u, dtype=np.float64)
    if reference[:1].lower() == "u":
        return u
    m = re.findall(r"[0-9.]+", reference)
    sigma = float(m[0]) if m else 4.0
    lib = _lib()
    flat = np.ascontiguousarray(u.ravel())
    y = np.empty_like(flat)
    _check(lib, lib.ttirt_truncnormal_map_host(flat.size, sigma, flat.ctypes.data_as(_dp), y.ctypes.data_as(_dp)), "ttirt_truncnormal_map_host")
    return y.reshape(u.shape)


def _stats(lFex, lFapp, want_weights):
    lib = _lib()
    fe = np.ascontiguousarray(np.asarray(lFex, dtype=np.float64).ravel())
    fa = np.ascontiguousarray(np.asarray(lFapp, dtype=np.float64).ravel())
    if fe.size != fa.size or fe.size < 1:
        raise ValueError("lFex and lFapp must have the same, non-zero length")
    w = np.empty_like(fe) if want_weights else None
    out = np.zeros(6)
    _check(lib, lib.ttirt_iw_stats_host(fe.size, fe.ctypes.data_as(_dp), fa.ctypes.data_as(_dp),
                                        w.ctypes.data_as(_dp) if want_weights else None, out.ctypes.data_as(_dp)), "ttirt_iw_stats_host")
    return w, out


def iw_prune(lFex, lFapp):
    """[lFex_, isstd, max_ratio, err1] = iw_prune(lFex, lFapp) (iw_prune.m:16-31).  lFex: (M,) or (M, c) with the log
    exact density in column 0 and quantities of interest in the others."""
    lFex = np.asarray(lFex, dtype=np.float64)
    col0 = lFex if lFex.ndim == 1 else lFex[:, 0]
    w, out = _stats(col0, lFapp, True)
    scaled = lFex * (w if lFex.ndim == 1 else w[:, None])                       # iw_prune.m:28
    return scaled, out[0], out[1], out[2]


def essinv(lFex, lFapp):
    """tau = N / ESS (essinv.m:11-15)."""
    return _stats(lFex, lFapp, False)[1][3]


def hellinger(lFex, lFapp):
    """Hellinger distance estimate (hellinger.m:11-17)."""
    return _stats(lFex, lFapp, False)[1][4]


def mcmc_prune(y, lFex, lFapp, u, rej_hist_len=64):
    """[y, lFex, lFapp, num_of_rejects, rej_distribution] = mcmc_prune(y, lFex, lFapp) (mcmc_prune.m:17-46) with the
    M-1 uniforms the reference draws inside its loop passed in as u.  The accept / reject chain runs on the device;
    the rows are gathered here."""
    lib = _lib()
    lFex = np.asarray(lFex, dtype=np.float64)
    col0 = np.ascontiguousarray(lFex if lFex.ndim == 1 else lFex[:, 0])
    fa = np.ascontiguousarray(np.asarray(lFapp, dtype=np.float64).ravel())
    M = fa.size
    uu = np.ascontiguousarray(np.asarray(u, dtype=np.float64).ravel())
    if col0.size != M or (M > 1 and uu.size < M - 1):
        raise ValueError("need M log-densities and M-1 uniforms")
    src = np.zeros(M, dtype=np.int32)
    nrej = c_longlong(0)
    hist = np.zeros(max(1, rej_hist_len), dtype=np.int64)
    _check(lib, lib.ttirt_mcmc_prune_host(M, col0.ctypes.data_as(_dp), fa.ctypes.data_as(_dp), uu.ctypes.data_as(_dp),
                                          src.ctypes.data_as(c_void_p), nrej, hist.ctypes.data_as(_lp), int(rej_hist_len)), "ttirt_mcmc_prune_host")
    y = np.asarray(y)
    return y[src], lFex[src], np.asarray(lFapp)[src], int(nrej.value), hist, src
