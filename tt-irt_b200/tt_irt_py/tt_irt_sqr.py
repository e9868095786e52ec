"""Host-side mirror of the reference's squared-density sampler (reference matlab/samplers/tt_irt_sqr.m:1):

    xq, lFapp = tt_irt_sqr(xsf, f, q)

Same argument order and meaning as the Matlab function: `xsf` the grids of all dimensions stacked (with or without the two
boundary points per dimension that the cores lack, :33-39), `f` the TT of the SQUARE ROOT of the density (a ttpy
tt.tensor or the TTTensor container of tt_irt.py), `q` seeds in [0,1], M x D with 0 < D <= d (D < d samples the marginal
of the first D variables, :9, :105).  The work is done by the B200 library next to this file (include/tt_irt_sqr.h);
there is no CPU fallback: without the library or a CUDA device the call raises.

`SqrModel` keeps the swept model resident on the device for repeated sampling (what tt_dirt_sample.m:46,71 does per layer).
"""
from ctypes import c_int, c_double, c_longlong, c_void_p, POINTER, byref

import numpy as np

from .tt_irt import load_library, _packed_cores, _raise_last

_bound = False


def _lib():
    global _bound
    lib = load_library()
    if not _bound:
        ip, dp, lp = POINTER(c_int), POINTER(c_double), POINTER(c_longlong)
        lib.tt_irt_sqr.restype = None
        lib.tt_irt_sqr.argtypes = [c_int, ip, c_int, dp, ip, dp, c_int, c_int, dp, dp, dp]
        lib.ttirt_sqr_model_create.restype = c_void_p
        lib.ttirt_sqr_model_create.argtypes = [c_longlong, lp, c_longlong, dp, lp, dp, c_int]
        lib.ttirt_sqr_model_destroy.restype = None
        lib.ttirt_sqr_model_destroy.argtypes = [c_void_p]
        lib.ttirt_sqr_model_get_sweep.restype = c_int
        lib.ttirt_sqr_model_get_sweep.argtypes = [c_void_p, c_longlong, dp, dp]
        lib.ttirt_sqr_model_mode_size.restype = c_longlong
        lib.ttirt_sqr_model_mode_size.argtypes = [c_void_p, c_longlong]
        lib.ttirt_sqr_sample_device.restype = c_int
        lib.ttirt_sqr_sample_device.argtypes = [c_void_p, c_longlong, c_longlong, c_void_p, c_longlong, c_void_p, c_longlong,
                                                c_void_p, c_void_p, c_void_p]
        lib.ttirt_sqr_sample_host.restype = c_int
        lib.ttirt_sqr_sample_host.argtypes = [c_void_p, c_longlong, c_longlong, dp, dp, dp, c_void_p, c_longlong]
        lib.ttirt_sqr_profile_enable.restype = None
        lib.ttirt_sqr_profile_enable.argtypes = [c_void_p, c_int]
        lib.ttirt_sqr_profile_read.restype = c_int
        lib.ttirt_sqr_profile_read.argtypes = [c_void_p, dp, lp, dp]
        lib.tt_rt_sqr.restype = None
        lib.tt_rt_sqr.argtypes = [c_int, ip, c_int, dp, ip, dp, c_int, c_int, dp, dp, dp]
        lib.ttirt_sqr_forward_host.restype = c_int
        lib.ttirt_sqr_forward_host.argtypes = [c_void_p, c_longlong, c_longlong, dp, dp, dp, c_void_p, c_longlong]
        lib.ttirt_sqr_forward_device.restype = c_int
        lib.ttirt_sqr_forward_device.argtypes = [c_void_p, c_longlong, c_longlong, c_void_p, c_longlong, c_void_p, c_longlong,
                                                 c_void_p, c_void_p, c_void_p]
        lib.ttirt_dirt_inverse_host.restype = c_int
        lib.ttirt_dirt_inverse_host.argtypes = [c_longlong, POINTER(c_void_p), c_double, c_longlong, dp, dp, dp, c_longlong]
        lib.ttirt_dirt_inverse_device.restype = c_int
        lib.ttirt_dirt_inverse_device.argtypes = [c_longlong, POINTER(c_void_p), c_double, c_longlong, c_void_p, c_longlong, c_void_p,
                                                  c_longlong, c_void_p, c_void_p]
        lib.ttirt_dirt_sample_host.restype = c_int
        lib.ttirt_dirt_sample_host.argtypes = [c_longlong, POINTER(c_void_p), c_double, c_longlong, dp, dp, dp, c_longlong]
        lib.ttirt_dirt_sample_device.restype = c_int
        lib.ttirt_dirt_sample_device.argtypes = [c_longlong, POINTER(c_void_p), c_double, c_longlong, c_void_p, c_longlong, c_void_p,
                                                 c_longlong, c_void_p, c_void_p]
        lib.ttirt_tracemult_host.restype = c_int
        lib.ttirt_tracemult_host.argtypes = [c_longlong] * 5 + [dp, dp, dp, dp]
        _bound = True
    return lib


def tracemult(A, j, B=None):
    """ C(:,:,i) = A(:,:,i)*B(:,:,j(i))  or  C(i) = A(i,j(i))  -- reference matlab/utils/tracemult.c:5-8 (real case).
        A: (p, m, n) [with B] or (n, s) [without]; j: n one-based indices; B: (m, k, s).  Returns (p, k, n) or (n,)."""
    lib = _lib()
    dp = POINTER(c_double)
    jd = np.ascontiguousarray(np.asarray(j, dtype=np.float64).ravel())
    n = jd.size
    if B is None:
        A = np.asfortranarray(A, dtype=np.float64)
        if A.ndim != 2 or A.shape[0] != n:
            raise ValueError("size(j,1) differs from size(A)")
        C = np.zeros(n)
        rc = lib.ttirt_tracemult_host(0, 0, 0, n, A.shape[1], A.ctypes.data_as(dp), jd.ctypes.data_as(dp), None, C.ctypes.data_as(dp))
    else:
        A = np.asfortranarray(A, dtype=np.float64)
        B = np.asfortranarray(B, dtype=np.float64)
        A = A.reshape(A.shape + (1,) * (3 - A.ndim), order="F")
        B = B.reshape(B.shape + (1,) * (3 - B.ndim), order="F")
        p, m, na = A.shape
        if na != n:
            raise ValueError("size(j,1) differs from size(A)")
        if B.shape[0] != m:
            raise ValueError("size(A,2) differs from size(B,1)")
        k, s = B.shape[1], B.shape[2]
        C = np.zeros((p, k, n), order="F")
        rc = lib.ttirt_tracemult_host(p, m, k, n, s, A.ctypes.data_as(dp), jd.ctypes.data_as(dp), B.ctypes.data_as(dp), C.ctypes.data_as(dp))
    if rc != 0:
        _raise_last(lib, "ttirt_tracemult_host")
    return C


def tt_rt_sqr(xsf, f, x):
    """ Forward (Rosenblatt) transform through the square root of the density (reference tt_rt_sqr.m:1-15): same inputs as
        tt_irt_sqr with points x (M x D) instead of seeds; returns (q, lFapp), q the CDF values in [0,1]."""
    return _sqr_call("tt_rt_sqr", xsf, f, x)


def tt_irt_sqr(xsf, f, q):
    """ Inverse CDF (Rosenblatt) transform through the square root of the density (reference tt_irt_sqr.m:1-15)
        Inputs:
          xsf: grid points (inc. boundaries) of all dimensions stacked (np.float64, size sum(f.n) or sum(f.n + 2))
          f: TT of SQRT(PDF) on that grid, with or without the boundary points in each variable
          q: seed points from [0,1]^D (np.float64 M x D), 0 < D <= d
        Returns:
          xq: samples mapped from q by the inverse CDF (M x D, Fortran shaped)
          lFapp: log(approximate PDF) at xq (M)
    """
    return _sqr_call("tt_irt_sqr", xsf, f, q)


def _sqr_call(symbol, xsf, f, q):
    lib = _lib()
    q = np.asfortranarray(q, dtype=np.float64)
    if q.ndim == 1:
        q = np.asfortranarray(q[:, None])
    xsf = np.ascontiguousarray(np.concatenate([np.asarray(x, dtype=np.float64).ravel() for x in xsf])
                               if isinstance(xsf, (list, tuple)) else np.asarray(xsf, dtype=np.float64).ravel(order="F"))
    core = _packed_cores(f)
    n = np.ascontiguousarray(f.n, dtype=np.int32)
    rf = np.ascontiguousarray(f.r, dtype=np.int32)
    M, D = q.shape
    if D < 1 or D > f.d:
        raise ValueError("q must have between 1 and d columns")
    if xsf.size not in (int(n.sum()), int((n + 2).sum())):
        raise ValueError("number of grid points (with or without boundaries) in xsf should be sum of mode sizes in f")
    xq = np.zeros((M, D), dtype=np.float64, order="F")
    lFapp = np.zeros(M, dtype=np.float64)
    dp, ip = POINTER(c_double), POINTER(c_int)
    getattr(lib, symbol)(c_int(f.d), n.ctypes.data_as(ip), c_int(xsf.size), xsf.ctypes.data_as(dp), rf.ctypes.data_as(ip),
                         core.ctypes.data_as(dp), c_int(M), c_int(D), q.ctypes.data_as(dp), xq.ctypes.data_as(dp),
                         lFapp.ctypes.data_as(dp))
    from .tt_irt import _check_void_call
    _check_void_call(lib, symbol, M, lFapp)
    return xq, lFapp


class SqrModel(object):
    """TT of sqrt(density) swept once on one B200 (tt_irt_sqr.m:41-82), sampled many times (:85-208)."""

    def __init__(self, n, xs, ranks, cores, device=0):
        lib = _lib()
        self._lib = lib
        self.n = np.ascontiguousarray(n, dtype=np.int64)
        self.r = np.ascontiguousarray(ranks, dtype=np.int64)
        self.d = int(self.n.size)
        xs = np.ascontiguousarray(np.asarray(xs, dtype=np.float64).ravel(order="F"))
        cores = np.ascontiguousarray(np.asarray(cores, dtype=np.float64).ravel(order="F"))
        if self.r.size != self.d + 1 or cores.size != int((self.r[:-1] * self.n * self.r[1:]).sum()):
            raise ValueError("inconsistent TT description")
        lp, dp = POINTER(c_longlong), POINTER(c_double)
        self._h = lib.ttirt_sqr_model_create(self.d, self.n.ctypes.data_as(lp), int(xs.size), xs.ctypes.data_as(dp),
                                             self.r.ctypes.data_as(lp), cores.ctypes.data_as(dp), int(device))
        if not self._h:
            _raise_last(lib, "ttirt_sqr_model_create")
        self.n_ext = np.array([lib.ttirt_sqr_model_mode_size(self._h, k) for k in range(self.d)], dtype=np.int64)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.ttirt_sqr_model_destroy(self._h)
            self._h = None

    __del__ = close

    def sweep(self, k):
        """(P{k} of tt_irt_sqr.m:80 as (r_k^2, n_k) F-order, R'R of the factor left of core k or None for k = 0)."""
        r0, nk = int(self.r[k]), int(self.n_ext[k])
        G = np.zeros((r0 * r0, nk), order="F")
        RR = np.zeros((r0, r0), order="F") if k > 0 else None
        dp = POINTER(c_double)
        if self._lib.ttirt_sqr_model_get_sweep(self._h, k, G.ctypes.data_as(dp), RR.ctypes.data_as(dp) if k > 0 else None) != 0:
            _raise_last(self._lib, "ttirt_sqr_model_get_sweep")
        return G, RR

    def sample(self, q, want_idx=False):
        q = np.asfortranarray(q, dtype=np.float64)
        if q.ndim == 1:
            q = np.asfortranarray(q[:, None])
        M, D = q.shape
        Z = np.zeros((M, D), dtype=np.float64, order="F")
        lF = np.zeros(M, dtype=np.float64)
        idx = np.zeros((M, D), dtype=np.int32, order="F") if want_idx else None
        dp = POINTER(c_double)
        rc = self._lib.ttirt_sqr_sample_host(self._h, M, D, q.ctypes.data_as(dp), Z.ctypes.data_as(dp), lF.ctypes.data_as(dp),
                                             idx.ctypes.data_as(c_void_p) if want_idx else None, M)
        if rc != 0:
            _raise_last(self._lib, "ttirt_sqr_sample_host")
        return (Z, lF, idx) if want_idx else (Z, lF)

    def forward(self, x, want_idx=False):
        """tt_rt_sqr on the resident model: points x (M x D) -> (q, lFapp[, idx])."""
        x = np.asfortranarray(x, dtype=np.float64)
        if x.ndim == 1:
            x = np.asfortranarray(x[:, None])
        M, D = x.shape
        Q = np.zeros((M, D), dtype=np.float64, order="F")
        lF = np.zeros(M, dtype=np.float64)
        idx = np.zeros((M, D), dtype=np.int32, order="F") if want_idx else None
        dp = POINTER(c_double)
        rc = self._lib.ttirt_sqr_forward_host(self._h, M, D, x.ctypes.data_as(dp), Q.ctypes.data_as(dp), lF.ctypes.data_as(dp),
                                              idx.ctypes.data_as(c_void_p) if want_idx else None, M)
        if rc != 0:
            _raise_last(self._lib, "ttirt_sqr_forward_host")
        return (Q, lF, idx) if want_idx else (Q, lF)

    def sample_device(self, M, D, q_ptr, ldq, z_ptr, ldz, lf_ptr, idx_ptr=None, stream=None):
        rc = self._lib.ttirt_sqr_sample_device(self._h, int(M), int(D), c_void_p(q_ptr), int(ldq), c_void_p(z_ptr), int(ldz),
                                               c_void_p(lf_ptr), c_void_p(idx_ptr) if idx_ptr else None,
                                               c_void_p(stream) if stream else None)
        if rc != 0:
            _raise_last(self._lib, "ttirt_sqr_sample_device")

    def profile_enable(self, on=True):
        self._lib.ttirt_sqr_profile_enable(self._h, 1 if on else 0)

    def profile_read(self):
        """(summed pdf-kernel ms, launches timed, algorithmic flops of those launches)."""
        ms, n, fl = c_double(0), c_longlong(0), c_double(0)
        if self._lib.ttirt_sqr_profile_read(self._h, byref(ms), byref(n), byref(fl)) != 0:
            _raise_last(self._lib, "ttirt_sqr_profile_read")
        return ms.value, n.value, fl.value


def flops_per_sample(ns, ranks, D=None):
    """Algorithmic FP64 flops per sample of the squared-density transform: per sampled dimension the symmetric half of the
    conditional contraction (:107-112), r_k (r_k + 1) n_k + r_k (r_k + 1) / 2, plus 4 r_k r_{k+1} per interface update (:205)."""
    ns = np.asarray(ns, dtype=np.int64)
    ranks = np.asarray(ranks, dtype=np.int64)
    D = ns.size if D is None else int(D)
    r0 = ranks[:D]
    w = (r0 * (r0 + 1) * ns[:D]).sum() + (r0 * (r0 + 1) // 2).sum()
    w += (4 * ranks[:D - 1] * ranks[1:D]).sum()
    return int(w)


def _reference_sigma(reference):
    """tt_dirt_sample.m:21-30: 0 for a uniform reference, else the half-width of the truncated normal ('Normal' -> 4)."""
    if str(reference)[0].lower() == "u":
        return 0.0
    digits = "".join(ch for ch in str(reference) if ch == "." or ch.isdigit())
    try:
        return float(digits)
    except ValueError:
        return 4.0


class Dirt(object):
    """A Deep Inverse Rosenblatt Transform resident on one B200: IRTstruct.F0 on IRTstruct.x0 and IRTstruct.F{1..nlvl} on
    IRTstruct.x (reference matlab/samplers/tt_dirt_sample.m, fields of tt_dirt_approx.m:158-162, 257, 321), each swept once.
    levels[0] = (n, x0, ranks, cores) of F0, levels[j] = (n, x, ranks, cores) of F{j}; spline interpolation only."""

    def __init__(self, levels, reference="uni", device=0):
        self.sigma = _reference_sigma(reference)
        self.models = [SqrModel(n, xs, rk, c, device=device) for (n, xs, rk, c) in levels]
        self.d = self.models[0].d
        self._lib = self.models[0]._lib
        self._arr = (c_void_p * len(self.models))(*[m._h for m in self.models])

    def close(self):
        for m in getattr(self, "models", []):
            m.close()
        self.models = []

    def sample(self, q):
        """[z, lFapp] = tt_dirt_sample(IRTstruct, q): q on [0,1]^d (uniform reference) or [-S,S]^d (truncated normal)."""
        q = np.asfortranarray(q, dtype=np.float64)
        M, d = q.shape
        if d != self.d:
            raise ValueError("q must be M x d")
        z = np.zeros((M, d), order="F")
        lF = np.zeros(M)
        dp = POINTER(c_double)
        rc = self._lib.ttirt_dirt_sample_host(len(self.models), self._arr, self.sigma, M, q.ctypes.data_as(dp), z.ctypes.data_as(dp),
                                              lF.ctypes.data_as(dp), M)
        if rc != 0:
            _raise_last(self._lib, "ttirt_dirt_sample_host")
        return z, lF

    def inverse(self, x):
        """[q, lFapp] = tt_dirt_inverse(IRTstruct, x): points of the target space back to the reference space."""
        x = np.asfortranarray(x, dtype=np.float64)
        M, d = x.shape
        if d != self.d:
            raise ValueError("x must be M x d")
        q = np.zeros((M, d), order="F")
        lF = np.zeros(M)
        dp = POINTER(c_double)
        rc = self._lib.ttirt_dirt_inverse_host(len(self.models), self._arr, self.sigma, M, x.ctypes.data_as(dp), q.ctypes.data_as(dp),
                                               lF.ctypes.data_as(dp), M)
        if rc != 0:
            _raise_last(self._lib, "ttirt_dirt_inverse_host")
        return q, lF

    def sample_device(self, M, q_ptr, ldq, z_ptr, ldz, lf_ptr, stream=None):
        rc = self._lib.ttirt_dirt_sample_device(len(self.models), self._arr, self.sigma, int(M), c_void_p(q_ptr), int(ldq), c_void_p(z_ptr),
                                                int(ldz), c_void_p(lf_ptr), c_void_p(stream) if stream else None)
        if rc != 0:
            _raise_last(self._lib, "ttirt_dirt_sample_device")


def tt_dirt_inverse(IRTstruct, x):
    """[q, lFapp] = tt_dirt_inverse(IRTstruct, x)  (reference tt_dirt_inverse.m:1); IRTstruct as for tt_dirt_sample."""
    drt = _dirt_from_struct(IRTstruct)
    try:
        return drt.inverse(x)
    finally:
        drt.close()


def tt_dirt_sample(IRTstruct, q):
    """[z, lFapp] = tt_dirt_sample(IRTstruct, q)  (reference tt_dirt_sample.m:1; the exact-density evaluation of :76-82 stays
    with the caller).  IRTstruct: mapping with the fields tt_dirt_approx leaves behind -- 'x0', 'F0', 'x', 'F' (list),
    'reference', and optionally 'interpolation' / 'crossmethod'; F0 / F{j} are ttpy tensors or TTTensor containers."""
    drt = _dirt_from_struct(IRTstruct)
    try:
        return drt.sample(q)
    finally:
        drt.close()


def _dirt_from_struct(IRTstruct):
    if str(IRTstruct.get("crossmethod", "amen_cross_s")) == "build_ftt" or str(IRTstruct.get("interpolation", "spline"))[0] != "s":
        raise NotImplementedError("only the spline / TT-cross branch of tt_dirt_sample (tt_irt_sqr, :46, :71) is built")

    def level(f, x):
        x = np.concatenate([np.asarray(v, dtype=np.float64).ravel() for v in x]) if isinstance(x, (list, tuple)) else np.asarray(x, dtype=np.float64).ravel()
        return (np.asarray(f.n, dtype=np.int64), x, np.asarray(f.r, dtype=np.int64), _packed_cores(f))
    levels = [level(IRTstruct["F0"], IRTstruct["x0"])] + [level(f, IRTstruct["x"]) for f in IRTstruct["F"]]
    return Dirt(levels, IRTstruct.get("reference", "uni"))
