"""Host-side mirror of the reference's ctypes wrapper (reference python/tt_irt_py/tt_irt.py:1-53).

Same entry point, argument meaning and return values:

    Z, lPz = tt_irt1(q, f, xsf)

but the library found next to this file (glob "tt_irt1*", as reference tt_irt.py:8-11 does) is the
B200-native one built by tt-irt_b200/Makefile.  `f` may be a ttpy `tt.tensor` (attributes d, n, r,
core, ps -- exactly what the reference reads) or the minimal `TTTensor` container below; ttpy itself
is not required.  There is no CPU fallback: if the library or a CUDA device is missing the call raises.

Additions beyond the reference wrapper (they do not change tt_irt1): `Model`, a cached device-resident
TT density for repeated sampling, used by bench.py and the tests.
"""
from ctypes import cdll, c_int, c_double, c_longlong, c_void_p, c_char_p, POINTER
import glob
import os

import numpy as np

MODE_FAST, MODE_STRICT = 0, 1

_lib = None


def load_library():
    """cdll-load the first file matching tt_irt1* next to this module (reference tt_irt.py:8-11)."""
    global _lib
    if _lib is None:
        pat = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tt_irt1*")
        hits = sorted(p for p in glob.glob(pat) if p.endswith(".so"))
        if not hits:
            raise RuntimeError("tt_irt1*.so not found next to %s: run `make -C tt-irt_b200` "
                               "(this package has no CPU fallback)" % __file__)
        # TTIRT_LIBRARY: an alternative build of the same library (kernel experiments); default is the in-tree one
        lib = cdll.LoadLibrary(os.environ.get("TTIRT_LIBRARY", hits[0]))
        ip, dp, lp = POINTER(c_int), POINTER(c_double), POINTER(c_longlong)
        lib.tt_irt1.restype = None
        lib.tt_irt1.argtypes = [c_int, ip, dp, ip, dp, c_int, dp, dp, dp]
        lib.ttirt_model_create.restype = c_void_p
        lib.ttirt_model_create.argtypes = [c_longlong, lp, dp, lp, dp, c_int]
        lib.ttirt_model_destroy.restype = None
        lib.ttirt_model_destroy.argtypes = [c_void_p]
        lib.ttirt_model_get_sweep.restype = c_int
        lib.ttirt_model_get_sweep.argtypes = [c_void_p, dp, dp]
        lib.ttirt_sample_device.restype = c_int
        lib.ttirt_sample_device.argtypes = [c_void_p, c_longlong, c_void_p, c_longlong, c_void_p, c_longlong,
                                            c_void_p, c_void_p, c_int, c_void_p]
        lib.ttirt_sample_host.restype = c_int
        lib.ttirt_sample_host.argtypes = [c_void_p, c_longlong, dp, dp, dp, c_void_p, c_longlong, c_int]
        lib.ttirt_run_host.restype = c_int
        lib.ttirt_run_host.argtypes = [c_longlong, lp, dp, lp, dp, c_longlong, dp, dp, dp, c_void_p, c_int, c_int, c_int]
        lib.ttirt_kernel_launches.restype = c_longlong
        lib.ttirt_profile_enable.restype = None
        lib.ttirt_profile_enable.argtypes = [c_void_p, c_int]
        lib.ttirt_profile_read.restype = c_int
        lib.ttirt_profile_read.argtypes = [c_void_p, dp, lp, dp]
        lib.ttirt_last_error.restype = c_char_p
        lib.ttirt_device_count.restype = c_int
        lib.ttirt_set_chunk.restype = None
        lib.ttirt_set_chunk.argtypes = [c_longlong]
        _lib = lib
    return _lib


class TTTensor(object):
    """Minimal stand-in for ttpy's tt.tensor: d, n, r, core (TT2.0 contiguous storage), ps (1-based)."""

    def __init__(self, n, r, core):
        self.n = np.asarray(n, dtype=np.int32)
        self.r = np.asarray(r, dtype=np.int32)
        self.d = int(self.n.size)
        self.core = np.ascontiguousarray(core, dtype=np.float64).ravel()
        sizes = self.r[:-1].astype(np.int64) * self.n * self.r[1:]
        self.ps = np.concatenate([[1], 1 + np.cumsum(sizes)]).astype(np.int64)
        if self.core.size != int(sizes.sum()):
            raise ValueError("core has %d entries, ranks/modes need %d" % (self.core.size, int(sizes.sum())))


def _packed_cores(f):
    """Cores repacked contiguously by position vector ps (reference tt_irt.py:27-34)."""
    core = np.zeros(int(np.sum(np.asarray(f.r[:-1], dtype=np.int64) * np.asarray(f.n, dtype=np.int64)
                               * np.asarray(f.r[1:], dtype=np.int64))), dtype=np.float64)
    ps_my = 0
    for i in range(0, f.d):
        sz = int(f.r[i]) * int(f.n[i]) * int(f.r[i + 1])
        core[ps_my:ps_my + sz] = np.asarray(f.core)[int(f.ps[i]) - 1:int(f.ps[i]) - 1 + sz]
        ps_my += sz
    return core


def _raise_last(lib, what):
    msg = lib.ttirt_last_error()
    raise RuntimeError("%s failed: %s" % (what, msg.decode() if msg else "unknown error"))


def _check_void_call(lib, what, M, lPz):
    """The reference's entry points return nothing; the B200 library reports failure (no device, unsupported shape, CUDA
    error) by NaN-filling its outputs and keeping a message.  The Python mirror turns that into an exception."""
    if M > 0 and np.isnan(lPz[0]) and np.isnan(lPz).all():
        msg = lib.ttirt_last_error()
        raise RuntimeError("%s failed: %s" % (what, msg.decode() if msg else "outputs are NaN (see stderr)"))


def tt_irt1(q, f, xsf):
    """ Inverse Rosenblatt sampler, linear splines (reference tt_irt.py:13-53)
        Inputs:
          q: seed samples from [0,1]^d (np.float64 M x d Fortran shaped)
          f: tt.tensor of the PDF (dimension d), constructed on a grid specified in xsf
          xsf: vector of grid points of all variables stacked together (np.float64, size sum(f.n))
        Returns:
          Z: transformed samples (np.float64 M x d Fortran shaped)
          lPz: values of log(sampling density) at Z (np.float64 M x 1)
    """
    lib = load_library()
    q = np.asfortranarray(q, dtype=np.float64)
    xsf = np.ascontiguousarray(np.asarray(xsf, dtype=np.float64).ravel(order="F"))
    core = _packed_cores(f)
    n = np.ascontiguousarray(f.n, dtype=np.int32)
    rf = np.ascontiguousarray(f.r, dtype=np.int32)
    if q.ndim != 2 or q.shape[1] != f.d:
        raise ValueError("q must be M x d")
    if xsf.size != int(n.sum()):
        raise ValueError("xsf must stack all grids (size sum(f.n))")
    Z = np.zeros([q.shape[0], q.shape[1]], dtype=np.float64, order='F')
    lPz = np.zeros([q.shape[0]], dtype=np.float64, order='F')
    dp, ip = POINTER(c_double), POINTER(c_int)
    # Sampler is actually here
    lib.tt_irt1(c_int(f.d), n.ctypes.data_as(ip), xsf.ctypes.data_as(dp), rf.ctypes.data_as(ip),
                core.ctypes.data_as(dp), c_int(q.shape[0]), q.ctypes.data_as(dp), Z.ctypes.data_as(dp),
                lPz.ctypes.data_as(dp))
    _check_void_call(lib, "tt_irt1", q.shape[0], lPz)
    return (Z, lPz)


class Model(object):
    """A TT density resident on one B200: cores uploaded and marginalised once, sampled many times."""

    def __init__(self, n, xs, ranks, cores, device=0):
        lib = load_library()
        self._lib = lib
        self.n = np.ascontiguousarray(n, dtype=np.int64)
        self.r = np.ascontiguousarray(ranks, dtype=np.int64)
        self.d = int(self.n.size)
        xs = np.ascontiguousarray(np.asarray(xs, dtype=np.float64).ravel(order="F"))
        cores = np.ascontiguousarray(np.asarray(cores, dtype=np.float64).ravel(order="F"))
        if self.r.size != self.d + 1 or xs.size != int(self.n.sum()) or \
                cores.size != int((self.r[:-1] * self.n * self.r[1:]).sum()):
            raise ValueError("inconsistent TT description")
        lp, dp = POINTER(c_longlong), POINTER(c_double)
        self._h = lib.ttirt_model_create(self.d, self.n.ctypes.data_as(lp), xs.ctypes.data_as(dp),
                                         self.r.ctypes.data_as(lp), cores.ctypes.data_as(dp), int(device))
        if not self._h:
            _raise_last(lib, "ttirt_model_create")
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.ttirt_model_destroy(self._h)
            self._h = None

    __del__ = close

    def sweep(self):
        """(P_k list (r_k, n_k) F-order, marginals list) computed on the device."""
        pk = np.zeros(int((self.r[:-1] * self.n).sum()))
        mg = np.zeros(int(self.r[1:].sum()))
        dp = POINTER(c_double)
        if self._lib.ttirt_model_get_sweep(self._h, pk.ctypes.data_as(dp), mg.ctypes.data_as(dp)) != 0:
            _raise_last(self._lib, "ttirt_model_get_sweep")
        P, Mg, op, om = [], [], 0, 0
        for k in range(self.d):
            sz = int(self.r[k] * self.n[k])
            P.append(pk[op:op + sz].reshape((int(self.r[k]), int(self.n[k])), order="F"))
            op += sz
            Mg.append(mg[om:om + int(self.r[k + 1])])
            om += int(self.r[k + 1])
        return P, Mg

    def sample(self, q, mode=MODE_FAST, want_idx=False):
        """Host buffers in, host buffers out (chunked H2D -> kernels -> D2H on this model's device)."""
        q = np.asfortranarray(q, dtype=np.float64)
        M = q.shape[0]
        if q.ndim != 2 or q.shape[1] != self.d:
            raise ValueError("q must be M x d")
        Z = np.zeros((M, self.d), dtype=np.float64, order="F")
        lPz = np.zeros(M, dtype=np.float64)
        idx = np.zeros((M, self.d), dtype=np.int32, order="F") if want_idx else None
        dp = POINTER(c_double)
        rc = self._lib.ttirt_sample_host(self._h, M, q.ctypes.data_as(dp), Z.ctypes.data_as(dp), lPz.ctypes.data_as(dp),
                                         idx.ctypes.data_as(c_void_p) if want_idx else None, M, int(mode))
        if rc != 0:
            _raise_last(self._lib, "ttirt_sample_host")
        return (Z, lPz, idx) if want_idx else (Z, lPz)

    def sample_lattice(self, l, genvec, shift, m0=0, M=None, mode=MODE_FAST, want_q=False):
        """tt_irt1 on the 2^l-point shifted rank-1 lattice of qmcnodes.m, generated on the device (no q upload).
        Returns (Z, lPz) or (Z, lPz, q)."""
        N = 1 << int(l)
        M = N - m0 if M is None else int(M)
        z = np.ascontiguousarray(np.asarray(genvec)[:self.d], dtype=np.int64)
        sh = np.ascontiguousarray(np.asarray(shift, dtype=np.float64).ravel()[:self.d])
        Z = np.zeros((M, self.d), dtype=np.float64, order="F")
        lPz = np.zeros(M, dtype=np.float64)
        q = np.zeros((M, self.d), dtype=np.float64, order="F") if want_q else None
        dp, lp = POINTER(c_double), POINTER(c_longlong)
        f = self._lib.ttirt_sample_lattice_host
        f.argtypes = [c_void_p, c_longlong, c_longlong, c_longlong, lp, dp, dp, dp, dp, c_longlong, c_int]
        f.restype = c_int
        rc = f(self._h, M, int(m0), N, z.ctypes.data_as(lp), sh.ctypes.data_as(dp), q.ctypes.data_as(dp) if want_q else None,
               Z.ctypes.data_as(dp), lPz.ctypes.data_as(dp), M, int(mode))
        if rc != 0:
            _raise_last(self._lib, "ttirt_sample_lattice_host")
        return (Z, lPz, q) if want_q else (Z, lPz)

    def sample_uniform(self, M, seed, m0=0, mode=MODE_FAST, want_q=False):
        """tt_irt1 on M reproducible uniform seed points (Philox4x32-10) generated on the device (no q upload)."""
        from ctypes import c_ulonglong
        Z = np.zeros((M, self.d), dtype=np.float64, order="F")
        lPz = np.zeros(M, dtype=np.float64)
        q = np.zeros((M, self.d), dtype=np.float64, order="F") if want_q else None
        dp = POINTER(c_double)
        f = self._lib.ttirt_sample_uniform_host
        f.argtypes = [c_void_p, c_longlong, c_longlong, c_ulonglong, dp, dp, dp, c_longlong, c_int]
        f.restype = c_int
        rc = f(self._h, int(M), int(m0), int(seed) & 0xFFFFFFFFFFFFFFFF, q.ctypes.data_as(dp) if want_q else None,
               Z.ctypes.data_as(dp), lPz.ctypes.data_as(dp), int(M), int(mode))
        if rc != 0:
            _raise_last(self._lib, "ttirt_sample_uniform_host")
        return (Z, lPz, q) if want_q else (Z, lPz)

    def profile_enable(self, on=True):
        self._lib.ttirt_profile_enable(self._h, 1 if on else 0)

    def profile_read(self):
        """(summed transition-kernel ms, launches timed, algorithmic flops of those launches)."""
        ms, n, fl = c_double(0), c_longlong(0), c_double(0)
        from ctypes import byref
        if self._lib.ttirt_profile_read(self._h, byref(ms), byref(n), byref(fl)) != 0:
            _raise_last(self._lib, "ttirt_profile_read")
        return ms.value, n.value, fl.value

    def sample_device(self, M, q_ptr, ldq, z_ptr, ldz, lpz_ptr, idx_ptr=None, mode=MODE_FAST, stream=None):
        """Raw device pointers (ints); enqueues on `stream` (cudaStream_t as int) without synchronising."""
        rc = self._lib.ttirt_sample_device(self._h, int(M), c_void_p(q_ptr), int(ldq), c_void_p(z_ptr), int(ldz),
                                           c_void_p(lpz_ptr), c_void_p(idx_ptr) if idx_ptr else None, int(mode),
                                           c_void_p(stream) if stream else None)
        if rc != 0:
            _raise_last(self._lib, "ttirt_sample_device")


def run_host(n, xs, ranks, cores, q, mode=MODE_FAST, first_device=0, n_devices=1, want_idx=False):
    """One-shot call on host buffers, rows sharded over n_devices GPUs (what the C tt_irt1 runs)."""
    lib = load_library()
    n = np.ascontiguousarray(n, dtype=np.int64)
    r = np.ascontiguousarray(ranks, dtype=np.int64)
    xs = np.ascontiguousarray(np.asarray(xs, dtype=np.float64).ravel(order="F"))
    cores = np.ascontiguousarray(np.asarray(cores, dtype=np.float64).ravel(order="F"))
    q = np.asfortranarray(q, dtype=np.float64)
    M, d = q.shape
    Z = np.zeros((M, d), dtype=np.float64, order="F")
    lPz = np.zeros(M, dtype=np.float64)
    idx = np.zeros((M, d), dtype=np.int32, order="F") if want_idx else None
    lp, dp = POINTER(c_longlong), POINTER(c_double)
    rc = lib.ttirt_run_host(d, n.ctypes.data_as(lp), xs.ctypes.data_as(dp), r.ctypes.data_as(lp), cores.ctypes.data_as(dp),
                            M, q.ctypes.data_as(dp), Z.ctypes.data_as(dp), lPz.ctypes.data_as(dp),
                            idx.ctypes.data_as(c_void_p) if want_idx else None, int(mode), int(first_device), int(n_devices))
    if rc != 0:
        _raise_last(lib, "ttirt_run_host")
    return (Z, lPz, idx) if want_idx else (Z, lPz)


def run_uniform_host(n, xs, ranks, cores, M, seed, m0=0, mode=MODE_FAST, first_device=0, n_devices=1, want_q=False):
    """One-shot call with the seeds generated on the devices (Philox indices [m0, m0 + M)), rows sharded over n_devices
    GPUs: no q upload; the result does not depend on the device count."""
    from ctypes import c_ulonglong
    lib = load_library()
    n = np.ascontiguousarray(n, dtype=np.int64)
    r = np.ascontiguousarray(ranks, dtype=np.int64)
    xs = np.ascontiguousarray(np.asarray(xs, dtype=np.float64).ravel(order="F"))
    cores = np.ascontiguousarray(np.asarray(cores, dtype=np.float64).ravel(order="F"))
    d = int(n.size)
    Z = np.zeros((M, d), dtype=np.float64, order="F")
    lPz = np.zeros(M, dtype=np.float64)
    q = np.zeros((M, d), dtype=np.float64, order="F") if want_q else None
    lp, dp = POINTER(c_longlong), POINTER(c_double)
    f = lib.ttirt_run_uniform_host
    f.restype = c_int
    f.argtypes = [c_longlong, lp, dp, lp, dp, c_longlong, c_longlong, c_ulonglong, dp, dp, dp, c_int, c_int, c_int]
    rc = f(d, n.ctypes.data_as(lp), xs.ctypes.data_as(dp), r.ctypes.data_as(lp), cores.ctypes.data_as(dp), int(M), int(m0),
           int(seed) & 0xFFFFFFFFFFFFFFFF, q.ctypes.data_as(dp) if want_q else None, Z.ctypes.data_as(dp), lPz.ctypes.data_as(dp),
           int(mode), int(first_device), int(n_devices))
    if rc != 0:
        _raise_last(lib, "ttirt_run_uniform_host")
    return (Z, lPz, q) if want_q else (Z, lPz)


def kernel_launches():
    return int(load_library().ttirt_kernel_launches())


def device_count():
    return int(load_library().ttirt_device_count())
