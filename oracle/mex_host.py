"""TEST INFRASTRUCTURE ONLY: calls the reference's MEX function tracemult (matlab/utils/tracemult.c, compiled UNMODIFIED against
the stand-in mex.h under oracle/mexstub/ into oracle/_ref/libref_tracemult.so by oracle/Makefile) from Python.

    ref_tracemult(A, j)        C(i) = A(i, j(i))                      A: n x s
    ref_tracemult(A, j, B)     C(:,:,i) = A(:,:,i) * B(:,:,j(i))      A: p x m x n, B: m x k x s   (tracemult.c:103-112, 131-136)
Arrays go in and come out in Matlab's column-major order; j is 1-based, as the MEX file reads it.

It also calls the reference's MEX GATEWAY of the hot path, matlab/utils/tt_irt_mex.c (what Matlab runs for
[Z, lPz] = tt_irt_mex(f.n, cell2mat(xsf), f.r, f.core, Z), install.m:169), compiled unmodified twice by oracle/Makefile:
    tt_irt_mex("reference", ...)   gateway + the reference's own tt_irt1_int64.c            (oracle/_ref/libref_tt_irt_mex.so)
    tt_irt_mex("b200", ...)        the same gateway linked against tt-irt_b200/lib/libtt_irt1_int64.so instead, i.e. the swap
                                   INTEGRATION.md describes to a maintainer                 (oracle/_ref/libref_tt_irt_mex_b200.so)
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def available():
    return os.path.exists(os.path.join(_HERE, "_ref", "libref_tracemult.so"))


def _lib():
    global _LIB
    if _LIB is None:
        lib = ctypes.CDLL(os.path.join(_HERE, "_ref", "libref_tracemult.so"))
        vp, sz = ctypes.c_void_p, ctypes.c_size_t
        lib.mexstub_wrap.restype = vp
        lib.mexstub_wrap.argtypes = [vp, sz, ctypes.POINTER(sz)]
        lib.mexstub_free_wrapper.argtypes = [vp]
        lib.mexstub_free_array.argtypes = [vp]
        lib.mexstub_ndim.restype = sz
        lib.mexstub_ndim.argtypes = [vp]
        lib.mexstub_dim.restype = sz
        lib.mexstub_dim.argtypes = [vp, sz]
        lib.mexstub_data.restype = ctypes.POINTER(ctypes.c_double)
        lib.mexstub_data.argtypes = [vp]
        lib.mexstub_call1.restype = vp
        lib.mexstub_call1.argtypes = [ctypes.c_int, ctypes.POINTER(vp)]
        _LIB = lib
    return _LIB


def ref_tracemult(A, j, B=None):
    lib = _lib()
    keep, wrapped = [], []
    for x in ([A, j] if B is None else [A, j, B]):
        a = np.asfortranarray(np.asarray(x, dtype=np.float64))
        if a.ndim < 2:
            a = a.reshape((-1, 1), order="F")
        keep.append(a)
        dims = (ctypes.c_size_t * a.ndim)(*a.shape)
        wrapped.append(lib.mexstub_wrap(a.ctypes.data_as(ctypes.c_void_p), a.ndim, dims))
    prhs = (ctypes.c_void_p * len(wrapped))(*wrapped)
    out = lib.mexstub_call1(len(wrapped), prhs)
    for w in wrapped:
        lib.mexstub_free_wrapper(w)
    if not out:
        raise RuntimeError("tracemult returned no output (it prints its complaint to stderr and returns)")
    nd = int(lib.mexstub_ndim(out))
    shape = [int(lib.mexstub_dim(out, i)) for i in range(nd)]
    n = int(np.prod(shape))
    res = np.ctypeslib.as_array(lib.mexstub_data(out), shape=(max(n, 1),))[:n].copy().reshape(shape, order="F")
    lib.mexstub_free_array(out)
    while res.ndim > 2 and res.shape[-1] == 1:
        res = res.reshape(res.shape[:-1], order="F")
    return np.asfortranarray(res)


_GATEWAYS = {"reference": "libref_tt_irt_mex.so", "b200": "libref_tt_irt_mex_b200.so"}
_GW = {}


def gateway_available(which):
    return os.path.exists(os.path.join(_HERE, "_ref", _GATEWAYS[which]))


def _gateway(which):
    if which not in _GW:
        _GW[which] = _bind(ctypes.CDLL(os.path.join(_HERE, "_ref", _GATEWAYS[which])))
    return _GW[which]


def _bind(lib):
    vp, sz = ctypes.c_void_p, ctypes.c_size_t
    lib.mexstub_wrap.restype = vp
    lib.mexstub_wrap.argtypes = [vp, sz, ctypes.POINTER(sz)]
    lib.mexstub_free_wrapper.argtypes = [vp]
    lib.mexstub_free_array.argtypes = [vp]
    lib.mexstub_ndim.restype = sz
    lib.mexstub_ndim.argtypes = [vp]
    lib.mexstub_dim.restype = sz
    lib.mexstub_dim.argtypes = [vp, sz]
    lib.mexstub_data.restype = ctypes.POINTER(ctypes.c_double)
    lib.mexstub_data.argtypes = [vp]
    lib.mexFunction.restype = None
    lib.mexFunction.argtypes = [ctypes.c_int, ctypes.POINTER(vp), ctypes.c_int, ctypes.POINTER(vp)]
    return lib


def call_mex(lib, args, nlhs):
    """mexFunction(nlhs, plhs, nrhs, prhs) of a gateway built against oracle/mexstub/ (lib: ctypes.CDLL or a path); args: arrays,
    handed over as double mxArrays in column-major order (1-d inputs become columns).  Returns the nlhs outputs as numpy arrays."""
    if isinstance(lib, str):
        lib = _bind(ctypes.CDLL(lib))
    keep, wrapped = [], []
    for x in args:
        a = np.asfortranarray(np.asarray(x, dtype=np.float64))
        if a.ndim < 2:
            a = a.reshape((-1, 1), order="F")
        keep.append(a)
        dims = (ctypes.c_size_t * a.ndim)(*a.shape)
        wrapped.append(lib.mexstub_wrap(a.ctypes.data_as(ctypes.c_void_p), a.ndim, dims))
    prhs = (ctypes.c_void_p * len(wrapped))(*wrapped)
    plhs = (ctypes.c_void_p * max(1, nlhs))(*([None] * max(1, nlhs)))
    lib.mexFunction(nlhs, plhs, len(wrapped), prhs)
    for w in wrapped:
        lib.mexstub_free_wrapper(w)
    if any(not plhs[i] for i in range(nlhs)):
        raise RuntimeError("the MEX function assigned no outputs (it prints its complaint and returns)")
    outs = []
    for i in range(nlhs):
        o = plhs[i]
        shape = [int(lib.mexstub_dim(o, k)) for k in range(int(lib.mexstub_ndim(o)))]
        cnt = int(np.prod(shape))
        outs.append(np.asfortranarray(np.ctypeslib.as_array(lib.mexstub_data(o), shape=(max(cnt, 1),))[:cnt].copy().reshape(shape, order="F")))
        lib.mexstub_free_array(o)
    return outs


def tt_irt_mex(which, n, xs, ttrank, ttcore, q):
    """[Z, lPz] = tt_irt_mex(n, xs, ttrank, ttcore, q) through the reference's gateway source (tt_irt_mex.c:7-42): everything is
    passed as Matlab passes it, double arrays (n and ttrank included; the gateway converts them to mwIndex, :23-29).
    Returns Z (M x d, column-major) and lPz (M,)."""
    Z, l = call_mex(_gateway(which), [np.asarray(n, dtype=np.float64).reshape(-1, 1), np.asarray(xs, dtype=np.float64).reshape(-1, 1),
                                      np.asarray(ttrank, dtype=np.float64).reshape(-1, 1), np.asarray(ttcore, dtype=np.float64).reshape(-1, 1), q], 2)
    return Z, l.reshape(-1)
