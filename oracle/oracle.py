"""TEST INFRASTRUCTURE ONLY: ctypes loaders for the CPU oracle and the compiled reference.

`oracle_run`  -> oracle/liboracle_tt_irt1.so   (restatement of tt_irt1_int32.c:34-193)
`ref_run`     -> oracle/_ref/libref_tt_irt1_int{32,64}_{shim,openblas}.so, i.e. the
                 UNMODIFIED reference sources compiled by oracle/Makefile.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ORACLE_DIR = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}


def build(quiet=True):
    """Compile the oracle (and oracle/_ref when /root/reference is present)."""
    out = subprocess.run(["make", "-C", ORACLE_DIR], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if not quiet:
        print(out.stdout)


def _load(path):
    if path not in _LIBS:
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (run `make -C oracle`)")
        if path.endswith("_openblas.so"):
            _preload_blas_deps(path)
        _LIBS[path] = C.CDLL(path)
    return _LIBS[path]


def _preload_blas_deps(path):
    """The wheel-bundled OpenBLAS needs its sibling libquadmath/libgfortran, which carry no RUNPATH."""
    import glob
    import re
    out = subprocess.run(["readelf", "-d", path], capture_output=True, text=True).stdout
    m = re.search(r"(?:RUNPATH|RPATH).*\[(.*?)\]", out)
    if not m:
        return
    for pat in ("libquadmath*", "libgfortran*"):
        for f in sorted(glob.glob(os.path.join(m.group(1), pat))):
            try:
                C.CDLL(f, mode=C.RTLD_GLOBAL)
            except OSError:
                pass


def ref_lib_path(width=32, blas="shim"):
    return os.path.join(ORACLE_DIR, "_ref", "libref_tt_irt1_int%d_%s.so" % (width, blas))


def have_ref(width=32, blas="shim"):
    return os.path.exists(ref_lib_path(width, blas))


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).ravel(order="F"))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def oracle_run(n, xs, ranks, cores, q, rows=None, extras=False):
    """Run the restatement. q is (M, d) (any order; read column-major). Returns Z (M,d) F-order, lPz (M,)
    and, with extras=True, also idx (M,d) int32, kappa (M,d), gap (M,d), cond (M,d), lsens (M,d)
    (cond: sensitivity of x_k to rounding, lsens: |d log p / d x_k|; both feed the parity tolerances)."""
    lib = _load(os.path.join(ORACLE_DIR, "liboracle_tt_irt1.so"))
    q = np.asfortranarray(q, dtype=np.float64)
    M, d = q.shape
    n64 = np.ascontiguousarray(n, dtype=np.int64)
    r64 = np.ascontiguousarray(ranks, dtype=np.int64)
    xs = _f64(xs)
    cores = _f64(cores)
    assert n64.size == d and r64.size == d + 1 and xs.size == int(n64.sum())
    assert cores.size == int((r64[:-1] * n64 * r64[1:]).sum())
    Z = np.zeros((M, d), dtype=np.float64, order="F")
    lPz = np.zeros(M, dtype=np.float64)
    idx = np.zeros((M, d), dtype=np.int32, order="F") if extras else None
    kap = np.zeros((M, d), dtype=np.float64, order="F") if extras else None
    gap = np.zeros((M, d), dtype=np.float64, order="F") if extras else None
    cond = np.zeros((M, d), dtype=np.float64, order="F") if extras else None
    lsens = np.zeros((M, d), dtype=np.float64, order="F") if extras else None
    m0, m1 = (0, M) if rows is None else rows
    fn = lib.tt_irt1_oracle_rows
    fn.restype = C.c_int
    ip = C.POINTER(C.c_longlong)
    fn.argtypes = [C.c_longlong, ip, C.POINTER(C.c_double), ip, C.POINTER(C.c_double), C.c_longlong,
                   C.c_longlong, C.c_longlong, C.POINTER(C.c_double), C.POINTER(C.c_double),
                   C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_double),
                   C.POINTER(C.c_double), C.POINTER(C.c_double)]
    rc = fn(d, n64.ctypes.data_as(ip), _dp(xs), r64.ctypes.data_as(ip), _dp(cores), M, m0, m1,
            _dp(q), _dp(Z), _dp(lPz),
            idx.ctypes.data_as(C.POINTER(C.c_int)) if extras else None,
            _dp(kap) if extras else None, _dp(gap) if extras else None, _dp(cond) if extras else None, _dp(lsens) if extras else None)
    if rc != 0:
        raise RuntimeError("oracle failed")
    if extras:
        return Z, lPz, idx, kap, gap, cond, lsens
    return Z, lPz


def oracle_sweep(n, xs, ranks, cores):
    """Right-to-left marginalisation sweep only: returns (list of P_k (r_k, n_k) F-order, list of marg_k (r_{k+1},))."""
    lib = _load(os.path.join(ORACLE_DIR, "liboracle_tt_irt1.so"))
    n64 = np.ascontiguousarray(n, dtype=np.int64)
    r64 = np.ascontiguousarray(ranks, dtype=np.int64)
    d = n64.size
    xs = _f64(xs)
    cores = _f64(cores)
    pk = np.zeros(int((r64[:-1] * n64).sum()))
    mg = np.zeros(int(r64[1:].sum()))
    ip = C.POINTER(C.c_longlong)
    fn = lib.tt_irt1_oracle_sweep
    fn.restype = C.c_int
    fn.argtypes = [C.c_longlong, ip, C.POINTER(C.c_double), ip, C.POINTER(C.c_double),
                   C.POINTER(C.c_double), C.POINTER(C.c_double)]
    if fn(d, n64.ctypes.data_as(ip), _dp(xs), r64.ctypes.data_as(ip), _dp(cores), _dp(pk), _dp(mg)) != 0:
        raise RuntimeError("oracle sweep failed")
    P, Mg, op, om = [], [], 0, 0
    for k in range(d):
        sz = int(r64[k] * n64[k])
        P.append(pk[op:op + sz].reshape((int(r64[k]), int(n64[k])), order="F"))
        op += sz
        Mg.append(mg[om:om + int(r64[k + 1])])
        om += int(r64[k + 1])
    return P, Mg


def ref_run(n, xs, ranks, cores, q, width=32, blas="shim"):
    """Run the UNMODIFIED reference tt_irt1 (compiled into oracle/_ref). Returns Z (M,d) F-order, lPz (M,)."""
    lib = _load(ref_lib_path(width, blas))
    q = np.asfortranarray(q, dtype=np.float64)
    M, d = q.shape
    it, ct = (np.int32, C.c_int) if width == 32 else (np.int64, C.c_longlong)
    nn = np.ascontiguousarray(n, dtype=it)
    rr = np.ascontiguousarray(ranks, dtype=it)
    xs = _f64(xs)
    cores = _f64(cores)
    Z = np.zeros((M, d), dtype=np.float64, order="F")
    lPz = np.zeros(M, dtype=np.float64)
    fn = lib.tt_irt1
    fn.restype = None
    ip = C.POINTER(ct)
    fn.argtypes = [ct, ip, C.POINTER(C.c_double), ip, C.POINTER(C.c_double), ct,
                   C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    fn(d, nn.ctypes.data_as(ip), _dp(xs), rr.ctypes.data_as(ip), _dp(cores), M, _dp(q), _dp(Z), _dp(lPz))
    return Z, lPz
