"""End-to-end timing of the drop-in tt_irt1 symbol: pageable vs page-locked caller buffers."""
import os, sys, time, ctypes
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tt-irt_b200"))
from tt_irt_py import synth, tt_irt
log2m = int(sys.argv[1]) if len(sys.argv) > 1 else 24
d, n, r, M = 32, 65, 64, 1 << log2m
ns, xs, rk, c = synth.make_tt(d, n, r, seed=2026)
f = tt_irt.TTTensor(ns, rk, c)
q = np.asfortranarray(np.random.default_rng(0).random((d, M)).T)
for it in range(3):
    t = time.perf_counter(); Z, l = tt_irt.tt_irt1(q, f, xs); dt = time.perf_counter() - t
    print("wrapper tt_irt1 (numpy pageable, allocates Z): %.3f s  %.2f M samples/s  TTIRT_PIN=%s" % (dt, M / dt / 1e6, os.environ.get("TTIRT_PIN")))
lib = tt_irt.load_library()
Z = np.zeros((M, d), order="F"); l = np.zeros(M)
n32 = ns.astype(np.int32); r32 = rk.astype(np.int32)
dp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int)
for it in range(3):
    t = time.perf_counter()
    lib.tt_irt1(d, n32.ctypes.data_as(ip), xs.ctypes.data_as(dp), r32.ctypes.data_as(ip), c.ctypes.data_as(dp), M, q.ctypes.data_as(dp), Z.ctypes.data_as(dp), l.ctypes.data_as(dp))
    dt = time.perf_counter() - t
    print("C symbol, preallocated pageable outputs: %.3f s  %.2f M samples/s" % (dt, M / dt / 1e6))
