"""Per-phase cycle breakdown of the transition kernel (debug build with -DTTIRT_PHASE_TIMING).
usage: python tools/phase_timing.py tools/pt_w8.so [log2m]"""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tt-irt_b200"))
from tt_irt_py import synth
import torch
lib = ctypes.CDLL(os.path.abspath(sys.argv[1]))
log2m = int(sys.argv[2]) if len(sys.argv) > 2 else 20
d, n, r = [int(x) for x in os.environ.get("PT_SHAPE", "32,65,64").split(",")]
M = 1 << log2m
ns, xs, rk, c = synth.make_tt(d, n, r, seed=2026)
lp, dp = ctypes.POINTER(ctypes.c_longlong), ctypes.POINTER(ctypes.c_double)
lib.ttirt_model_create.restype = ctypes.c_void_p
lib.ttirt_model_create.argtypes = [ctypes.c_longlong, lp, dp, lp, dp, ctypes.c_int]
lib.ttirt_sample_device.argtypes = [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
n64 = ns.astype(np.int64); r64 = rk.astype(np.int64)
md = lib.ttirt_model_create(d, n64.ctypes.data_as(lp), xs.ctypes.data_as(dp), r64.ctypes.data_as(lp), c.ctypes.data_as(dp), 0)
q = torch.rand((d, M), dtype=torch.float64, device="cuda"); z = torch.empty_like(q); l = torch.empty(M, dtype=torch.float64, device="cuda")
out = (ctypes.c_ulonglong * 40)()
for it in range(3):
    lib.ttirt_sample_device(ctypes.c_void_p(md), M, ctypes.c_void_p(q.data_ptr()), M, ctypes.c_void_p(z.data_ptr()), M, ctypes.c_void_p(l.data_ptr()), None, 0, None)
    torch.cuda.synchronize()
    lib.ttirt_debug_phase_cycles(out)
v = np.array(list(out)[:8], dtype=np.float64)
per_warp = np.array(list(out)[8:], dtype=np.float64)
names = ["slab/bin", "wait TMA", "update MMA", "issue+F' store", "pdf MMA+tailcol", "cdf+search", "inversion tail+out", "loop head"]
tot = v.sum()
warps = int(os.environ.get("PT_WARPS", "8")); mt = int(os.environ.get("PT_MT", "2"))
ntiles = (d - 1.0) * M / (8 * mt)      # warp-tiles per call (d-1 transitions)
for nm, x in zip(names, v):
    print("%-20s %6.2f%%  %8.0f cycles per warp-tile" % (nm, 100 * x / tot, x / ntiles))
print("total %.0f cycles per warp-tile" % (tot / ntiles))
tail = np.array(list(out)[24:30], dtype=np.float64)
uses = (d - 1.0) * M / 16      # parked tiles per call
for nm, x in zip(["tail: wait FULL", "tail: pass", "tail: search+release", "tail: inversion+out (per pair)", "tail: loop"], tail):
    print("%-32s %8.0f cycles per parked tile" % (nm, x / uses))
print("loop time per warp id (mean cycles per launch per CTA):", np.round(per_warp[:warps] / ((d - 1.0) * 148)).astype(int).tolist())

# timeline of the two MMA warps of sub-partition 0 of CTA 0 (last launch): clock at the end of every phase
tr = (ctypes.c_longlong * (2 * 24 * 8))()
lib.ttirt_debug_trace(tr)
t = np.array(list(tr), dtype=np.int64).reshape(2, 24, 8)
t0 = t[t > 0].min()
order = [7, 0, 1, 2, 3, 4, 5, 6]   # loop top, slab/bin done, rows arrived, update done, gather issued, pdf done, buffer free, parked
print("timeline (cycles since first mark): tile | producer 0: top slab rows upd issue pdf free parked | producer 1: same")
for i in range(2, 14):
    print("%2d | %s | %s" % (i, " ".join("%7d" % (t[0, i, k] - t0) for k in order), " ".join("%7d" % (t[1, i, k] - t0) for k in order)))
