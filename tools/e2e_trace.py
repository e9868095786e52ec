"""Where the time of one drop-in call goes (TTIRT_TRACE=1 lines on stderr): three BASELINE shapes, pinned and pageable host arrays."""
import ctypes, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tt-irt_b200"))
import torch
from tt_irt_py import synth, tt_irt
lib = tt_irt.load_library()
dp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int)
os.environ["TTIRT_DEVICES"] = "1"
for (d, n, r, log2m) in [(8, 17, 8, 14), (11, 17, 16, 20), (40, 33, 32, 22), (32, 65, 64, 24)]:
    M = 1 << log2m
    ns, xs, rk, c = synth.make_tt(d, n, r, seed=1)
    n32, r32 = ns.astype(np.int32), rk.astype(np.int32)
    qh = torch.rand((d, M), dtype=torch.float64).pin_memory(); zh = torch.empty((d, M), dtype=torch.float64).pin_memory(); lh = torch.empty(M, dtype=torch.float64).pin_memory()
    def call():
        lib.tt_irt1(d, n32.ctypes.data_as(ip), xs.ctypes.data_as(dp), r32.ctypes.data_as(ip), c.ctypes.data_as(dp), M,
                    ctypes.cast(qh.data_ptr(), dp), ctypes.cast(zh.data_ptr(), dp), ctypes.cast(lh.data_ptr(), dp))
    quiet = os.dup(2)
    if os.environ.get("TTIRT_TRACE"):      # warm-up calls and the timing loop without the trace lines
        dn = os.open(os.devnull, os.O_WRONLY); os.dup2(dn, 2)
    for _ in range(3): call()
    reps = 20 if log2m < 22 else 3
    t = time.perf_counter()
    for _ in range(reps): call()
    dt = (time.perf_counter() - t) / reps
    os.dup2(quiet, 2)
    print("shape d=%d n=%d r=%d M=2^%d pinned: %.3f ms per call, %.1f M samples/s" % (d, n, r, log2m, 1e3 * dt, M / dt / 1e6), flush=True)
    if os.environ.get("TTIRT_TRACE"):
        call()
        sys.stderr.write("---- trace of one call, d=%d n=%d r=%d M=2^%d\n" % (d, n, r, log2m)); sys.stderr.flush()
