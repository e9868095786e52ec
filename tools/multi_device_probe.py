"""The drop-in symbol on several GPUs inside ONE process (TTIRT_DEVICES): BASELINE configs[4], M = 2^26 seed points
sharded over 1 / 2 / 4 / 8 B200 by ttirt_run_host (one host thread per device, no collective), host arrays pinned."""
import ctypes, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tt-irt_b200"))
import torch
from tt_irt_py import synth, tt_irt
d, n, r = 32, 65, 64
log2m = int(sys.argv[1]) if len(sys.argv) > 1 else 26
M = 1 << log2m
ns, xs, rk, c = synth.make_tt(d, n, r, seed=2026)
lib = tt_irt.load_library()
q = torch.empty((d, M), dtype=torch.float64, pin_memory=True)
q.uniform_(0.0, 1.0)
z = torch.empty((d, M), dtype=torch.float64, pin_memory=True); l = torch.empty(M, dtype=torch.float64, pin_memory=True)
n32, r32 = ns.astype(np.int32), rk.astype(np.int32)
dp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int)
res = {}
ref = None
only = [int(x) for x in os.environ.get("PROBE_DEVICES", "1,2,4,8").split(",")]
for g in [x for x in only if x <= torch.cuda.device_count()]:
    os.environ["TTIRT_DEVICES"] = str(g)
    def call():
        lib.tt_irt1(d, n32.ctypes.data_as(ip), xs.ctypes.data_as(dp), r32.ctypes.data_as(ip), c.ctypes.data_as(dp), M,
                    ctypes.cast(q.data_ptr(), dp), ctypes.cast(z.data_ptr(), dp), ctypes.cast(l.data_ptr(), dp))
    call()
    dt = 1e9
    for _ in range(3):
        t = time.perf_counter(); call(); dt = min(dt, time.perf_counter() - t)
    chk = float(l[: 1 << 16].sum())
    ref = chk if ref is None else ref
    res["devices_%d" % g] = {"samples_per_s": M / dt, "seconds": dt, "same_result_as_1_device": chk == ref}
print(json.dumps({"workload": "d=32 n=65 r=64 M=2^%d, one process, tt_irt1 with TTIRT_DEVICES" % log2m, **res}, indent=1))
