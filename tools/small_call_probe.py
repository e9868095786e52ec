"""Latency of small drop-in calls (BASELINE configs[0] shape): where the time of one tt_irt1 call goes."""
import os, sys, time, ctypes
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tt-irt_b200"))
from tt_irt_py import synth, tt_irt
for (d, n, r, log2m) in [(8, 17, 8, 14), (8, 17, 16, 16), (11, 17, 16, 16)]:
    M = 1 << log2m
    ns, xs, rk, c = synth.make_tt(d, n, r, seed=1)
    f = tt_irt.TTTensor(ns, rk, c)
    q = synth.make_q(M, d, seed=2)
    for _ in range(3):
        tt_irt.tt_irt1(q, f, xs)
    t = time.perf_counter()
    reps = 50
    for _ in range(reps):
        tt_irt.tt_irt1(q, f, xs)
    dt = (time.perf_counter() - t) / reps
    md = tt_irt.Model(ns, xs, rk, c)
    for _ in range(3):
        md.sample(q)
    t = time.perf_counter()
    for _ in range(reps):
        md.sample(q)
    dm = (time.perf_counter() - t) / reps
    md.close()
    print("d=%d n=%d r=%d M=2^%d: tt_irt1 %.3f ms per call (%.1f M samples/s), cached Model.sample %.3f ms" % (d, n, r, log2m, dt * 1e3, M / dt / 1e6, dm * 1e3))
os.environ["TTIRT_TRACE"] = "1"
