"""Development check on a GPU box for the wide path (csrc/ttirt_wide.cu): parity against the oracle on shapes beyond the
fused transition kernel, time against the strict kernel, and device-resident throughput against the DMMA roofline."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tt-irt_b200")); sys.path.insert(0, ROOT)
from tt_irt_py import synth, tt_irt
import oracle
from oracle import parity

shapes = [((1, 96, 96, 96, 1), (129, 129, 129, 129), 3000, "uniform"), ((1, 128, 128, 1), (33, 33, 33), 2500, "uniform"),
          ((1, 6, 6, 6, 1), (80, 300, 80, 80), 4000, "uniform"), ((1, 70, 130, 9, 1), (12, 90, 75, 5), 2000, "uniform"),
          ((1, 65, 67, 1), (73, 2, 74), 1500, "normal")]
for ranks, ns, M, cores in shapes:
    rk = np.array(ranks, dtype=np.int64); ns = np.array(ns, dtype=np.int64); d = ns.size
    rng = np.random.default_rng(int(rk.sum() + ns.sum()))
    xs = np.concatenate([np.sort(rng.uniform(-1.5, 2.5, size=n)) for n in ns])
    size = int((rk[:-1] * ns * rk[1:]).sum())
    c = rng.random(size) if cores == "uniform" else rng.standard_normal(size)
    q = synth.make_q(M, d, seed=5); q[0, :] = 0.0; q[1, :] = 1.0
    Zo, lo, io, kap, gap, cond, lsens = oracle.oracle_run(ns, xs, rk, c, q, extras=True)
    md = tt_irt.Model(ns, xs, rk, c)
    t = time.time(); Zs, ls, isx = md.sample(q, mode=tt_irt.MODE_STRICT, want_idx=True); ts = time.time() - t
    l0 = tt_irt.kernel_launches()
    t = time.time(); Zf, lf, ifx = md.sample(q, mode=tt_irt.MODE_FAST, want_idx=True); tf = time.time() - t
    launches = tt_irt.kernel_launches() - l0
    st_f, f_f = parity.compare(Zf, lf, ifx, Zo, lo, io, cond, gap, lsens=lsens)
    print(json.dumps({"ranks": ranks, "ns": [int(v) for v in ns], "M": M, "cores": cores, "launches": launches,
                      "strict_bitexact": bool(np.array_equal(Zs, Zo) and np.array_equal(isx, io)),
                      "fast": {**st_f, "fails": f_f}, "sec_strict": round(ts, 3), "sec_fast": round(tf, 3)}), flush=True)
    md.close()

if "--perf" in sys.argv:
    import torch
    peak = 37.1
    for (d, n, r, log2m) in [(8, 129, 128, 17), (8, 257, 256, 16), (6, 65, 96, 18)]:
        M = 1 << log2m
        ns, xs, rk, c = synth.make_tt(d, n, r, seed=3)
        Wf = synth.flops_per_sample(ns, rk)
        md = tt_irt.Model(ns, xs, rk, c, device=0)
        q = torch.rand((d, M), dtype=torch.float64, device="cuda:0")
        z = torch.empty_like(q); l = torch.empty((M,), dtype=torch.float64, device="cuda:0")
        st = torch.cuda.current_stream()
        res = {}
        for mode, name, steps in ((tt_irt.MODE_FAST, "fast", 3), (tt_irt.MODE_STRICT, "strict", 1)):
            md.sample_device(M, q.data_ptr(), M, z.data_ptr(), M, l.data_ptr(), None, mode, st.cuda_stream)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record(st)
            for _ in range(steps):
                md.sample_device(M, q.data_ptr(), M, z.data_ptr(), M, l.data_ptr(), None, mode, st.cuda_stream)
            e1.record(st); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            res[name] = {"ms": ms, "samples_per_s": M / ms * 1e3, "tflops": M * Wf / ms / 1e9, "frac_of_dmma_peak_37.1": M * Wf / ms / 1e9 / peak,
                         "finite": bool(torch.isfinite(l).all())}
        md.close()
        print(json.dumps({"perf": [d, n, r, M], "flops_per_sample": Wf, **res}), flush=True)
