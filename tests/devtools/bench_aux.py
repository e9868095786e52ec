"""Measurement of the rows either side of tt_irt1 on one B200: device-resident
throughput with CUDA events against the HBM roofline (MEASURED_PEAKS.json), the numpy restatement timed beside it.
usage (under gpurun): python tests/devtools/bench_aux.py > gpurun_out/aux_bench.json"""
import ctypes, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tt-irt_b200")); sys.path.insert(0, ROOT)
import torch
from tt_irt_py import synth, tt_irt
from oracle import samplers_oracle as so

lib = tt_irt.load_library()
vp, ll, ull, dbl = ctypes.c_void_p, ctypes.c_longlong, ctypes.c_ulonglong, ctypes.c_double
lib.ttirt_seeds_lattice_device.argtypes = [ll, ll, ll, ll, vp, vp, vp, ll, vp]
lib.ttirt_seeds_uniform_device.argtypes = [ll, ll, ll, ull, vp, ll, vp]
lib.ttirt_truncnormal_map_device.argtypes = [ll, dbl, vp, vp, vp]
lib.ttirt_iw_stats_device.argtypes = [ll, vp, vp, vp, ctypes.POINTER(dbl), vp]
lib.ttirt_mcmc_prune_device.argtypes = [ll, vp, vp, vp, vp, ctypes.POINTER(ll), vp, ll, vp]
try:
    HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    HBM = 6650.0
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream().cuda_stream


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


def cpu_timed(fn):
    t = time.perf_counter(); fn(); return time.perf_counter() - t


out = {}
d, M = 32, 1 << 24
q = torch.empty((d, M), dtype=torch.float64, device=dev)
z = torch.arange(1, 2 * d, 2, dtype=torch.float64, device=dev); sh = torch.rand(d, dtype=torch.float64, device=dev)
t = timed(lambda: lib.ttirt_seeds_lattice_device(d, M, 0, M, z.data_ptr(), sh.data_ptr(), q.data_ptr(), M, st))
tc = cpu_timed(lambda: so.qmc_lattice(d, 20, np.arange(1, 2 * d, 2), np.full(d, 0.3)))
out["seeds_lattice"] = {"elements_per_s": d * M / t, "GB_per_s": 8 * d * M / t / 1e9, "hbm_frac": 8 * d * M / t / 1e9 / HBM,
                        "cpu_numpy_elements_per_s": d * (1 << 20) / tc, "shape": "d=32, M=2^24 (config 3 seed matrix)"}
t = timed(lambda: lib.ttirt_seeds_uniform_device(d, M, 0, 12345, q.data_ptr(), M, st))
tc = cpu_timed(lambda: so.uniform_philox(d, 1 << 16, 12345))
out["seeds_uniform_philox"] = {"elements_per_s": d * M / t, "GB_per_s": 8 * d * M / t / 1e9, "hbm_frac": 8 * d * M / t / 1e9 / HBM,
                               "cpu_numpy_elements_per_s": d * (1 << 16) / tc}
y = torch.empty_like(q)
t = timed(lambda: lib.ttirt_truncnormal_map_device(d * M, 4.0, q.data_ptr(), y.data_ptr(), st))
un = np.random.default_rng(0).random(1 << 22)
tc = cpu_timed(lambda: so.truncnormal_map(un, 4.0))
out["truncnormal_map"] = {"elements_per_s": d * M / t, "GB_per_s": 16 * d * M / t / 1e9, "hbm_frac": 16 * d * M / t / 1e9 / HBM,
                          "cpu_scipy_elements_per_s": un.size / tc}
del y

Mi = 1 << 26
lfapp = torch.randn(Mi, dtype=torch.float64, device=dev) * 2 - 5
lfex = lfapp + 0.3 * torch.randn(Mi, dtype=torch.float64, device=dev) + 1.7
w = torch.empty(Mi, dtype=torch.float64, device=dev)
res = (dbl * 6)()
t = timed(lambda: lib.ttirt_iw_stats_device(Mi, lfex.data_ptr(), lfapp.data_ptr(), w.data_ptr(), res, st), reps=3)
fe, fa = lfex[: 1 << 22].cpu().numpy(), lfapp[: 1 << 22].cpu().numpy()
tc = cpu_timed(lambda: (so.iw_prune(fe, fa), so.essinv(fe, fa), so.hellinger(fe, fa)))
byts = Mi * (16 * 3 + 8)
out["iw_stats"] = {"samples_per_s": Mi / t, "GB_per_s": byts / t / 1e9, "hbm_frac": byts / t / 1e9 / HBM, "cpu_numpy_samples_per_s": (1 << 22) / tc,
                   "bytes_per_sample": 56, "note": "three passes over (lFex, lFapp) + weights written; M=2^26"}

Mm = 1 << 20
for spread in (0.2, 1.0, 3.0):
    fa = torch.randn(Mm, dtype=torch.float64, device=dev) - 3
    fe = fa + spread * torch.randn(Mm, dtype=torch.float64, device=dev)
    u = torch.rand(Mm, dtype=torch.float64, device=dev)
    src = torch.empty(Mm, dtype=torch.int32, device=dev)
    nrej = ll(0)
    t = timed(lambda: lib.ttirt_mcmc_prune_device(Mm, fe.data_ptr(), fa.data_ptr(), u.data_ptr(), src.data_ptr(), ctypes.byref(nrej), None, 0, st), reps=2)
    k = 1 << 16
    tc = cpu_timed(lambda: so.mcmc_prune(fe[:k].cpu().numpy(), fa[:k].cpu().numpy(), u[:k].cpu().numpy()))
    out["mcmc_prune_spread_%g" % spread] = {"samples_per_s": Mm / t, "ms_for_2^20": t * 1e3, "rejection_rate": nrej.value / Mm,
                                            "cpu_python_loop_samples_per_s": k / tc}
del q

# sampling end to end with seeds generated on the device vs uploaded seeds (pageable numpy outputs)
ns, xs, rk, c = synth.make_tt(32, 65, 64, seed=2026)
md = tt_irt.Model(ns, xs, rk, c)
Ms = 1 << 23
qh = synth.make_q(Ms, 32, seed=1)
md.sample(qh); md.sample_uniform(Ms, 7)   # warm-up at full size: the pinned bounce buffers are allocated on first use (~0.3 s per GB)
t0 = time.perf_counter(); md.sample(qh); t_up = time.perf_counter() - t0
t0 = time.perf_counter(); md.sample_uniform(Ms, 7); t_dev = time.perf_counter() - t0
out["sample_e2e_2^23"] = {"uploaded_seeds_samples_per_s": Ms / t_up, "device_seeds_samples_per_s": Ms / t_dev,
                          "note": "Model.sample (numpy q in, numpy Z out, Z allocated per call) vs Model.sample_uniform (no q upload)"}
print(json.dumps(out, indent=1))
