#!/usr/bin/env python
"""Measurement of the squared-density transform, same contract as bench.py's line:

    python tests/devtools/bench_sqr.py [--shape d,n,r] [--log2m L] [--steps K] [--warmup W] [--no-cpu]

value        device-resident samples/s (ttirt_sqr_sample_device, q / Z in HBM, CUDA events on the launching stream)
e2e          the C symbol tt_irt_sqr on pinned HOST buffers: cores upload, sweep, copies, kernels inside the timed region
roofline     the conditional-pdf kernel (sqr_pdf_kernel): algorithmic flops (r (r + 1) n + r (r + 1) / 2 per sample and
             dimension, the symmetric half of tt_irt_sqr.m:107-112) / its summed launch time, against the live DMMA probe
cpu_baseline the numpy oracle (kind "port": the routine is Matlab-only, nothing to compile) on a bounded sample, one core
It lives under tests/ because it executes oracle/ for the CPU leg.  One JSON line on stdout.
"""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "tt-irt_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")

import bench  # noqa: E402  (ClockSampler, measured_fp64_peak)
from tt_irt_py import synth, tt_irt, tt_irt_sqr  # noqa: E402


def _ncu_traffic(d, n, r, M):
    """DRAM bytes per launch of sqr_pdf_kernel from the committed ncu capture of the same shape, scaled by the rows per launch
    (the capture ran one 2^18-row chunk; traffic is per row: interface row in, conditional out; the default chunk of this
    shape class is 2^19)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_sqr_ncu_traffic.json")) as f:
            j = json.load(f)
        if j["shape"] == "d=%d n=%d r=%d" % (d, n, r):
            chunk = int(os.environ.get("TTIRT_SQR_CHUNK", 0)) or (1 << (19 if n > 40 else 18))
            rows = min(M, chunk)
            return int((int(j["dram_bytes_read"]) + int(j["dram_bytes_write"])) * rows / int(j["rows_per_launch"]))
    except Exception:
        pass
    return None


def dirt_main(a):
    """DIRT sampler loop (tt_dirt_sample.m:17-73): nlvl + 1 tt_irt_sqr levels with the reference maps between them, samples
    resident on the device; synthetic levels of one shape, truncated-normal reference on [-4, 4]."""
    import torch
    d, n, r = [int(v) for v in a.shape.split(",")]
    M = 1 << a.log2m
    dev = torch.device("cuda", 0)
    levels = [synth.make_tt(d, n, r, seed=5, lo=-2.0, hi=3.0)] + [synth.make_tt(d, n, r, seed=6 + j, lo=-4.0, hi=4.0) for j in range(a.dirt)]
    drt = tt_irt_sqr.Dirt(levels, "Normal 4")
    gen = torch.Generator(device=dev)
    gen.manual_seed(1)
    q = (torch.rand((d, M), dtype=torch.float64, device=dev, generator=gen) * 2.0 - 1.0) * 3.99
    z = torch.empty((d, M), dtype=torch.float64, device=dev)
    lf = torch.empty((M,), dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream()

    def step():
        drt.sample_device(M, q.data_ptr(), M, z.data_ptr(), M, lf.data_ptr(), stream.cuda_stream)
    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    l1 = tt_irt.kernel_launches()
    sampler = bench.ClockSampler(0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(a.steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    clocks = sampler.stop()
    W = tt_irt_sqr.flops_per_sample(levels[0][0], levels[0][2]) * (a.dirt + 1)
    out = {"metric": "dirt_samples_per_sec", "value": M / (ms * 1e-3), "unit": "samples/s", "n_gpus": 1, "steps": a.steps, "warmup": a.warmup,
           "ms_per_step": ms, "higher_is_better": True, "dtype": "f64", "data": "synthetic",
           "config": {"workload": "DIRT sampler loop (tt_dirt_sample.m), %d levels of d=%d n=%d r=%d, truncated-normal reference on [-4,4], M=2^%d"
                                  % (a.dirt + 1, d, n, r, a.log2m)},
           "algorithmic_tflops_whole_step": W * M / (ms * 1e-3) / 1e12, "levels": a.dirt + 1,
           "gpu_launches": int(tt_irt.kernel_launches() - l1), "clocks": clocks,
           "finite": bool(torch.isfinite(lf).all() and torch.isfinite(z).all())}
    if not a.no_cpu:
        from oracle.dirt_oracle import tt_dirt_sample_oracle
        qs = np.asfortranarray((2.0 * synth.make_q(a.cpu_samples, d, seed=2) - 1.0) * 3.99)
        t0 = time.perf_counter()
        tt_dirt_sample_oracle(levels, qs, "Normal 4")
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": a.cpu_samples / dt, "unit": "samples/s", "cores": 1, "kind": "port",
                               "sample": "%d samples through oracle/dirt_oracle.py (numpy, one thread), %.1f s" % (a.cpu_samples, dt)}
    drt.close()
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="32,65,64")
    ap.add_argument("--log2m", type=int, default=20)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-samples", type=int, default=1 << 12)
    ap.add_argument("--dirt", type=int, default=-1, help="time the DIRT sampler loop (tt_dirt_sample.m) with this many levels above level 0, "
                                                         "normal reference on [-4, 4], instead of a single tt_irt_sqr")
    a = ap.parse_args()
    if a.dirt >= 0:
        return dirt_main(a)
    d, n, r = [int(v) for v in a.shape.split(",")]
    M = 1 << a.log2m
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench_sqr.py: no CUDA device (there is no CPU fallback)")
    dev = torch.device("cuda", 0)
    ns, xs, rk, c = synth.make_tt(d, n, r, seed=5)
    W = tt_irt_sqr.flops_per_sample(ns, rk)
    md = tt_irt_sqr.SqrModel(ns, xs, rk, c)     # (the first one pays context and module loading)
    t0 = time.perf_counter()
    md2 = tt_irt_sqr.SqrModel(ns, xs, rk, c)
    t_model = time.perf_counter() - t0
    md2.close()
    gen = torch.Generator(device=dev)
    gen.manual_seed(1)
    q = torch.rand((d, M), dtype=torch.float64, device=dev, generator=gen)
    z = torch.empty((d, M), dtype=torch.float64, device=dev)
    lf = torch.empty((M,), dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream()
    l0 = tt_irt.kernel_launches()

    def step():
        md.sample_device(M, d, q.data_ptr(), M, z.data_ptr(), M, lf.data_ptr(), None, stream.cuda_stream)

    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    md.profile_enable(True)
    sampler = bench.ClockSampler(0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l1 = tt_irt.kernel_launches()
    e0.record(stream)
    for _ in range(a.steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    launches = tt_irt.kernel_launches() - l1
    clocks = sampler.stop()
    kms, kn, kfl = md.profile_read()
    md.profile_enable(False)
    peak, peak_src = bench.measured_fp64_peak()
    finite = bool(torch.isfinite(lf).all() and torch.isfinite(z).all())
    out = {"metric": "irt_sqr_samples_per_sec", "value": M / (ms * 1e-3), "unit": "samples/s", "n_gpus": 1, "steps": a.steps, "warmup": a.warmup,
           "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": "squared-density IRT (tt_irt_sqr.m), synthetic random TT of sqrt(density) d=%d n=%d r=%d, M=2^%d uniform seeds; "
                                  "inputs larger than L2 (q+Z %.1f GB)" % (d, n, r, a.log2m, 16.0 * M * d / 1e9)},
           "roofline": {"bound": "tensor", "achieved": kfl / kms / 1e9 if kms > 0 else None, "peak": peak, "unit": "TFLOP/s",
                        "frac": (kfl / kms / 1e9) / peak if kms > 0 else None, "traffic": _ncu_traffic(d, n, r, M), "kernel": "sqr_pdf_kernel",
                        "kernel_avg_ms": kms / max(kn, 1), "kernel_share_of_step": kms / (ms * a.steps), "peak_source": peak_src,
                        "algorithmic_flops_per_sample": W},
           "algorithmic_tflops_whole_step": W * M / (ms * 1e-3) / 1e12,
           "gpu_launches": int(launches), "clocks": clocks, "model_create_ms": t_model * 1e3, "finite": finite}
    if not a.no_e2e:
        qh = torch.empty((d, M), dtype=torch.float64, pin_memory=True)
        qh.copy_(q)
        zh = torch.empty((d, M), dtype=torch.float64, pin_memory=True)
        lh = torch.empty((M,), dtype=torch.float64, pin_memory=True)
        lib = md._lib
        n32 = np.ascontiguousarray(ns, dtype=np.int32)
        r32 = np.ascontiguousarray(rk, dtype=np.int32)
        dp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int)

        def e2e_step():
            lib.tt_irt_sqr(d, n32.ctypes.data_as(ip), int(xs.size), xs.ctypes.data_as(dp), r32.ctypes.data_as(ip), c.ctypes.data_as(dp),
                           M, d, ctypes.cast(qh.data_ptr(), dp), ctypes.cast(zh.data_ptr(), dp), ctypes.cast(lh.data_ptr(), dp))
        for _ in range(max(1, a.warmup - 1)):
            e2e_step()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            e2e_step()
        dt = (time.perf_counter() - t0) / a.steps
        out["e2e"] = {"value": M / dt, "unit": "samples/s", "h2d_bytes_per_step": int(8 * M * d + 8 * c.size + 8 * xs.size),
                      "d2h_bytes_per_step": int(8 * M * (d + 1)), "ms_per_step": dt * 1e3,
                      "agrees_with_device_run": bool(np.array_equal(zh[:, :4096].numpy(), z[:, :4096].cpu().numpy()))}
    md.close()
    if not a.no_cpu:
        from oracle.tt_irt_sqr_oracle import tt_irt_sqr_oracle
        qs = synth.make_q(a.cpu_samples, d, seed=2)
        t0 = time.perf_counter()
        tt_irt_sqr_oracle(ns, xs, rk, c, qs)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": a.cpu_samples / dt, "unit": "samples/s", "cores": 1, "kind": "port",
                               "sample": "%d samples through oracle/tt_irt_sqr_oracle.py (numpy, OPENBLAS_NUM_THREADS=1; includes its sweep), %.1f s"
                                         % (a.cpu_samples, dt)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
