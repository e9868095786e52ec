#!/usr/bin/env python
"""bench.py -- IRT samples/s of the B200-native tt_irt1 on BASELINE.json's metric configuration.

    python bench.py [--gpus N] [--steps K] [--warmup W]          (N>1: launched by torch.distributed.run)
    python bench.py --impl reference [...]                        the reference's own CPU path, host cores

A "step" is one pass of the hot path (all d dimensions) over one batch of M synthetic seed points.

`value`  BASELINE.json configs[2] (synthetic random TT density d=32, n=65, r=64, M=2^24 uniform points) per GPU, through
         the device-resident entry point (ttirt_sample_device, inputs already in HBM), CUDA events on the launching
         stream, max over ranks.  At N>1 these are N replicas of the one-GPU job ("scaling": "weak"): samples are
         independent (tt_irt1_int32.c:88-181), there is no collective, so this line only shows that the GPUs do not
         disturb each other.
`e2e`    the reference's own C-ABI symbol `tt_irt1` on pinned HOST buffers, everything inside the timed region (cores
         upload, marginalisation sweep, H2D of q, kernels, D2H of Z and lPz).
           N=1 : configs[2], M=2^24, one GPU.
           N>1 : configs[4], ONE call of tt_irt1 by rank 0 on M=2^26 seed points with TTIRT_DEVICES=N -- the product's own
                 multi-GPU path (ttirt_run_host: one host thread per device, contiguous row shards, cores fanned out
                 by peer copies); the other ranks wait on a host-side (gloo) barrier.  Beside it: the same call with the
                 seeds generated on the devices (ttirt_run_uniform_host, no q upload) and the box's host-copy ceiling
                 measured with plain pinned copies on the same buffers.
`roofline` the dominant kernel (transition_kernel) timed launch by launch with CUDA events in a serialised pass of the
         same workload right after the timed region (the timed region itself overlaps chunks on two streams, where
         per-launch events would also measure the overlap), against the live FP64 DMMA probe.
One JSON line on stdout (rank 0).
"""
import argparse
import glob
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tt-irt_b200"))
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")

from tt_irt_py import synth  # noqa: E402

METRIC = "irt_samples_per_sec"
UNIT = "samples/s"
FP64_PEAK_FALLBACK_TFLOPS = 37.17  # profiles/r01_fp64_peak.md (measured DMMA, this pool); MEASURED_PEAKS.json has no FP64 entry


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log2m", type=int, default=24, help="log2 of samples per GPU per step (default: configs[2], 2^24)")
    ap.add_argument("--log2m-sharded", type=int, default=26, help="log2 of the samples of the one sharded call at N>1 (configs[4], 2^26)")
    ap.add_argument("--shape", default="32,65,64", help="d,n,r (default: the metric configuration)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-next-rows", action="store_true", help="skip the informational tt_irt_sqr figure appended at N=1")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the other BASELINE shapes appended at N=1")
    ap.add_argument("--cpu-samples-per-core", type=int, default=1 << 14)
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# CPU baseline: the UNMODIFIED reference C (oracle/_ref, kind "reference") or the oracle port, sharded
# over all host cores in separate processes (the reference is single-threaded by construction)
# ------------------------------------------------------------------------------------------------
_CPU_CTX = {}


def _cpu_worker(i):
    # inputs are inherited through fork (no pickling of the 64 MB cores)
    kind, width, blas, ns, xs, rk, cores = _CPU_CTX["args"]
    q = _CPU_CTX["qs"][i]
    import oracle
    t0 = time.perf_counter()
    if kind == "reference":
        oracle.ref_run(ns, xs, rk, cores, q, width=width, blas=blas)
    else:
        oracle.oracle_run(ns, xs, rk, cores, q)
    return time.perf_counter() - t0


def _cpu_worker_warm(i):
    kind, width, blas, ns, xs, rk, cores = _CPU_CTX["args"]
    import oracle
    q = _CPU_CTX["qs"][i][:64]
    if kind == "reference":
        oracle.ref_run(ns, xs, rk, cores, q, width=width, blas=blas)
    else:
        oracle.oracle_run(ns, xs, rk, cores, q)
    return 0


def cpu_baseline(ns, xs, rk, cores, d, samples_per_core, repeats=1):
    """Samples/s of the reference's CPU path on all host cores for a bounded sample of the workload."""
    import multiprocessing as mp
    import oracle
    ncores = os.cpu_count() or 1
    if oracle.have_ref(32, "openblas"):
        kind, blas = "reference", "openblas"
    elif oracle.have_ref(32, "shim"):
        kind, blas = "reference", "shim"
    else:
        kind, blas = "port", ""
        oracle.build()
    # The library is loaded (and run once) in THIS process before the workers fork: they inherit the mapping, and the
    # driver's loaded-library record of this process shows oracle/_ref/... (the workers themselves end with the pool).
    qw = synth.make_q(64, d, seed=999)
    if kind == "reference":
        oracle.ref_run(ns, xs, rk, cores, qw, width=32, blas=blas)
        loaded = os.path.relpath(oracle.ref_lib_path(32, blas), ROOT)
    else:
        oracle.oracle_run(ns, xs, rk, cores, qw)
        loaded = "oracle/liboracle_tt_irt1.so"
    _CPU_CTX["args"] = (kind, 32, blas, ns, xs, rk, cores)
    _CPU_CTX["qs"] = [synth.make_q(samples_per_core, d, seed=1000 + i) for i in range(ncores)]
    ctx = mp.get_context("fork")
    best = None
    pool = ctx.Pool(ncores)
    try:
        pool.map(_cpu_worker_warm, range(ncores))  # first touch outside the timed region
        for _ in range(repeats):
            t0 = time.perf_counter()
            pool.map(_cpu_worker, range(ncores), chunksize=1)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    finally:
        pool.close()
        pool.join()
    total = samples_per_core * ncores
    return {"value": total / best, "unit": UNIT, "cores": ncores, "kind": kind, "library": loaded,
            "sample": "%d samples (%d per core x %d processes, OPENBLAS_NUM_THREADS=1%s), %.1f s wall" %
                      (total, samples_per_core, ncores, ", BLAS=" + blas if blas else "", best)}, best


# ------------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler(object):
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            out = self.p.communicate(timeout=5)[0]
        except Exception:
            self.p.kill()
            out = ""
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        busy = [s for s, p in zip(sm, pw) if p > 300.0] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_fp64_peak():
    """FP64 roofline denominator: tools/fp64_peak (DMMA/DFMA register loops) run live on this GPU when the
    binary is present, else the figure recorded in profiles/r01_fp64_peak.md."""
    exe = os.path.join(ROOT, "tools", "fp64_peak")
    if os.path.exists(exe):
        try:
            out = subprocess.run([exe, "--quick"], capture_output=True, text=True, timeout=120).stdout
            j = json.loads(out.strip().splitlines()[-1])
            pk = max(j["dmma_sustained_tflops"], j["dfma_sustained_tflops"])
            if 5.0 < pk < 100.0:
                return pk, "tools/fp64_peak live on this GPU (DMMA %.2f, DFMA %.2f TFLOP/s sustained); MEASURED_PEAKS.json has no FP64 entry" % (
                    j["dmma_sustained_tflops"], j["dfma_sustained_tflops"])
        except Exception:
            pass
    return FP64_PEAK_FALLBACK_TFLOPS, "recorded DMMA peak of profiles/r01_fp64_peak.md (fallback); MEASURED_PEAKS.json has no FP64 entry"


def _c_symbol(lib, width):
    """(callable, int numpy type) of the reference's entry point in the library of the given integer width."""
    import ctypes as C
    ct = C.c_int if width == 32 else C.c_longlong
    dp, ip = C.POINTER(C.c_double), C.POINTER(ct)
    lib.tt_irt1.restype = None
    lib.tt_irt1.argtypes = [ct, ip, dp, ip, dp, ct, dp, dp, dp]
    return lib.tt_irt1, (np.int32 if width == 32 else np.int64), ip, dp


def main():
    a = parse()
    d, n, r = [int(x) for x in a.shape.split(",")]
    M = 1 << a.log2m
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    ns, xs, rk, cores = synth.make_tt(d, n, r, seed=2026)
    W = synth.flops_per_sample(ns, rk)
    workload = "synthetic random TT density d=%d n=%d r=%d, M=2^%d uniform points per GPU (BASELINE.json configs[2]%s)" % (
        d, n, r, a.log2m, "; inputs >> L2" if M * d * 8 > (1 << 28) else "")
    config = {"workload": workload, "d": d, "n": n, "r": r, "M_per_gpu": M, "flops_per_sample": W,
              "bytes_per_sample": 8 * (2 * d + 1), "l2_policy": "inputs (%.1f GB of q and Z per step) far exceed the 126 MB L2" % (2 * M * d * 8 / 1e9)}
    if world > 1:
        config["value_is"] = "%d independent replicas of the one-GPU job (no data-path collective exists); the product's sharded call is e2e" % world

    # ---------------------------------------------------------------- reference arm (CPU) ----
    if a.impl == "reference":
        if rank != 0:
            return 0
        spc = a.cpu_samples_per_core
        # warm-up steps are run too (page cache, library load); every step is the same bounded sample
        for _ in range(max(0, min(a.warmup, 1))):
            cpu_baseline(ns, xs, rk, cores, d, max(256, spc // 16))
        times = []
        cb = None
        for _ in range(a.steps):
            cb, dt = cpu_baseline(ns, xs, rk, cores, d, spc)
            times.append(dt)
        total = spc * (os.cpu_count() or 1)
        val = total * len(times) / sum(times)
        cb["value"] = val
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
                "warmup": a.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config, "cpu_baseline": cb,
                "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0,
                "note": "reference CPU path (single-threaded C + BLAS) sharded over all host cores; each step is a bounded sample of the workload"}
        print(json.dumps(line))
        return 0

    # ---------------------------------------------------------------- B200 arm ----------------
    import torch
    from tt_irt_py import tt_irt
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (this framework has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    host_group = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        host_group = dist.new_group(backend="gloo")   # host-side waits: an NCCL barrier would spin ON the GPUs rank 0 is timing
    dev = torch.device("cuda", local_rank)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    md = tt_irt.Model(ns, xs, rk, cores, device=local_rank)
    lib = tt_irt.load_library()
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    q = torch.rand((d, M), dtype=torch.float64, device=dev, generator=gen)  # column-major M x d
    z = torch.empty((d, M), dtype=torch.float64, device=dev)
    lpz = torch.empty((M,), dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream()

    def step():
        md.sample_device(M, q.data_ptr(), M, z.data_ptr(), M, lpz.data_ptr(), None, tt_irt.MODE_FAST, stream.cuda_stream)

    for _ in range(a.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    l0 = tt_irt.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(a.steps):
        step()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = tt_irt.kernel_launches() - l0
    clocks = sampler.stop() if sampler else None
    # serialised pass of the same workload for the per-launch kernel time (see the module docstring)
    md.profile_enable(True)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record(stream)
    for _ in range(a.steps):
        step()
    p1.record(stream)
    torch.cuda.synchronize()
    ms_serial = p0.elapsed_time(p1)
    k_ms, k_launches, k_flops = md.profile_read()
    md.profile_enable(False)
    tmax = torch.tensor([ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_all = float(tmax.item())
    value = world * M * a.steps / (ms_all * 1e-3)
    checksum = float(lpz[: 1 << 10].sum().item())  # D2H read of a result so nothing is elided

    e2e = None
    if not a.no_e2e:
        del z, lpz
        if world == 1:
            e2e = e2e_single(a, lib, torch, ns, xs, rk, cores, d, M, q, local_rank)
            del q
        else:
            del q
            torch.cuda.empty_cache()
            dist.barrier(group=host_group)
            if rank == 0:
                e2e = e2e_sharded(a, lib, torch, ns, xs, rk, cores, d, world)
            dist.barrier(group=host_group)   # ranks != 0 sleep here on the host while rank 0 drives all the GPUs

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    peak, peak_src = measured_fp64_peak()
    ach = (k_flops / (k_ms * 1e-3)) / 1e12 if k_ms > 0 else 0.0
    traffic, traffic_src = _ncu_traffic(M)
    roofline = {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak if peak else None,
                "traffic": traffic, "traffic_unit": "bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)",
                "traffic_source": traffic_src,
                "kernel": "ttirt::transition_kernel<RT,NT,EXACT,TAIL1> (FP64 DMMA mma.sync.m8n8k4; 8 MMA warps + 4 tail warps per CTA)",
                "kernel_launches_timed": k_launches, "kernel_avg_ms": k_ms / max(1, k_launches),
                "kernel_timing": "CUDA events around every launch in a serialised pass of the same %d steps right after the timed region" % a.steps,
                "kernel_share_of_serial_step": k_ms / ms_serial if ms_serial > 0 else None,
                "serial_ms_per_step": ms_serial / a.steps,
                "flops_per_launch": k_flops / max(1, k_launches),
                "peak_source": peak_src,
                "whole_step_frac": (value / world * W / 1e12) / peak,
                "hbm_frac_of_measured": (value / world * 8 * (2 * d + 1) / 1e9) / _hbm_peak()}
    cb = None
    if not a.no_cpu and world == 1:
        cb, _ = cpu_baseline(ns, xs, rk, cores, d, a.cpu_samples_per_core)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_all / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config, "roofline": roofline, "cpu_baseline": cb, "e2e": e2e,
            "gpu_launches": int(launches), "clocks": clocks, "checksum_lpz_1k": checksum}
    if world == 1 and not a.no_other_configs and a.shape == "32,65,64":
        md.close()
        torch.cuda.empty_cache()
        lib.ttirt_cache_clear()
        line["other_configs"] = other_configs(torch, peak)
    if world == 1 and not a.no_next_rows:
        line["next_rows"] = _next_rows()
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


# ------------------------------------------------------------------------------------------------
# e2e, one GPU: the drop-in symbol on pinned host buffers (and on pageable numpy arrays, for information)
# ------------------------------------------------------------------------------------------------
def e2e_single(a, lib, torch, ns, xs, rk, cores, d, M, q_dev, local_rank):
    from ctypes import c_int, cast
    fn, it, ip, dp = _c_symbol(lib, 32)
    qh = torch.empty((d, M), dtype=torch.float64, pin_memory=True)
    qh.copy_(q_dev)
    del q_dev
    torch.cuda.empty_cache()
    zh = torch.empty((d, M), dtype=torch.float64, pin_memory=True)
    lh = torch.empty((M,), dtype=torch.float64, pin_memory=True)
    n32 = np.ascontiguousarray(ns, dtype=it)
    r32 = np.ascontiguousarray(rk, dtype=it)
    os.environ["TTIRT_DEVICE"] = str(local_rank)
    os.environ["TTIRT_DEVICES"] = "1"

    def e2e_step():
        fn(c_int(d), n32.ctypes.data_as(ip), xs.ctypes.data_as(dp), r32.ctypes.data_as(ip), cores.ctypes.data_as(dp),
           c_int(M), cast(qh.data_ptr(), dp), cast(zh.data_ptr(), dp), cast(lh.data_ptr(), dp))

    for _ in range(min(a.warmup, 2)):
        e2e_step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        e2e_step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if not bool(torch.isfinite(lh[: 1 << 12]).all()):
        raise SystemExit("bench.py: e2e produced non-finite lPz")
    e2e = {"value": M * a.steps / dt, "unit": UNIT, "h2d_bytes_per_step": int(M * d * 8 + cores.nbytes + xs.nbytes),
           "d2h_bytes_per_step": int(M * d * 8 + M * 8), "ms_per_step": 1e3 * dt / a.steps,
           "workload": "BASELINE.json configs[2]: M=2^%d, one GPU" % a.log2m,
           "api": "tt_irt1 (C-ABI symbol of tt_irt1_int32.so) on pinned host buffers; includes cores upload and marginalisation sweep"}
    # the same call as the reference's own Python caller makes it: ordinary pageable numpy arrays (tt_irt.py:44-51)
    qn = np.array(qh.numpy().T, order="F", copy=True)   # M x d, F-order, a pageable copy
    zn = np.zeros((M, d), order="F"); ln = np.zeros(M)

    def e2e_np():
        fn(c_int(d), n32.ctypes.data_as(ip), xs.ctypes.data_as(dp), r32.ctypes.data_as(ip), cores.ctypes.data_as(dp),
           c_int(M), qn.ctypes.data_as(dp), zn.ctypes.data_as(dp), ln.ctypes.data_as(dp))
    e2e_np()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        e2e_np()
    dtn = time.perf_counter() - t0
    e2e["pageable_numpy_value"] = M * a.steps / dtn
    e2e["pageable_numpy_note"] = "same call on pageable numpy arrays (bounce-buffer pipeline, %s host copy threads)" % os.environ.get("TTIRT_COPY_THREADS", "default")
    if not np.array_equal(ln[:4096], lh[:4096].numpy()):
        raise SystemExit("bench.py: pageable and pinned e2e results differ")
    return e2e


# ------------------------------------------------------------------------------------------------
# e2e, N GPUs: ONE call of the drop-in symbol, sharded over the N devices by the library (BASELINE configs[4])
# ------------------------------------------------------------------------------------------------
def e2e_sharded(a, lib, torch, ns, xs, rk, cores, d, ndev):
    import ctypes as C
    fn, it, ip, dp = _c_symbol(lib, 32)
    log2m = a.log2m_sharded
    qh = zh = lh = None
    while log2m >= 22:   # the host may not have 2 x 17 GB to page-lock: halve (and say so) rather than fail
        try:
            Ms = 1 << log2m
            qh = torch.empty((d, Ms), dtype=torch.float64, pin_memory=True)
            zh = torch.empty((d, Ms), dtype=torch.float64, pin_memory=True)
            lh = torch.empty((Ms,), dtype=torch.float64, pin_memory=True)
            break
        except RuntimeError:
            qh = zh = lh = None
            log2m -= 1
    if qh is None:
        return {"unavailable": "could not page-lock host buffers for the sharded call"}
    Ms = 1 << log2m
    gen = torch.Generator(device="cuda:0")
    gen.manual_seed(4321)
    step_rows = 1 << 22
    for m0 in range(0, Ms, step_rows):   # seeds made on GPU 0, parked in the pinned array
        blk = torch.rand((d, min(step_rows, Ms - m0)), dtype=torch.float64, device="cuda:0", generator=gen)
        qh[:, m0:m0 + blk.shape[1]].copy_(blk)
    del blk
    torch.cuda.empty_cache()
    n32 = np.ascontiguousarray(ns, dtype=it)
    r32 = np.ascontiguousarray(rk, dtype=it)
    os.environ["TTIRT_DEVICE"] = "0"
    os.environ["TTIRT_DEVICES"] = str(ndev)
    qp, zp, lp_ = C.cast(qh.data_ptr(), dp), C.cast(zh.data_ptr(), dp), C.cast(lh.data_ptr(), dp)

    def call():
        fn(C.c_int(d), n32.ctypes.data_as(ip), xs.ctypes.data_as(dp), r32.ctypes.data_as(ip), cores.ctypes.data_as(dp),
           C.c_int(Ms), qp, zp, lp_)

    def timed(f, steps, warm):
        for _ in range(warm):
            f()
        t0 = time.perf_counter()
        for _ in range(steps):
            f()
        return (time.perf_counter() - t0) / steps

    dt = timed(call, a.steps, min(a.warmup, 2))
    if not bool(torch.isfinite(lh[:: max(1, Ms // 4096)]).all()):
        raise SystemExit("bench.py: sharded e2e produced non-finite lPz")
    chk_rows = slice(Ms - 4096, Ms)    # rows of the last device's shard, compared with a one-device call on the same seeds below
    last = lh[chk_rows].clone()
    e2e = {"value": Ms / dt, "unit": UNIT, "h2d_bytes_per_step": int(Ms * d * 8 + cores.nbytes + xs.nbytes),
           "d2h_bytes_per_step": int(Ms * d * 8 + Ms * 8), "ms_per_step": 1e3 * dt,
           "workload": "BASELINE.json configs[4]: ONE tt_irt1 call, M=2^%d seed points sharded over %d B200 by the library (TTIRT_DEVICES=%d)" % (log2m, ndev, ndev),
           "api": "tt_irt1 (C-ABI symbol of tt_irt1_int32.so) called once by rank 0 on pinned host buffers; ttirt_run_host underneath: one host "
                  "thread per device, chunks of rows dealt out from a shared queue, cores uploaded to device 0 and fanned out by peer copies, no collective",
           "devices": ndev}
    if log2m != a.log2m_sharded:
        e2e["note"] = "host memory allowed only 2^%d seed points to be page-locked (asked for 2^%d)" % (log2m, a.log2m_sharded)
    # same seeds, last 2^20 rows on ONE device: the shard boundaries must not change a bit
    os.environ["TTIRT_DEVICES"] = "1"
    sub = 1 << 20
    q1 = qh[:, Ms - sub:].contiguous().pin_memory()
    z1 = torch.empty((d, sub), dtype=torch.float64, pin_memory=True); l1 = torch.empty((sub,), dtype=torch.float64, pin_memory=True)
    fn(C.c_int(d), n32.ctypes.data_as(ip), xs.ctypes.data_as(dp), r32.ctypes.data_as(ip), cores.ctypes.data_as(dp),
       C.c_int(sub), C.cast(q1.data_ptr(), dp), C.cast(z1.data_ptr(), dp), C.cast(l1.data_ptr(), dp))
    e2e["same_bits_as_one_device_on_last_rows"] = bool(torch.equal(l1[-4096:], last)) and bool(torch.equal(z1[:, -4096:], zh[:, chk_rows]))
    del q1, z1, l1
    os.environ["TTIRT_DEVICES"] = str(ndev)

    # ---- the same call without the q upload: seeds generated on the devices ----
    n64 = np.ascontiguousarray(ns, dtype=np.int64)
    r64 = np.ascontiguousarray(rk, dtype=np.int64)
    lp64 = C.POINTER(C.c_longlong)
    ru = lib.ttirt_run_uniform_host
    ru.restype = C.c_int
    ru.argtypes = [C.c_longlong, lp64, dp, lp64, dp, C.c_longlong, C.c_longlong, C.c_ulonglong, dp, dp, dp, C.c_int, C.c_int, C.c_int]

    def call_nu():
        rc = ru(d, n64.ctypes.data_as(lp64), xs.ctypes.data_as(dp), r64.ctypes.data_as(lp64), cores.ctypes.data_as(dp), Ms, 0, 2026,
                None, zp, lp_, 0, 0, ndev)
        if rc != 0:
            raise SystemExit("bench.py: ttirt_run_uniform_host failed")

    dtn = timed(call_nu, a.steps, 1)
    e2e["no_upload"] = {"value": Ms / dtn, "unit": UNIT, "ms_per_step": 1e3 * dtn, "h2d_bytes_per_step": int(cores.nbytes + xs.nbytes),
                        "d2h_bytes_per_step": int(Ms * d * 8 + Ms * 8),
                        "api": "ttirt_run_uniform_host: same sharded call, Philox seeds generated on the devices (no q upload), Z and lPz to pinned host buffers"}
    lib.ttirt_cache_clear()
    torch.cuda.empty_cache()

    # ---- the box's host-copy ceiling: plain pinned copies of the same shards on all devices at once ----
    try:
        e2e["host_copy_ceiling"] = host_copy_ceiling(torch, qh, zh, d, Ms, ndev)
        cg = e2e["host_copy_ceiling"]["duplex_GBps"]
        bps = 8 * (2 * d + 1)
        e2e["host_copy_ceiling"]["samples_per_s_at_ceiling"] = cg * 1e9 / bps
        e2e["frac_of_host_copy_ceiling"] = e2e["value"] * bps / (cg * 1e9)
        e2e["no_upload"]["frac_of_d2h_ceiling"] = e2e["no_upload"]["value"] * (8 * (d + 1)) / (e2e["host_copy_ceiling"]["d2h_only_GBps"] * 1e9)
    except Exception as ex:   # noqa: BLE001
        e2e["host_copy_ceiling"] = {"unavailable": "%s: %s" % (type(ex).__name__, str(ex)[:200])}
    return e2e


def host_copy_ceiling(torch, qh, zh, d, Ms, ndev):
    """GB/s of plain cudaMemcpy2DAsync between the pinned arrays of the sharded call and device buffers, every device at
    once, each copying its own row shard in chunks of 2^20 rows (d column segments per chunk, exactly the copies the
    pipeline issues): H2D alone, D2H alone and both directions together.  This is the ceiling of any end-to-end figure
    that streams q in and Z out on this box."""
    import ctypes as C
    rt = None
    for nm in ("libcudart.so.12", "libcudart.so"):
        try:
            rt = C.CDLL(nm)
            break
        except OSError:
            continue
    if rt is None:
        import glob as _g
        hits = _g.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so*"))
        rt = C.CDLL(hits[0])
    cp2d = rt.cudaMemcpy2DAsync
    cp2d.restype = C.c_int
    cp2d.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p]
    H2D, D2H = 1, 2
    chunk = 1 << 20
    bufs = []
    for g in range(ndev):
        with torch.cuda.device(g):
            bufs.append((torch.empty((d, chunk), dtype=torch.float64, device="cuda:%d" % g),
                         torch.empty((d, chunk), dtype=torch.float64, device="cuda:%d" % g),
                         [torch.cuda.Stream(device=g) for _ in range(2)], [torch.cuda.Stream(device=g) for _ in range(2)]))

    import itertools
    import threading
    rt.cudaStreamSynchronize.argtypes = [C.c_void_p]
    rt.cudaSetDevice.argtypes = [C.c_int]

    def run(h2d, d2h):
        """Chunks dealt out from one queue to one host thread per device (as ttirt_run_host does): a device whose link is
        faster copies more of them."""
        for g in range(ndev):
            torch.cuda.synchronize(g)
        counter = itertools.count()
        lock = threading.Lock()
        nchunks = (Ms + chunk - 1) // chunk
        err = []

        def worker(g):
            rt.cudaSetDevice(g)
            di, do, sa, sb = bufs[g]
            k = 0
            while True:
                # two chunks in flight per direction and device; a new chunk is claimed when the one before last is done
                s1, s2 = sa[k & 1], sb[k & 1]
                rt.cudaStreamSynchronize(s1.cuda_stream)
                rt.cudaStreamSynchronize(s2.cuda_stream)
                with lock:
                    i = next(counter)
                if i >= nchunks:
                    break
                m0 = i * chunk
                w = min(chunk, Ms - m0)
                if h2d and cp2d(di.data_ptr(), 8 * chunk, qh.data_ptr() + 8 * m0, 8 * Ms, 8 * w, d, H2D, s1.cuda_stream) != 0:
                    err.append("H2D")
                if d2h and cp2d(zh.data_ptr() + 8 * m0, 8 * Ms, do.data_ptr(), 8 * chunk, 8 * w, d, D2H, s2.cuda_stream) != 0:
                    err.append("D2H")
                k += 1
            for st in sa + sb:
                rt.cudaStreamSynchronize(st.cuda_stream)
        t0 = time.perf_counter()
        th = [threading.Thread(target=worker, args=(g,)) for g in range(ndev)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        dt = time.perf_counter() - t0
        if err:
            raise RuntimeError("cudaMemcpy2DAsync failed: %s" % err[0])
        return (int(h2d) + int(d2h)) * Ms * d * 8 / dt / 1e9

    run(True, True)
    out = {"h2d_only_GBps": run(True, False), "d2h_only_GBps": run(False, True), "duplex_GBps": run(True, True),
           "how": "cudaMemcpy2DAsync, pinned host arrays of the sharded call <-> device buffers, %d devices at once, chunks of 2^20 rows x %d column "
                  "segments dealt out from one queue to one host thread per device, one stream per direction and device" % (ndev, d)}
    return out


# ------------------------------------------------------------------------------------------------
# the other BASELINE shapes (N=1 only): device-resident and through the C symbol, a few steps each
# ------------------------------------------------------------------------------------------------
def other_configs(torch, peak_tflops):
    """BASELINE.json configs[3], [1] and [0] (parity-test shapes, not the metric): samples/s device-resident
    (ttirt_sample_device, CUDA events) and through the drop-in C symbol on pinned host buffers, with the two rooflines that
    could bound them (FP64 pipe: flops per sample x rate / measured DMMA peak; HBM: 8(2d+1) bytes per sample x rate /
    measured copy bandwidth).  Information for the record; failures here never touch the headline line."""
    import ctypes as C
    from tt_irt_py import tt_irt
    out = {}
    cases = [("configs[3] lorenz-40 shape, int64 ABI", 40, 33, 32, 22, 64, (-3.0, 3.0), 3, 3),
             ("configs[1] inverse-diffusion shape", 11, 17, 16, 20, 32, (-3.0 ** 0.5, 3.0 ** 0.5), 10, 3),
             ("configs[0] shock-absorber shape", 8, 17, 8, 14, 32, (0.0, 1.0), 50, 5),
             # not a BASELINE config: a shape beyond the fused kernel (r > 64, n > 72), served by the unfused DMMA path of
             # csrc/ttirt_wide.cu (the reference handles any shape on one code path, tt_irt1_int32.c:41-53)
             ("wide shape d=8 n=129 r=128 (beyond the fused kernel)", 8, 129, 128, 18, 32, (-1.0, 1.0), 3, 2)]
    for name, d, n, r, log2m, width, (lo, hi), steps, warm in cases:
        try:
            M = 1 << log2m
            ns, xs, rk, cores = synth.make_tt(d, n, r, seed=77, lo=lo, hi=hi)
            Wf = synth.flops_per_sample(ns, rk)
            md = tt_irt.Model(ns, xs, rk, cores, device=0)
            q = torch.rand((d, M), dtype=torch.float64, device="cuda:0")
            z = torch.empty_like(q); l = torch.empty((M,), dtype=torch.float64, device="cuda:0")
            st = torch.cuda.current_stream()

            def step():
                md.sample_device(M, q.data_ptr(), M, z.data_ptr(), M, l.data_ptr(), None, tt_irt.MODE_FAST, st.cuda_stream)
            for _ in range(warm):
                step()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            l0 = tt_irt.kernel_launches()
            e0.record(st)
            for _ in range(steps):
                step()
            e1.record(st)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            launches = (tt_irt.kernel_launches() - l0) // steps
            val = M / (ms * 1e-3)
            so = os.path.join(ROOT, "tt-irt_b200", "tt_irt_py", "tt_irt1_int32.so") if width == 32 else os.path.join(ROOT, "tt-irt_b200", "lib", "libtt_irt1_int64.so")
            lw = tt_irt.load_library() if width == 32 else C.CDLL(so)
            fn, it, ip, dp = _c_symbol(lw, width)
            ct = C.c_int if width == 32 else C.c_longlong
            qh = torch.empty((d, M), dtype=torch.float64, pin_memory=True); qh.copy_(q)
            zh = torch.empty((d, M), dtype=torch.float64, pin_memory=True); lh = torch.empty((M,), dtype=torch.float64, pin_memory=True)
            nn, rr = np.ascontiguousarray(ns, dtype=it), np.ascontiguousarray(rk, dtype=it)
            os.environ["TTIRT_DEVICE"] = "0"; os.environ["TTIRT_DEVICES"] = "1"

            cores_b = cores.copy()
            cores_b[-1] = np.nextafter(cores_b[-1], 2.0)   # a second TT, one ulp away in one entry: forces upload + sweep

            def call(cc=cores):
                fn(ct(d), nn.ctypes.data_as(ip), xs.ctypes.data_as(dp), rr.ctypes.data_as(ip), cc.ctypes.data_as(dp), ct(M),
                   C.cast(qh.data_ptr(), dp), C.cast(zh.data_ptr(), dp), C.cast(lh.data_ptr(), dp))
            # (a) a different TT on every call: grid + cores upload, sweep and operand packing inside every call
            for i in range(warm):
                call(cores_b if i & 1 else cores)
            t0 = time.perf_counter()
            for i in range(steps):
                call(cores_b if i & 1 else cores)
            dt = (time.perf_counter() - t0) / steps
            # (b) the same TT again and again (the MH / IW drivers' pattern): models up to 512 KB are recognised by an exact
            #     byte comparison and stay resident
            call(); call()
            t0 = time.perf_counter()
            for _ in range(steps):
                call()
            dts = (time.perf_counter() - t0) / steps
            ok = bool(torch.isfinite(lh).all()) and bool(torch.equal(lh, l.cpu()))
            out[name] = {"d": d, "n": n, "r": r, "M": M, "abi_width": width, "value": val, "ms_per_step": ms, "launches_per_step": int(launches),
                         "e2e_value": M / dt, "e2e_ms_per_call": 1e3 * dt, "e2e_ms_per_call_same_tt_again": 1e3 * dts,
                         "e2e_matches_device_resident_bit_for_bit": ok,
                         "flops_per_sample": Wf, "frac_of_fp64_peak": val * Wf / 1e12 / peak_tflops,
                         "frac_of_hbm_stream_bound": val * 8 * (2 * d + 1) / 1e9 / _hbm_peak(),
                         "binding_roofline": "fp64" if Wf / (8.0 * (2 * d + 1)) > peak_tflops * 1e3 / _hbm_peak() else "hbm"}
            md.close()
            del q, z, l, qh, zh, lh
            lw.ttirt_cache_clear()
            torch.cuda.empty_cache()
        except Exception as ex:   # noqa: BLE001
            out[name] = {"unavailable": "%s: %s" % (type(ex).__name__, str(ex)[:200])}
    return out


def _next_rows():
    """Information only (not part of the metric): the SURVEY section 8(f) rank-4 row, the squared-density transform
    tt_irt_sqr, measured by its own script at its metric shape in a child process (device-resident, no CPU leg, so
    nothing under oracle/ runs).  A failure here never touches the headline line."""
    try:
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "devtools", "bench_sqr.py"), "--no-cpu", "--no-e2e",
                              "--steps", "2", "--warmup", "2"], capture_output=True, text=True, timeout=180)
        j = json.loads(out.stdout.strip().splitlines()[-1])
        return {"tt_irt_sqr": {"metric": j["metric"], "value": j["value"], "unit": j["unit"], "workload": j["config"]["workload"],
                               "roofline": {k: j["roofline"][k] for k in ("bound", "achieved", "peak", "unit", "frac", "kernel", "kernel_share_of_step")},
                               "algorithmic_tflops_whole_step": j["algorithmic_tflops_whole_step"], "gpu_launches": j["gpu_launches"]}}
    except Exception as e:   # noqa: BLE001
        return {"tt_irt_sqr": {"unavailable": "%s: %s" % (type(e).__name__, str(e)[:200])}}


def _ncu_traffic(M):
    """DRAM bytes per launch of the dominant kernel from the newest committed `ncu --set full` capture of this kernel
    (profiles/rNN_ncu_traffic.json, written from the round's own capture by tools/ncu_traffic.py; same rows per launch), else None."""
    best = (None, None)
    for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_traffic.json"))):
        try:
            with open(f) as fh:
                j = json.load(fh)
            if "transition_kernel" in j.get("kernel", "") and min(M, 1 << 22) == int(j["rows_per_launch"]):
                best = (int(j["dram_bytes_read"]) + int(j["dram_bytes_write"]), os.path.relpath(f, ROOT))
        except Exception:
            continue
    return best


def _hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"])
    except Exception:
        return 6650.0


if __name__ == "__main__":
    sys.exit(main())
