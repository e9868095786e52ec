#!/usr/bin/env python
"""bench.py -- IRT samples/s of the B200-native tt_irt1 on BASELINE.json's metric configuration.

    python bench.py [--gpus N] [--steps K] [--warmup W]          (N>1: launched by torch.distributed.run)
    python bench.py --impl reference [...]                        the reference's own CPU path, host cores

A "step" is one pass of the hot path (all d dimensions) over one batch of M synthetic seed points.
Workload: BASELINE.json configs[2], synthetic random TT density d=32, n=65, r=64, M=2^24 uniform points
per GPU (weak scaling: N=4 is configs[4]'s M=2^26).  `value` times the device-resident entry point
(ttirt_sample_device, inputs already in HBM) with CUDA events on the launching stream; `e2e` times the
reference's own C-ABI symbol `tt_irt1` on pinned HOST buffers (cores upload, marginalisation sweep, H2D of q,
kernels, D2H of Z and lPz all inside the timed region).  One JSON line on stdout (rank 0).
"""
import argparse
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
    ap.add_argument("--shape", default="32,65,64", help="d,n,r (default: the metric configuration)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-next-rows", action="store_true", help="skip the informational tt_irt_sqr figure appended at N=1")
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
    _CPU_CTX["args"] = (kind, 32, blas, ns, xs, rk, cores)
    _CPU_CTX["qs"] = [synth.make_q(samples_per_core, d, seed=1000 + i) for i in range(ncores)]
    ctx = mp.get_context("fork")
    best = None
    with ctx.Pool(ncores) as pool:
        pool.map(_cpu_worker_warm, range(ncores))  # library load + first touch outside the timed region
        for _ in range(repeats):
            t0 = time.perf_counter()
            pool.map(_cpu_worker, range(ncores), chunksize=1)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    total = samples_per_core * ncores
    return {"value": total / best, "unit": UNIT, "cores": ncores, "kind": kind,
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


def bind_to_gpu_numa_node(gpu_index):
    """Pin this rank (and, by first touch, its page-locked buffers) to the CPUs next to its GPU: with 8 ranks streaming
    q / Z through host memory at once, remote-socket traffic is what limits the end-to-end figure."""
    try:
        out = subprocess.run(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=20).stdout.strip()
        bdf = out.splitlines()[0].strip()
        bdf = bdf.lower()
        if bdf.startswith("00000000:"):
            bdf = bdf[4:]
        path = "/sys/bus/pci/devices/%s/local_cpulist" % bdf
        with open(path) as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            if "-" in part:
                lo, hi = part.split("-"); cpus.update(range(int(lo), int(hi) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return "cpus %s (local to GPU %d at %s)" % (spec, gpu_index, bdf)
    except Exception as e:  # no sysfs / no permission: run unbound
        return "unbound (%s)" % type(e).__name__
    return "unbound"


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
    binding = bind_to_gpu_numa_node(local_rank) if world > 1 else "unbound (single rank)"
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
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
    md.profile_enable(True)
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
    k_ms, k_launches, k_flops = md.profile_read()
    md.profile_enable(False)
    tmax = torch.tensor([ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_all = float(tmax.item())
    value = world * M * a.steps / (ms_all * 1e-3)
    checksum = float(lpz[: 1 << 10].sum().item())  # D2H read of a result so nothing is elided

    # ---- e2e: the reference's own symbol tt_irt1 on pinned host buffers ---------------------------
    e2e = None
    if not a.no_e2e:
        from ctypes import POINTER, c_double, c_int, cast
        del z, lpz
        qh = torch.empty((d, M), dtype=torch.float64, pin_memory=True)
        qh.copy_(q)
        del q
        torch.cuda.empty_cache()
        zh = torch.empty((d, M), dtype=torch.float64, pin_memory=True)
        lh = torch.empty((M,), dtype=torch.float64, pin_memory=True)
        n32 = np.ascontiguousarray(ns, dtype=np.int32)
        r32 = np.ascontiguousarray(rk, dtype=np.int32)
        dp, ip = POINTER(c_double), POINTER(c_int)
        os.environ["TTIRT_DEVICE"] = str(local_rank)
        os.environ["TTIRT_DEVICES"] = "1"

        def e2e_step():
            lib.tt_irt1(c_int(d), n32.ctypes.data_as(ip), xs.ctypes.data_as(dp), r32.ctypes.data_as(ip), cores.ctypes.data_as(dp),
                        c_int(M), cast(qh.data_ptr(), dp), cast(zh.data_ptr(), dp), cast(lh.data_ptr(), dp))

        for _ in range(min(a.warmup, 2)):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            e2e_step()
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        if not bool(torch.isfinite(lh[: 1 << 12]).all()):
            raise SystemExit("bench.py: e2e produced non-finite lPz")
        e2e = {"value": world * M * a.steps / dt, "unit": UNIT, "h2d_bytes_per_step": int(world * (M * d * 8 + cores.nbytes + xs.nbytes)),
               "d2h_bytes_per_step": int(world * (M * d * 8 + M * 8)), "ms_per_step": 1e3 * dt / a.steps,
               "api": "tt_irt1 (C-ABI symbol of tt_irt1_int32.so) on pinned host buffers; includes cores upload and marginalisation sweep"}
        # the same call as the reference's own Python caller makes it: ordinary pageable numpy arrays (tt_irt.py:44-51)
        if world == 1:
            qn = np.array(qh.numpy().T, order="F", copy=True)   # M x d, F-order, a pageable copy
            zn = np.zeros((M, d), order="F"); ln = np.zeros(M)

            def e2e_np():
                lib.tt_irt1(c_int(d), n32.ctypes.data_as(ip), xs.ctypes.data_as(dp), r32.ctypes.data_as(ip), cores.ctypes.data_as(dp),
                            c_int(M), qn.ctypes.data_as(dp), zn.ctypes.data_as(dp), ln.ctypes.data_as(dp))
            e2e_np()
            t0 = time.perf_counter()
            for _ in range(a.steps):
                e2e_np()
            dtn = time.perf_counter() - t0
            e2e["pageable_numpy_value"] = M * a.steps / dtn
            e2e["pageable_numpy_note"] = "same call on pageable numpy arrays (bounce-buffer pipeline, %s host copy threads)" % os.environ.get("TTIRT_COPY_THREADS", "4")
            if not np.array_equal(ln[:4096], lh[:4096].numpy()):
                raise SystemExit("bench.py: pageable and pinned e2e results differ")
            del qn, zn, ln

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    peak, peak_src = measured_fp64_peak()
    ach = (k_flops / (k_ms * 1e-3)) / 1e12 if k_ms > 0 else 0.0
    roofline = {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak if peak else None,
                "traffic": _ncu_traffic(M), "traffic_unit": "bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)",
                "kernel": "ttirt::transition_kernel<RT,NT,EXACT,TAIL1> (FP64 DMMA mma.sync.m8n8k4; 8 MMA warps + 4 tail warps per CTA)",
                "kernel_launches_timed": k_launches, "kernel_avg_ms": k_ms / max(1, k_launches),
                "kernel_share_of_step": k_ms / ms if ms > 0 else None,
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
            "gpu_launches": int(launches), "clocks": clocks, "checksum_lpz_1k": checksum, "host_binding": binding}
    if world == 1 and not a.no_next_rows:
        line["next_rows"] = _next_rows()
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


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
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (same chunk size), else None."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")) as f:
            j = json.load(f)
        if min(M, 1 << 20) == int(j["rows_per_launch"]):
            return int(j["dram_bytes_read"]) + int(j["dram_bytes_write"])
    except Exception:
        pass
    return None


def _hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"])
    except Exception:
        return 6650.0


if __name__ == "__main__":
    sys.exit(main())
