"""Generates tests/golden/*.npz by running the UNMODIFIED reference tt_irt1 (compiled from
/root/reference into oracle/_ref by oracle/Makefile) on seeded inputs.  Run in the build container:

    make -C oracle && python tests/golden/make_golden.py

Each fixture stores the full inputs when they are small (cores <= 1 MB) or, for the BASELINE shapes,
the seeds plus a sha256 of the regenerated cores; and the reference outputs for both BLAS builds:
  Z_shim, lPz_shim         reference source + netlib-order BLAS shim (deterministic, self-contained)
  Z_openblas, lPz_openblas reference source + the wheel-bundled OpenBLAS (single thread)
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tt-irt_b200"))
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")

import oracle  # noqa: E402
from tt_irt_py import synth  # noqa: E402

CASES = [
    # name, d, n, r, M, width, cores, grid, (lo, hi)
    ("tiny_d3_n5_r3", 3, 5, 3, 64, 32, "uniform", "uniform", (-1.0, 1.0)),
    ("shock_d8_n17_r8", 8, 17, 8, 256, 32, "uniform", "uniform", (0.0, 1.0)),
    ("shock_d8_n17_r16_cheb", 8, 17, 16, 256, 32, "uniform", "chebyshev", (-1.0, 1.0)),
    ("diffusion_d11_n17_r16", 11, 17, 16, 256, 32, "uniform", "uniform", (-3.0 ** 0.5, 3.0 ** 0.5)),
    ("signed_d6_n20_r12", 6, 20, 12, 128, 32, "normal", "uniform", (-1.0, 1.0)),
    ("lorenz_d40_n33_r32_i64", 40, 33, 32, 128, 64, "uniform", "uniform", (-3.0, 3.0)),
    ("roofline_d32_n65_r64", 32, 65, 64, 128, 32, "uniform", "uniform", (-1.0, 1.0)),
]


def main():
    for name, d, n, r, M, width, cores, grid, (lo, hi) in CASES:
        seed = abs(hash(name)) % 1000 if False else sum(map(ord, name))
        ns, xs, rk, c = synth.make_tt(d, n, r, seed=seed, lo=lo, hi=hi, grid=grid, cores=cores)
        q = synth.make_q(M, d, seed=seed + 1)
        out = {"n": ns, "xs": xs, "ranks": rk, "q": q, "width": width, "seed": seed,
               "cores_kind": cores, "grid_kind": grid, "lo": lo, "hi": hi,
               "cores_sha256": hashlib.sha256(c.tobytes()).hexdigest()}
        if c.nbytes <= (1 << 20):
            out["cores"] = c
        for blas in ("shim", "openblas"):
            if oracle.have_ref(width, blas):
                Z, l = oracle.ref_run(ns, xs, rk, c, q, width=width, blas=blas)
                out["Z_" + blas] = Z
                out["lPz_" + blas] = l
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, {k: getattr(v, "shape", v) for k, v in out.items() if k.startswith(("Z_", "lPz_"))})


if __name__ == "__main__":
    main()
