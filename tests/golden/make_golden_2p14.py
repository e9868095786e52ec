"""Generates tests/golden/*_2p14.npz: the UNMODIFIED reference tt_irt1 (compiled from /root/reference into oracle/_ref
by oracle/Makefile, both BLAS builds) on M = 2^14 seeded seed points for the five BASELINE shapes.

    make -C oracle && python tests/golden/make_golden_2p14.py

Full outputs at this size would be ~25 MB of incompressible doubles, so each fixture stores
  * the generating seeds and sha256 of the regenerated cores and q (drift guards),
  * sha256 of the netlib-order build's complete Z and lPz: the oracle and the strict GPU mode must reproduce every bit,
  * every 16th row of both builds' outputs in full,
  * the reference's OWN noise when only its BLAS changes (OpenBLAS vs netlib order), over all 2^14 x d entries: fraction of
    Z entries beyond 1e-12 relative, largest |dZ|, largest relative dlPz, interval flips.  The fast GPU path is held to
    1.5x that fraction on the same inputs (tests/test_parity_2p14.py).
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

M = 1 << 14
CASES = [
    # name, d, n, r, width, cores, grid, (lo, hi)
    ("shock_d8_n17_r8", 8, 17, 8, 32, "uniform", "uniform", (0.0, 1.0)),                       # configs[0]
    ("shock_d8_n17_r16_cheb", 8, 17, 16, 32, "uniform", "chebyshev", (-1.0, 1.0)),             # configs[0], upper rank, stress grid
    ("diffusion_d11_n17_r16", 11, 17, 16, 32, "uniform", "uniform", (-3.0 ** 0.5, 3.0 ** 0.5)),  # configs[1]
    ("lorenz_d40_n33_r32_i64", 40, 33, 32, 64, "uniform", "uniform", (-3.0, 3.0)),             # configs[3]
    ("roofline_d32_n65_r64", 32, 65, 64, 32, "uniform", "uniform", (-1.0, 1.0)),               # configs[2] / [4]
]


def sha(a, order="F"):
    return hashlib.sha256(np.asarray(a).tobytes(order=order)).hexdigest()


def intervals(Z, ns, xs):
    off = np.concatenate([[0], np.cumsum(ns)])
    idx = np.zeros(Z.shape, dtype=np.int32)
    for k in range(ns.size):
        x = xs[off[k]:off[k + 1]]
        idx[:, k] = np.clip(np.searchsorted(x, Z[:, k], side="right") - 1, 0, ns[k] - 2)
    return idx


def main():
    for name, d, n, r, width, cores, grid, (lo, hi) in CASES:
        seed = sum(map(ord, name)) + 14
        ns, xs, rk, c = synth.make_tt(d, n, r, seed=seed, lo=lo, hi=hi, grid=grid, cores=cores)
        q = synth.make_q(M, d, seed=seed + 1)
        Zs, ls = oracle.ref_run(ns, xs, rk, c, q, width=width, blas="shim")
        Zb, lb = oracle.ref_run(ns, xs, rk, c, q, width=width, blas="openblas")
        rows = np.arange(0, M, 16)
        dz = np.abs(Zb - Zs)
        rel = dz > 1e-12 * np.maximum(1.0, np.abs(Zs))
        out = {"n": ns, "xs": xs, "ranks": rk, "width": width, "seed": seed, "M": M, "cores_kind": cores, "grid_kind": grid, "lo": lo, "hi": hi,
               "cores_sha256": sha(c, "C"), "q_sha256": sha(q),
               "Z_shim_sha256": sha(Zs), "lPz_shim_sha256": sha(ls, "C"), "rows": rows,
               "Z_shim_rows": Zs[rows], "lPz_shim_rows": ls[rows], "Z_openblas_rows": Zb[rows], "lPz_openblas_rows": lb[rows],
               "ref_blas_swap_z_frac_gt_1e-12": float(rel.mean()), "ref_blas_swap_z_max_abs": float(dz.max()),
               "ref_blas_swap_lpz_max_rel": float((np.abs(lb - ls) / np.maximum(1.0, np.abs(ls))).max()),
               "ref_blas_swap_interval_flips": int((intervals(Zb, ns, xs) != intervals(Zs, ns, xs)).sum())}
        np.savez_compressed(os.path.join(HERE, name + "_2p14.npz"), **out)
        print(name, {k: v for k, v in out.items() if k.startswith("ref_blas_swap")})


if __name__ == "__main__":
    main()
