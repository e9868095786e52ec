"""Seeded synthetic tensor-train densities and seed points for the BASELINE.json configs.

Shapes follow SURVEY.md section 8(d): cores i.i.d. U(0,1) (a positive, well-posed density),
ranks (1, r, ..., r, 1), a uniform or Chebyshev-like grid per dimension, q i.i.d. U(0,1).
Storage is the reference's "TT2.0" contiguous layout: core k is column-major
r_k x n_k x r_{k+1} at offset sum_{j<k} r_j n_j r_{j+1} (tt_irt1_int32.c:55-57).
"""
import numpy as np

CONFIGS = {
    # name: (d, n, r, M, grid_lo, grid_hi, abi width)
    "shock_d8_n17_r8": (8, 17, 8, 2 ** 14, 0.0, 1.0, 32),       # BASELINE configs[0] shape (D=6 -> d=8)
    "diffusion_d11_n17_r16": (11, 17, 16, 2 ** 20, -3.0 ** 0.5, 3.0 ** 0.5, 32),  # configs[1]
    "roofline_d32_n65_r64": (32, 65, 64, 2 ** 24, -1.0, 1.0, 32),  # configs[2], metric config
    "lorenz_d40_n33_r32": (40, 33, 32, 2 ** 22, -3.0, 3.0, 64),    # configs[3], int64 ABI
}


def make_tt(d, n, r, seed=0, lo=-1.0, hi=1.0, grid="uniform", cores="uniform", ranks=None, ns=None):
    """Return (n (d,), xs (sum n,), ranks (d+1,), cores (sum r n r,)) as numpy arrays."""
    rng = np.random.default_rng(seed)
    ns = np.full(d, n, dtype=np.int64) if ns is None else np.asarray(ns, dtype=np.int64)
    if ranks is None:
        ranks = np.array([1] + [r] * (d - 1) + [1], dtype=np.int64)
    ranks = np.asarray(ranks, dtype=np.int64)
    xs = []
    for k in range(d):
        nk = int(ns[k])
        if grid == "uniform":
            xs.append(lo + (hi - lo) * np.arange(nk) / (nk - 1))
        elif grid == "chebyshev":  # cf. tt_dirt_approx.m:306
            xs.append(lo + (hi - lo) * 0.5 * (np.cos(np.pi * (nk - 1 - np.arange(nk)) / (nk - 1)) + 1.0))
        else:
            raise ValueError(grid)
    xs = np.concatenate(xs)
    size = int((ranks[:-1] * ns * ranks[1:]).sum())
    if cores == "uniform":
        c = rng.random(size)
    elif cores == "normal":  # exercises the fabs() at tt_irt1_int32.c:105
        c = rng.standard_normal(size)
    else:
        raise ValueError(cores)
    return ns, xs, ranks, c


def make_q(M, d, seed=1):
    """Seed points q, (M, d) Fortran-ordered float64, as test_shock_absorber_tt.py:147-148 passes them."""
    rng = np.random.default_rng(seed)
    return np.asfortranarray(rng.random((d, M)).T)


def flops_per_sample(ns, ranks):
    """Algorithmic FP64 flops per sample, SURVEY.md section 8(d): sum 2 r_k n_k + sum_{k<d-1} 4 r_k r_{k+1}."""
    ns = np.asarray(ns, dtype=np.int64)
    ranks = np.asarray(ranks, dtype=np.int64)
    d = ns.size
    return int((2 * ranks[:d] * ns).sum() + (4 * ranks[:d - 1] * ranks[1:d]).sum())
