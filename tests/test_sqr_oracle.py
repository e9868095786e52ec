"""CPU tests of the squared-density oracle (oracle/tt_irt_sqr_oracle.py, the numpy restatement of the Matlab-only
reference matlab/samplers/tt_irt_sqr.m).  Matlab cannot run here; tests/test_matlab_pins.py pins the oracle against the reference's
source executed by a small Matlab-subset interpreter.  This file pins it independently of any interpreter: against the
unmodified reference C tt_irt1 where the two transforms coincide, by closed forms and by properties the reference's
construction implies; the committed tests/golden/sqr_*.npz freeze it."""
import hashlib
import importlib.util
import os

import numpy as np
import pytest

from oracle.tt_irt_sqr_oracle import sqr_sweep, tt_irt_sqr_oracle, tt_rt_sqr_oracle, split_cores, tracemult_oracle
from tt_irt_py import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

_spec = importlib.util.spec_from_file_location("make_golden_sqr", os.path.join(GOLDEN, "make_golden_sqr.py"))
mk = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(mk)


def _inv1(x, p, u):
    """Independent scalar inverse CDF of the piecewise-linear density with node values p (stable root form)."""
    h = np.diff(x)
    C = np.concatenate([[0.0], np.cumsum(0.5 * (p[:-1] + p[1:]) * h)])
    p = p / C[-1]
    C = C / C[-1]
    z, dens = np.zeros_like(u), np.zeros_like(u)
    for m, v in enumerate(u):
        i = min(max(int(np.searchsorted(C, v, side="left")) - 1, 0), x.size - 2)
        a = 0.5 * (p[i + 1] - p[i]) / h[i]
        t = 2.0 * (v - C[i]) / (p[i] + np.sqrt(p[i] ** 2 + 4.0 * a * (v - C[i])))
        z[m] = x[i] + t
        dens[m] = p[i] + (p[i + 1] - p[i]) * t / h[i]
    return z, dens


def test_rank_one_density_factorises_into_scalar_inverse_cdfs():
    """sqrt-density of rank 1: the transform is d independent 1-D inverse CDFs of the squared node values."""
    d, n = 3, 9
    rng = np.random.default_rng(0)
    vs = [rng.random(n) + 0.1 for _ in range(d)]
    grids = [np.linspace(-1, 1, n) ** 3 for _ in range(d)]
    q = synth.make_q(400, d, seed=5)
    Z, lF = tt_irt_sqr_oracle([n] * d, np.concatenate(grids), [1] * (d + 1), np.concatenate(vs), q)
    ref = np.zeros(400)
    for k in range(d):
        z, dens = _inv1(grids[k], vs[k] ** 2, q[:, k])
        np.testing.assert_allclose(Z[:, k], z, rtol=0, atol=5e-14)
        ref += np.log(dens)
    np.testing.assert_allclose(lF, ref, rtol=0, atol=2e-13)


def test_semi_marginals_are_squares_and_the_first_one_integrates_the_density():
    """P{k}(:, j) of tt_irt_sqr.m:80 is a Gram matrix (symmetric positive semi-definite), and P{1} equals the exact
    marginal of the squared linear interpolant under the trapezoid rule the reference uses (:49-51)."""
    d, n, r = 4, 7, 3
    ns, xs, rk, c = synth.make_tt(d, n, r, seed=2, cores="normal")
    sw = sqr_sweep(ns, xs, rk, c)
    for k in range(d):
        G = sw["P"][k].reshape((int(rk[k]), int(rk[k]), n), order="F")
        np.testing.assert_allclose(G, np.transpose(G, (1, 0, 2)), rtol=0, atol=1e-12 * np.abs(G).max())
        for j in range(n):
            assert np.linalg.eigvalsh(G[:, :, j]).min() > -1e-10 * np.abs(G).max()
    cores = split_cores(ns, rk, c)
    full = cores[0][0]                       # (n, r)
    for k in range(1, d):
        full = np.tensordot(full, cores[k], axes=([-1], [0]))
    full = full[..., 0] ** 2                 # squared node values of the sqrt-density, shape (n,)*d
    for k in range(d - 1, 0, -1):
        x = xs[k * n:(k + 1) * n]
        h = np.diff(x)
        w = np.concatenate([[h[0]], h[:-1] + h[1:], [h[-1]]]) * 0.5
        full = full @ w
    np.testing.assert_allclose(sw["P"][0][0], full, rtol=1e-11)


def test_samples_invert_the_conditional_cdf():
    """For every sample and dimension the piecewise-quadratic CDF of the conditional, evaluated at x_k, returns q_k."""
    d, n, r = 4, 9, 4
    ns, xs, rk, c = synth.make_tt(d, n, r, seed=8)
    q = synth.make_q(300, d, seed=9)
    Z, lF, idx, cond, gap, ls = tt_irt_sqr_oracle(ns, xs, rk, c, q, extras=True)
    sw = sqr_sweep(ns, xs, rk, c)
    f = np.ones((300, 1))
    lsum = np.zeros(300)
    for k in range(d):
        x = xs[k * n:(k + 1) * n]
        r0 = int(rk[k])
        G = sw["P"][k].reshape((r0, r0, n), order="F")
        p = np.einsum("ma,abj,mb->mj", f, G, f)
        h = np.diff(x)
        C = np.concatenate([np.zeros((300, 1)), np.cumsum(0.5 * (p[:, :-1] + p[:, 1:]) * h, axis=1)], axis=1)
        p = p / C[:, -1:]
        C = C / C[:, -1:]
        i0 = idx[:, k]
        ar = np.arange(300)
        t = Z[:, k] - x[i0]
        assert (t >= 0).all() and (t <= h[i0] * (1 + 1e-14)).all()
        slope = (p[ar, i0 + 1] - p[ar, i0]) / h[i0]
        cdf_at = C[ar, i0] + p[ar, i0] * t + 0.5 * slope * t * t
        np.testing.assert_allclose(cdf_at, q[:, k], rtol=0, atol=1e-11)
        lsum += np.log(p[ar, i0] + slope * t)
        if k < d - 1:
            core = sw["f"][k]
            w2 = t / h[i0]
            f = np.einsum("ma,amb->mb", f, core[:, i0, :]) * (1 - w2)[:, None] + np.einsum("ma,amb->mb", f, core[:, i0 + 1, :]) * w2[:, None]
    np.testing.assert_allclose(lF, lsum, rtol=0, atol=1e-10)


def test_marginal_sampling_is_a_prefix_of_the_full_transform():
    """q with D < d columns samples the marginal of the first D variables (tt_irt_sqr.m:9, :105)."""
    ns, xs, rk, c = synth.make_tt(5, 9, 4, seed=3)
    q = synth.make_q(200, 5, seed=4)
    Zf, lf = tt_irt_sqr_oracle(ns, xs, rk, c, q)
    Zm, lm = tt_irt_sqr_oracle(ns, xs, rk, c, q[:, :3])
    np.testing.assert_array_equal(Zm, Zf[:, :3])
    assert np.isfinite(lm).all() and not np.allclose(lm, lf)


def test_boundary_extension_equals_explicitly_extended_cores():
    """Cores without boundary nodes are extrapolated linearly (tt_irt_sqr.m:53-60); handing the extrapolated cores
    over with the same grid must give the same transform."""
    d, n, r = 4, 8, 3
    ns, xs, rk, c = mk.make_case(d, n, r, 21, -1.0, 1.0, "uniform", "uniform", True)
    q = synth.make_q(150, d, seed=6)
    Z1, l1 = tt_irt_sqr_oracle(ns, xs, rk, c, q)
    sw = sqr_sweep(ns, xs, rk, c)
    cext = np.concatenate([f.ravel(order="F") for f in sw["f"]])
    Z2, l2 = tt_irt_sqr_oracle(sw["n"], xs, rk, cext, q)
    np.testing.assert_array_equal(Z1, Z2)
    np.testing.assert_array_equal(l1, l2)
    with pytest.raises(ValueError):
        tt_irt_sqr_oracle(ns, xs[:-1], rk, c, q)


def test_zero_mass_conditional_falls_back_to_the_grid_density():
    """tt_irt_sqr.m:121-127: a conditional without mass is replaced by h (and its CDF by cumsum(h))."""
    n = 6
    x = np.linspace(0.0, 2.0, n)
    Z, lF = tt_irt_sqr_oracle([n], x, [1, 1], np.zeros(n), np.array([[0.1], [0.5], [0.95]]))
    assert np.isfinite(Z).all() and (Z >= x[0]).all() and (Z <= x[-1]).all()
    assert (np.diff(Z[:, 0]) > 0).all()


GOLD = ["sqr_tiny_d3_n5_r3", "sqr_shock_d8_n17_r8", "sqr_dirt_d6_n17_r16_cheb", "sqr_noboundary_d5_n15_r6",
        "sqr_marginal_d6_n33_r32", "sqr_roofline_d4_n65_r64"]


def load_sqr_golden(name):
    g = dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))
    ns, xs, rk, c = mk.make_case(int(g["d"]), int(g["n"]), int(g["r"]), int(g["seed"]), float(g["lo"]), float(g["hi"]),
                                 str(g["grid_kind"]), str(g["cores_kind"]), bool(g["ext"]))
    q = synth.make_q(int(g["M"]), int(g["D"]), seed=int(g["seed"]) + 1)
    if hashlib.sha256(c.tobytes()).hexdigest() != str(g["cores_sha256"]) or hashlib.sha256(q.tobytes()).hexdigest() != str(g["q_sha256"]):
        raise RuntimeError("regenerated inputs of golden %s do not match the recorded sha256 (numpy RNG drift)" % name)
    return g, ns, xs, rk, c, q


@pytest.mark.parametrize("name", GOLD)
def test_oracle_reproduces_its_committed_golden_outputs(name):
    g, ns, xs, rk, c, q = load_sqr_golden(name)
    Z, lF, idx, cond, gap, ls = tt_irt_sqr_oracle(ns, xs, rk, c, q, extras=True)
    np.testing.assert_array_equal(idx, g["idx"])
    # same numpy / OpenBLAS build reproduces bit for bit; a different BLAS may move entries by its summation order
    tol = 1e-12 * np.maximum(1.0, np.abs(g["xq"])) + 8 * np.finfo(float).eps * np.cumsum(g["cond"], axis=1)
    assert (np.abs(Z - g["xq"]) <= tol).all()
    np.testing.assert_allclose(lF, g["lFapp"], rtol=0, atol=1e-10)


def test_tracemult_oracle_matches_the_mex_loops():
    """matlab/utils/tracemult.c:103-112 (one dgemm per slice, B slice picked by the 1-based j) and :131-136."""
    rng = np.random.default_rng(3)
    p, m, k, n, s = 3, 4, 5, 7, 6
    A = rng.standard_normal((p, m, n))
    B = rng.standard_normal((m, k, s))
    j = rng.integers(1, s + 1, n).astype(float)
    C = tracemult_oracle(A, j, B)
    for i in range(n):
        for a in range(p):
            for b in range(k):
                assert abs(C[a, b, i] - sum(A[a, l, i] * B[l, b, int(j[i]) - 1] for l in range(m))) < 1e-13
    A2 = rng.standard_normal((n, s))
    np.testing.assert_array_equal(tracemult_oracle(A2, j), np.array([A2[i, int(j[i]) - 1] for i in range(n)]))
    # the two uses inside tt_irt_sqr.m: the Cartesian square of the interface (:109) and the slab product (:205)
    f = rng.standard_normal((4, 1, n))
    sq = tracemult_oracle(f, np.arange(1, n + 1), np.transpose(f, (1, 0, 2)))
    np.testing.assert_allclose(sq, np.einsum("ai,bi->abi", f[:, 0, :], f[:, 0, :]), rtol=0, atol=1e-15)


def _squared_tt(ns, rk, c):
    """TT of the elementwise square p = f o f: cores are the Kronecker squares of f's cores, ranks r^2."""
    cores = split_cores(ns, rk, c)
    out = []
    for G in cores:
        r0, n, r1 = G.shape
        K = np.einsum("ajb,cjd->acjbd", G, G).reshape((r0 * r0, n, r1 * r1), order="C")
        out.append(K.ravel(order="F"))
    return np.concatenate(out), np.asarray(rk, dtype=np.int64) ** 2


def _reference_tt_irt1(ns, xs, rk, c, q):
    """The unmodified reference C tt_irt1 (oracle/_ref) when it is built, else its bit-exact C restatement."""
    import oracle
    if oracle.have_ref(32, "shim"):
        return oracle.ref_run(ns, xs, rk, c, q, width=32, blas="shim")
    if not os.path.exists(os.path.join(oracle.ORACLE_DIR, "liboracle_tt_irt1.so")):
        oracle.build()
    return oracle.oracle_run(ns, xs, rk, c, q)


def test_pinned_by_the_reference_c_routine_on_separable_densities():
    """For a rank-1 sqrt-density the squared-density transform and the reference's linear-spline tt_irt1 on the squared node
    values are the same map (the scalar left interface normalises away): Z and the log-density agree with the live reference."""
    d, n = 4, 17
    rng = np.random.default_rng(11)
    f = [rng.random(n) + 0.2 for _ in range(d)]
    xs = np.concatenate([np.sort(rng.random(n)) * 3 - 1 for _ in range(d)])
    ns, rk = np.full(d, n), np.ones(d + 1, dtype=np.int64)
    q = synth.make_q(500, d, seed=12)
    Zs, ls = tt_irt_sqr_oracle(ns, xs, rk, np.concatenate(f), q)
    Zr, lr = _reference_tt_irt1(ns, xs, rk, np.concatenate([v ** 2 for v in f]), q)
    np.testing.assert_allclose(Zs, Zr, rtol=0, atol=2e-10)       # the reference's own root formula cancels in absolute coordinates
    assert np.median(np.abs(Zs - Zr)) < 1e-14
    np.testing.assert_allclose(ls, lr, rtol=0, atol=1e-9)


def test_sweep_pinned_by_the_reference_c_routine_through_the_first_coordinate():
    """General ranks: the first coordinate only sees the backward sweep.  tt_irt1 on the Kronecker-squared TT (rank r^2)
    integrates the same squared interpolant with the same trapezoid weights that tt_irt_sqr folds into its QR factors
    (:49-51, :66-72), so the first column of Z must agree with the live reference."""
    d, n, r = 4, 9, 3
    ns, xs, rk, c = synth.make_tt(d, n, r, seed=14, cores="normal")
    c2, rk2 = _squared_tt(ns, rk, c)
    q = synth.make_q(400, d, seed=15)
    Zs, ls = tt_irt_sqr_oracle(ns, xs, rk, c, q)
    Zr, lr = _reference_tt_irt1(ns, xs, rk2, c2, q)
    np.testing.assert_allclose(Zs[:, 0], Zr[:, 0], rtol=0, atol=1e-10)
    assert np.median(np.abs(Zs[:, 0] - Zr[:, 0])) < 1e-13


def test_forward_transform_inverts_the_inverse_transform():
    """tt_rt_sqr(tt_irt_sqr(q)) = q with the same log-density (the two reference routines share sweep and conditionals)."""
    for (d, n, r, cores, grid) in [(5, 17, 6, "uniform", "uniform"), (4, 12, 5, "normal", "chebyshev")]:
        ns, xs, rk, c = synth.make_tt(d, n, r, seed=3, cores=cores, grid=grid)
        q = synth.make_q(800, d, seed=4)
        Z, l = tt_irt_sqr_oracle(ns, xs, rk, c, q)
        q2, l2 = tt_rt_sqr_oracle(ns, xs, rk, c, Z)
        assert np.abs(q2 - q).max() < (1e-12 if cores == "uniform" else 1e-9)
        assert np.abs(l2 - l).max() < (1e-12 if cores == "uniform" else 1e-8)
    # marginal and boundary-less cores go through the same code path
    ns, xs, rk, c = mk.make_case(4, 8, 3, 21, -1.0, 1.0, "uniform", "uniform", True)
    q = synth.make_q(200, 2, seed=6)
    Z, l = tt_irt_sqr_oracle(ns, xs, rk, c, q)
    q2, l2 = tt_rt_sqr_oracle(ns, xs, rk, c, Z)
    assert q2.shape == (200, 2) and np.abs(q2 - q).max() < 1e-12 and np.abs(l2 - l).max() < 1e-12


def test_forward_transform_pinned_by_the_reference_c_routine_at_grid_nodes():
    """At a grid node x_j the forward transform returns the conditional CDF node value; for a separable density that is
    the normalised trapezoid sum the reference C tt_irt1 inverts, so feeding those q to the live reference returns x_j."""
    d, n = 3, 9
    rng = np.random.default_rng(5)
    f = [rng.random(n) + 0.2 for _ in range(d)]
    grids = [np.sort(rng.random(n)) * 2 - 1 for _ in range(d)]
    xs = np.concatenate(grids)
    ns, rk = np.full(d, n), np.ones(d + 1, dtype=np.int64)
    pts = np.asfortranarray(np.stack([g[rng.integers(1, n - 1, 300)] for g in grids], axis=1))
    qn, ln = tt_rt_sqr_oracle(ns, xs, rk, np.concatenate(f), pts)
    Zr, lr = _reference_tt_irt1(ns, xs, rk, np.concatenate([v ** 2 for v in f]), qn)
    np.testing.assert_allclose(Zr, pts, rtol=0, atol=1e-9)
    np.testing.assert_allclose(lr, ln, rtol=0, atol=1e-8)


def test_python_mirror_checks_its_arguments_before_touching_the_library():
    """tt_irt_sqr.py mirrors the Matlab function's own checks (tt_irt_sqr.m:37-39 grid size; q columns <= d) and the MEX's
    (tracemult.c:83, :87 shape messages); they fire on the host, without a device."""
    from tt_irt_py import tt_irt, tt_irt_sqr
    ns, xs, rk, c = synth.make_tt(3, 5, 2, seed=1)
    f = tt_irt.TTTensor(ns, rk, c)
    q = synth.make_q(8, 3, seed=2)
    with pytest.raises(ValueError):
        tt_irt_sqr.tt_irt_sqr(xs[:-1], f, q)
    with pytest.raises(ValueError):
        tt_irt_sqr.tt_irt_sqr(xs, f, synth.make_q(8, 4, seed=2))
    with pytest.raises(ValueError):
        tt_irt_sqr.tt_rt_sqr(xs[:-2], f, q)
    with pytest.raises(NotImplementedError):
        tt_irt_sqr.tt_dirt_sample({"x0": xs, "F0": f, "x": xs, "F": [], "reference": "uni", "interpolation": "fourier"}, q)
    with pytest.raises(NotImplementedError):
        tt_irt_sqr.tt_dirt_inverse({"x0": xs, "F0": f, "x": xs, "F": [], "reference": "uni", "crossmethod": "build_ftt"}, q)
    assert tt_irt_sqr._reference_sigma("uni") == 0.0 and tt_irt_sqr._reference_sigma("Normal") == 4.0
    assert tt_irt_sqr._reference_sigma("normal 2.5") == 2.5
    assert tt_irt_sqr.flops_per_sample(np.full(32, 65), np.array([1] + [64] * 31 + [1])) == 8938787
