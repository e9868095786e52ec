"""GPU parity tests (-m gpu) of the squared-density inverse Rosenblatt transform (include/tt_irt_sqr.h) against the numpy
oracle of matlab/samplers/tt_irt_sqr.m and the committed tests/golden/sqr_*.npz.

Bars (the protocol of oracle/parity.py, as for tt_irt1): interval indices equal to the oracle's (a flip is admissible only
where q sits on a CDF node to 1e-13; none observed), |dZ| <= 1e-12 max(1, |Z|) + 8 eps cumsum(cond) entry by entry,
lFapp within 1e-12 relative plus the same admitted Z perturbation carried through the density slope.  The sweep operands
(P{k} of :80 and R'R of :70) agree to 1e-12 relative.  Signed cores make the conditional contraction cancel; there the bar
is zero index flips and 1e-8 (the reference's own noise floor for such cores).
"""
import ctypes
import os

import numpy as np
import pytest

from oracle import parity
from oracle.tt_irt_sqr_oracle import sqr_sweep, tt_irt_sqr_oracle, tt_rt_sqr_oracle, tracemult_oracle
from test_sqr_oracle import GOLD, load_sqr_golden, mk
from tt_irt_py import synth, tt_irt, tt_irt_sqr

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB32 = os.path.join(ROOT, "tt-irt_b200", "tt_irt_py", "tt_irt1_int32.so")
LIB64 = os.path.join(ROOT, "tt-irt_b200", "lib", "libtt_irt1_int64.so")


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if tt_irt.device_count() < 1:
        pytest.fail("no CUDA device: the -m gpu tests need a B200 (there is no CPU fallback)")


SHAPES = [
    # d, n, r, M, grid, boundary-less cores, D
    (1, 9, 1, 500, "uniform", False, None),            # single dimension, r = 1
    (2, 2, 1, 300, "uniform", False, None),            # smallest grid
    (3, 5, 3, 400, "uniform", False, None),
    (4, 9, 4, 1000, "uniform", False, None),
    (8, 17, 8, 3000, "uniform", False, None),          # BASELINE configs[0] shape
    (6, 17, 16, 3000, "chebyshev", False, None),       # DIRT layer grid
    (5, 15, 6, 1500, "uniform", True, None),           # grid carries boundary points the cores lack (:33-36, :53-60)
    (6, 33, 32, 1500, "uniform", False, 4),            # marginal of the first 4 variables (:9, :105)
    (5, 12, 40, 1200, "uniform", False, None),         # rank not a multiple of 8, grid not 8 j + 1
    (4, 65, 64, 1024, "uniform", False, None),         # BASELINE configs[2] class (lone-column kernel variant)
    (3, 72, 64, 600, "chebyshev", False, None),        # largest supported shape
    (4, 6, 30, 800, "uniform", False, None),           # n s < r: rank-deficient QR factor
    (3, 40, 7, 777, "uniform", False, None),           # odd rank, ragged sample count
]


@pytest.mark.parametrize("d,n,r,M,grid,ext,D", SHAPES)
def test_sweep_and_samples_match_the_oracle(d, n, r, M, grid, ext, D):
    ns, xs, rk, c = mk.make_case(d, n, r, 300 + d + n + r, -1.0, 1.0, grid, "uniform", ext)
    D = d if D is None else D
    q = synth.make_q(M, D, seed=11)
    sw = sqr_sweep(ns, xs, rk, c)
    Zo, lo, io, cond, gap, lsens = tt_irt_sqr_oracle(ns, xs, rk, c, q, extras=True)
    md = tt_irt_sqr.SqrModel(ns, xs, rk, c)
    try:
        for k in range(d):
            G, RR = md.sweep(k)
            assert np.abs(G - sw["P"][k]).max() <= 1e-12 * np.abs(sw["P"][k]).max()
            if k > 0:
                RRo = sw["R"][k] @ sw["R"][k].T
                assert np.abs(RR - RRo).max() <= 1e-12 * np.abs(RRo).max()
        Z, lF, idx = md.sample(q, want_idx=True)
    finally:
        md.close()
    st, fails = parity.compare(Z, lF, idx, Zo, lo, io, cond, gap, lsens)
    assert not fails, (fails, st)
    assert st["idx_flips"] == 0


def test_signed_cores_keep_indices_and_stay_at_the_noise_floor():
    ns, xs, rk, c = synth.make_tt(5, 20, 12, seed=337, cores="normal")
    q = synth.make_q(2000, 5, seed=11)
    Zo, lo, io, cond, gap, lsens = tt_irt_sqr_oracle(ns, xs, rk, c, q, extras=True)
    md = tt_irt_sqr.SqrModel(ns, xs, rk, c)
    try:
        Z, lF, idx = md.sample(q, want_idx=True)
    finally:
        md.close()
    assert (idx != io).sum() == 0
    assert np.abs(Z - Zo).max() < 1e-8 and np.abs(lF - lo).max() < 1e-8


@pytest.mark.parametrize("name", GOLD)
def test_golden_fixtures(name):
    g, ns, xs, rk, c, q = load_sqr_golden(name)
    md = tt_irt_sqr.SqrModel(ns, xs, rk, c)
    try:
        Z, lF, idx = md.sample(q, want_idx=True)
    finally:
        md.close()
    st, fails = parity.compare(Z, lF, idx, g["xq"], g["lFapp"], g["idx"], g["cond"], g["gap"], g["lsens"])
    assert not fails, (fails, st)


def _c_call(lib_path, it, ct, ns, xs, rk, c, q):
    lib = ctypes.CDLL(lib_path)
    M, D = q.shape
    Z = np.zeros((M, D), order="F")
    lF = np.zeros(M)
    n_ = np.ascontiguousarray(ns, dtype=it)
    r_ = np.ascontiguousarray(rk, dtype=it)
    dp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ct)
    lib.tt_irt_sqr.restype = None
    lib.tt_irt_sqr.argtypes = [ct, ip, ct, dp, ip, dp, ct, ct, dp, dp, dp]
    lib.tt_irt_sqr(len(ns), n_.ctypes.data_as(ip), xs.size, xs.ctypes.data_as(dp), r_.ctypes.data_as(ip), c.ctypes.data_as(dp), M, D,
                   q.ctypes.data_as(dp), Z.ctypes.data_as(dp), lF.ctypes.data_as(dp))
    return Z, lF


def test_c_symbol_both_widths_and_python_mirror_agree_bit_for_bit():
    ns, xs, rk, c = mk.make_case(6, 17, 8, 77, -2.0, 2.0, "uniform", "uniform", True)
    q = synth.make_q(5000, 6, seed=3)
    md = tt_irt_sqr.SqrModel(ns, xs, rk, c)
    try:
        Z0, l0 = md.sample(q)
    finally:
        md.close()
    Z32, l32 = _c_call(LIB32, np.int32, ctypes.c_int, ns, xs, rk, c, q)
    Z64, l64 = _c_call(LIB64, np.int64, ctypes.c_longlong, ns, xs, rk, c, q)
    Zp, lp = tt_irt_sqr.tt_irt_sqr(xs, tt_irt.TTTensor(ns, rk, c), q)
    for Z, l in ((Z32, l32), (Z64, l64), (Zp, lp)):
        np.testing.assert_array_equal(Z, Z0)
        np.testing.assert_array_equal(l, l0)


def test_full_size_properties():
    """At a size the oracle cannot walk in seconds: row-subset invariance against the oracle, marginal = prefix of the full
    transform, chunk invariance, determinism, samples inside their grid cell, monotone first coordinate."""
    d, n, r, M = 6, 33, 32, 1 << 17
    ns, xs, rk, c = synth.make_tt(d, n, r, seed=19, lo=-3.0, hi=3.0)
    q = synth.make_q(M, d, seed=23)
    md = tt_irt_sqr.SqrModel(ns, xs, rk, c)
    try:
        Z, lF, idx = md.sample(q, want_idx=True)
        Z2, lF2 = md.sample(q)
        Zm, lm = md.sample(q[:, :3])
        os.environ["TTIRT_SQR_CHUNK"] = "10000"
        try:
            Zc, lc = md.sample(q)
        finally:
            del os.environ["TTIRT_SQR_CHUNK"]
    finally:
        md.close()
    np.testing.assert_array_equal(Z, Z2)
    np.testing.assert_array_equal(lF, lF2)
    np.testing.assert_array_equal(Z, Zc)
    np.testing.assert_array_equal(lF, lc)
    np.testing.assert_array_equal(Zm, Z[:, :3])
    rows = np.random.default_rng(1).choice(M, 512, replace=False)
    Zo, lo, io, cond, gap, lsens = tt_irt_sqr_oracle(ns, xs, rk, c, q[rows], extras=True)
    st, fails = parity.compare(Z[rows], lF[rows], idx[rows], Zo, lo, io, cond, gap, lsens)
    assert not fails, (fails, st)
    for k in range(d):
        x = xs[k * n:(k + 1) * n]
        assert (Z[:, k] >= x[idx[:, k]]).all() and (Z[:, k] <= x[idx[:, k] + 1]).all()
    o = np.argsort(q[:, 0], kind="stable")
    assert (np.diff(Z[o, 0]) >= 0).all()
    assert np.isfinite(lF).all()


def test_failures_are_loud():
    ns, xs, rk, c = synth.make_tt(3, 9, 4, seed=1)
    q = synth.make_q(64, 3, seed=2)
    Z, l = _c_call(LIB32, np.int32, ctypes.c_int, ns, xs[:-1].copy(), rk, c, q)      # grid size matches neither sum(n) nor sum(n + 2)
    assert np.isnan(Z).all() and np.isnan(l).all()
    with pytest.raises(RuntimeError):
        tt_irt_sqr.SqrModel(*synth.make_tt(3, 9, 70, seed=1))                        # rank beyond the supported shapes
    with pytest.raises(RuntimeError):
        tt_irt_sqr.SqrModel(*synth.make_tt(3, 80, 4, seed=1))                        # grid beyond the supported shapes
    md = tt_irt_sqr.SqrModel(ns, xs, rk, c)
    try:
        with pytest.raises(RuntimeError):
            md.sample(synth.make_q(16, 4, seed=2))                                   # more seed columns than dimensions
    finally:
        md.close()


def test_device_sharding_does_not_change_a_bit():
    """ttirt_sqr_run_host over 2 devices (contiguous row ranges, cores replicated) against one device."""
    if tt_irt.device_count() < 2:
        pytest.skip("needs two GPUs")
    ns, xs, rk, c = synth.make_tt(5, 17, 8, seed=4)
    q = synth.make_q(40000, 5, seed=5)
    Z1, l1 = _c_call(LIB32, np.int32, ctypes.c_int, ns, xs, rk, c, q)
    os.environ["TTIRT_DEVICES"] = "2"
    try:
        Z2, l2 = _c_call(LIB32, np.int32, ctypes.c_int, ns, xs, rk, c, q)
    finally:
        del os.environ["TTIRT_DEVICES"]
    np.testing.assert_array_equal(Z1, Z2)
    np.testing.assert_array_equal(l1, l2)


@pytest.mark.parametrize("p,m,k,n,s", [(1, 64, 64, 5000, 65), (8, 1, 8, 3000, 3000), (3, 4, 5, 777, 6), (16, 16, 1, 100, 17), (1, 1, 1, 1, 1)])
def test_tracemult_operator(p, m, k, n, s):
    """The standalone tracemult (reference matlab/utils/tracemult.c, real case) against its numpy restatement: the indexed
    batched product to 1e-13 relative (one fused multiply-add chain per element instead of dgemm's order), the pick bit-exact."""
    rng = np.random.default_rng(p + m + k)
    A = np.asfortranarray(rng.standard_normal((p, m, n)))
    B = np.asfortranarray(rng.standard_normal((m, k, s)))
    j = rng.integers(1, s + 1, n).astype(float)
    C = tt_irt_sqr.tracemult(A, j, B)
    Co = tracemult_oracle(A, j, B)
    assert C.shape == (p, k, n)
    assert np.abs(C - Co).max() <= 1e-13 * max(1.0, np.abs(Co).max()) * m
    A2 = np.asfortranarray(rng.standard_normal((n, s)))
    np.testing.assert_array_equal(tt_irt_sqr.tracemult(A2, j), tracemult_oracle(A2, j))
    with pytest.raises(RuntimeError):
        tt_irt_sqr.tracemult(A2, np.full(n, s + 1.0))          # the MEX would read out of bounds: fail instead


def test_edge_seeds_and_zero_mass_fallback():
    """Seeds exactly 0 and 1 (ties of the bisection go left, :139-142, q >= 1 ends in the last cell) and conditionals without
    mass (:121-127: the conditional is replaced by h, its CDF by cumsum(h)): same cells, same samples, same -inf pattern."""
    ns, xs, rk, c = synth.make_tt(3, 9, 4, seed=2)
    q = synth.make_q(64, 3, seed=3)
    q[:8, :] = 0.0
    q[8:16, :] = 1.0
    q[16:24, 0] = 0.0
    q[24:32, 2] = 1.0
    q = np.asfortranarray(q)
    for cores in (c, None):
        if cores is None:                       # a vanishing core: every conditional is zero, every dimension falls back
            cores = c.copy()
            off = int(rk[0] * ns[0] * rk[1])
            cores[off:off + int(rk[1] * ns[1] * rk[2])] = 0.0
        Zo, lo, io, cond, gap, lsens = tt_irt_sqr_oracle(ns, xs, rk, cores, q, extras=True)
        md = tt_irt_sqr.SqrModel(ns, xs, rk, cores)
        try:
            Z, lF, idx = md.sample(q, want_idx=True)
        finally:
            md.close()
        np.testing.assert_array_equal(idx, io)
        np.testing.assert_allclose(Z, Zo, rtol=0, atol=1e-12)
        np.testing.assert_array_equal(np.isfinite(lF), np.isfinite(lo))
        fin = np.isfinite(lo)
        np.testing.assert_allclose(lF[fin], lo[fin], rtol=0, atol=1e-11)
        np.testing.assert_array_equal(lF[~fin], lo[~fin])


@pytest.mark.parametrize("d,n,r,M,ext,D", [(4, 9, 4, 1000, False, None), (8, 17, 8, 3000, False, None), (5, 15, 6, 1500, True, None),
                                           (6, 33, 32, 1500, False, 4), (4, 65, 64, 1024, False, None), (5, 12, 40, 1200, False, None)])
def test_forward_transform_matches_the_oracle(d, n, r, M, ext, D):
    """tt_rt_sqr (reference matlab/samplers/tt_rt_sqr.m): CDF values to 1e-12 absolute (they live in [0, 1] and are sums of
    the same conditionals), cells equal, log-density to 1e-12 relative + the conditional's own conditioning."""
    ns, xs, rk, c = mk.make_case(d, n, r, 300 + d + n + r, -1.0, 1.0, "uniform", "uniform", ext)
    D = d if D is None else D
    q = synth.make_q(M, D, seed=11)
    X, _ = tt_irt_sqr_oracle(ns, xs, rk, c, q)           # points inside the support
    Qo, lo = tt_rt_sqr_oracle(ns, xs, rk, c, X)
    md = tt_irt_sqr.SqrModel(ns, xs, rk, c)
    try:
        Q, lF, idx = md.forward(X, want_idx=True)
        Z, lz = md.sample(q)
        Q2, l2 = md.forward(Z)                          # device round trip
    finally:
        md.close()
    assert np.abs(Q - Qo).max() <= 1e-12, np.abs(Q - Qo).max()
    assert np.abs(lF - lo).max() <= 1e-11 * max(1.0, np.abs(lo).max())
    assert np.abs(Q2 - q).max() <= 1e-11 and np.abs(l2 - lz).max() <= 1e-11 * max(1.0, np.abs(lz).max())
    Qc, lc = tt_irt_sqr.tt_rt_sqr(xs, tt_irt.TTTensor(ns, rk, c), X)      # the C symbol through the Python mirror
    np.testing.assert_array_equal(Qc, Q)
    np.testing.assert_array_equal(lc, lF)


def test_round_trip_at_full_size():
    """Encode -> decode at a size the oracle cannot walk: tt_rt_sqr(tt_irt_sqr(q)) = q for 2^17 seeds."""
    d, n, r, M = 6, 33, 32, 1 << 17
    ns, xs, rk, c = synth.make_tt(d, n, r, seed=19, lo=-3.0, hi=3.0)
    q = synth.make_q(M, d, seed=23)
    md = tt_irt_sqr.SqrModel(ns, xs, rk, c)
    try:
        Z, lF = md.sample(q)
        Q, l2 = md.forward(Z)
    finally:
        md.close()
    assert np.abs(Q - q).max() <= 1e-10 and np.median(np.abs(Q - q)) <= 1e-14
    assert np.abs(l2 - lF).max() <= 1e-10 * max(1.0, np.abs(lF).max())


def test_repeated_calls_with_changing_shapes_are_stable_and_bounded():
    """tt_irt_sqr keeps no state the caller can see; its device blocks go back to a per-device pool (bounded by TTIRT_POOL_MB).
    Alternating shapes and sizes must repeat bit for bit, and device memory must not creep once every shape has been seen."""
    import torch
    shapes = [(6, 17, 8, 1 << 14), (5, 33, 32, 70000), (6, 17, 8, 300000), (4, 65, 64, 20000), (3, 9, 4, 1000)]
    data, first = [], {}
    for i, (d, n, r, M) in enumerate(shapes):
        ns, xs, rk, c = synth.make_tt(d, n, r, seed=50 + i)
        data.append((tt_irt.TTTensor(ns, rk, c), xs, synth.make_q(M, d, seed=60 + i)))
    free0 = None
    for rep in range(8):
        for i, (f, xs, q) in enumerate(data):
            Z, l = tt_irt_sqr.tt_irt_sqr(xs, f, q)
            if i not in first:
                first[i] = (Z.copy(), l.copy())
                assert np.isfinite(Z).all() and np.isfinite(l).all()
            else:
                assert np.array_equal(Z, first[i][0]) and np.array_equal(l, first[i][1]), (rep, i)
        free = torch.cuda.mem_get_info(0)[0]
        if rep == 1:
            free0 = free
        if rep > 1:
            assert free >= free0 - (64 << 20), "device memory shrank by %d MB over repeated calls" % ((free0 - free) >> 20)
    before = torch.cuda.mem_get_info(0)[0]
    tt_irt.load_library().ttirt_cache_clear()
    assert torch.cuda.mem_get_info(0)[0] >= before          # the pool's idle blocks went back to the driver


@pytest.mark.parametrize("d,n,r,M,cores", [(4, 65, 64, 3000, "uniform"), (6, 33, 32, 5000, "uniform"), (8, 17, 16, 20000, "uniform"),
                                           (5, 20, 12, 4000, "normal"), (5, 12, 40, 1200, "uniform"), (4, 40, 7, 999, "uniform")])
def test_fused_and_separate_tail_agree_bit_for_bit(d, n, r, M, cores):
    """The tail fused into the conditional-pdf kernel (no CDF scratch: mass, then a guarded count against q * mass, the
    reference bisection only when the count is ambiguous or the CDF is not monotone) and the separate tail kernel
    (TTIRT_SQR_FUSED=0) must give identical bits, in both directions; seeds on CDF nodes and at 0 / 1 included."""
    ns, xs, rk, c = synth.make_tt(d, n, r, seed=70 + d, cores=cores)
    q = synth.make_q(M, d, seed=71)
    q[:16, :] = 0.0
    q[16:32, :] = 1.0
    q[32:48, 0] = 0.5
    q = np.asfortranarray(q)
    md = tt_irt_sqr.SqrModel(ns, xs, rk, c)
    try:
        # "2": fused tail in every class (the default leaves the r <= 64 class unfused); "1": fused tail without the fused
        # interface update of small cores (a different summation order, so not bit-comparable with the sorted DMMA update)
        os.environ["TTIRT_SQR_FUSED"] = "2" if n > 40 else "1"
        try:
            Zf, lf, idf = md.sample(q, want_idx=True)
            Qf, l2f = md.forward(Zf)
            os.environ["TTIRT_SQR_FUSED"] = "0"
            Zs, ls, ids = md.sample(q, want_idx=True)
            Qs, l2s = md.forward(Zf)
        finally:
            del os.environ["TTIRT_SQR_FUSED"]
    finally:
        md.close()
    np.testing.assert_array_equal(idf, ids)
    np.testing.assert_array_equal(Zf, Zs)
    np.testing.assert_array_equal(lf, ls)
    np.testing.assert_array_equal(Qf, Qs)
    np.testing.assert_array_equal(l2f, l2s)


@pytest.mark.parametrize("d,n,r,M", [(8, 17, 16, 20000), (11, 17, 16, 5000), (8, 17, 8, 7000), (6, 20, 12, 3000), (5, 9, 5, 999)])
def test_fully_fused_step_of_small_cores_matches_the_oracle_and_the_sorted_path(d, n, r, M):
    """Small cores (the whole core fits in shared memory next to the tiles): contraction, tail and interface update of a
    dimension run in one kernel, without the sort.  Against the oracle inside the protocol; against the sorted DMMA update
    (TTIRT_SQR_FUSED=1) inside the same protocol (the two differ by summation order only) with identical cells."""
    ns, xs, rk, c = synth.make_tt(d, n, r, seed=80 + d)
    q = synth.make_q(M, d, seed=81)
    Zo, lo, io, cond, gap, lsens = tt_irt_sqr_oracle(ns, xs, rk, c, q, extras=True)
    md = tt_irt_sqr.SqrModel(ns, xs, rk, c)
    try:
        l0 = tt_irt.kernel_launches()
        Z, lF, idx = md.sample(q, want_idx=True)
        launches = tt_irt.kernel_launches() - l0
        os.environ["TTIRT_SQR_FUSED"] = "1"
        try:
            Z1, l1, i1 = md.sample(q, want_idx=True)
        finally:
            del os.environ["TTIRT_SQR_FUSED"]
    finally:
        md.close()
    assert launches == d + 1                        # init + one kernel per dimension
    st, fails = parity.compare(Z, lF, idx, Zo, lo, io, cond, gap, lsens)
    assert not fails, (fails, st)
    st1, fails1 = parity.compare(Z, lF, idx, Z1, l1, i1, cond, gap, lsens)      # the two update arithmetics: same protocol
    assert not fails1 and st1["idx_flips"] == 0, (fails1, st1)
    assert np.median(np.abs(Z - Z1)) <= 1e-15
