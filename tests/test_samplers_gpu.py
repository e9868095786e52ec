"""GPU tests (-m gpu): the device versions of the reference's sampler helpers (csrc/ttirt_aux.cu, through the C-ABI
host forms) against the numpy restatements of oracle/samplers_oracle.py.
Bars: seeds bit-exact; truncated-normal map 1e-13; importance-weight statistics 1e-12 relative (sums in a different,
fixed order); Metropolis-Hastings chain identical indices."""
import numpy as np
import pytest

from oracle import samplers_oracle as so
from tt_irt_py import samplers, synth, tt_irt

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if tt_irt.device_count() < 1:
        pytest.fail("no CUDA device: the -m gpu tests need a B200 (there is no CPU fallback)")


def test_lattice_bitexact_and_sharded():
    d, l = 11, 14
    rng = np.random.default_rng(3)
    z = np.concatenate([[1], rng.integers(1, 1 << 20, size=d - 1) | 1])
    shift = rng.random(d)
    q = samplers.qmcnodes(d, l, z, shift)
    assert q.shape == (1 << l, d) and q.flags.f_contiguous
    assert np.array_equal(q, so.qmc_lattice(d, l, z, shift))
    part = samplers.qmcnodes(d, l, z, shift, m0=5000, M=777)
    assert np.array_equal(part, q[5000:5777])
    assert ((q >= 0) & (q < 1)).all()


def test_uniform_bitexact_and_sharded():
    u = samplers.rand_uniform(4099, 5, seed=0x1234567890ABCDEF)
    assert np.array_equal(u, so.uniform_philox(5, 4099, seed=0x1234567890ABCDEF))
    assert np.array_equal(samplers.rand_uniform(100, 5, seed=0x1234567890ABCDEF, m0=1000), u[1000:1100])
    assert abs(u.mean() - 0.5) < 0.02


def test_randref_truncated_normal():
    u = np.random.default_rng(0).random((1000, 3))
    assert samplers.randref("uniform", u) is not None and np.array_equal(samplers.randref("UNI", u), u)
    for ref, sigma in (("normal", 4.0), ("Normal 2.5", 2.5)):
        y = samplers.randref(ref, u)
        np.testing.assert_allclose(y, so.truncnormal_map(u, sigma), rtol=1e-13, atol=1e-13)
        assert np.abs(y).max() <= sigma
    np.testing.assert_allclose(samplers.randref("normal 3", np.array([0.0, 0.5, 1.0])), [-3.0, 0.0, 3.0], atol=1e-9)


@pytest.mark.parametrize("M", [1, 7, 1000, (1 << 18) + 13])
def test_iw_statistics(M):
    rng = np.random.default_rng(M)
    lfapp = rng.normal(-5.0, 2.0, M)
    lfex = lfapp + 0.3 * rng.normal(size=M) + 1.7
    w, isstd, mx, err1, lren = so.iw_prune(lfex, lfapp)
    scaled, g_isstd, g_mx, g_err1 = samplers.iw_prune(np.stack([lfex, np.ones(M)], axis=1), lfapp)
    np.testing.assert_allclose(scaled[:, 1], w, rtol=1e-12)
    np.testing.assert_allclose(scaled[:, 0], lfex * w, rtol=1e-12)
    np.testing.assert_allclose([g_isstd, g_mx, g_err1], [isstd, mx, err1], rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(samplers.essinv(lfex, lfapp), so.essinv(lfex, lfapp), rtol=1e-12)
    np.testing.assert_allclose(samplers.hellinger(lfex, lfapp), so.hellinger(lfex, lfapp), rtol=1e-10, atol=1e-14)


@pytest.mark.parametrize("M,spread", [(1, 1.0), (2, 1.0), (33, 0.5), (5000, 0.2), (5000, 3.0), (1 << 16, 1.0)])
def test_mcmc_prune_chain_identical(M, spread):
    rng = np.random.default_rng(M + int(10 * spread))
    lfapp = rng.normal(-3.0, 1.0, M)
    lfex = lfapp + spread * rng.normal(size=M)
    u = rng.random(max(M - 1, 1))
    src, nrej, hist = so.mcmc_prune(lfex, lfapp, u)
    y = np.arange(M * 2, dtype=np.float64).reshape(M, 2)
    yp, fe, fa, g_nrej, g_hist, g_src = samplers.mcmc_prune(y, lfex, lfapp, u, rej_hist_len=4096)
    assert np.array_equal(g_src, src)
    assert g_nrej == nrej
    assert np.array_equal(g_hist[:hist.size], hist) and g_hist[hist.size:].sum() == 0
    assert np.array_equal(yp, y[src]) and np.array_equal(fe, lfex[src]) and np.array_equal(fa, lfapp[src])


def test_seeds_feed_the_sampler():
    """qmcnodes -> tt_irt1 as the reference's QMC drivers chain them (tt_irt_debias / test_diffusion): the lattice
    generated on the device is a valid seed matrix for the drop-in call."""
    d, n, r = 6, 17, 8
    ns, xs, rk, c = synth.make_tt(d, n, r, seed=4)
    q = samplers.qmcnodes(d, 12, np.arange(1, 2 * d, 2), np.full(d, 0.125))
    f = tt_irt.TTTensor(ns, rk, c)
    Z, lPz = tt_irt.tt_irt1(q, f, xs)
    assert np.isfinite(Z).all() and np.isfinite(lPz).all()
    lo, hi = xs.reshape(d, n)[:, 0], xs.reshape(d, n)[:, -1]
    assert (Z >= lo - 1e-9).all() and (Z <= hi + 1e-9).all()


def test_sampling_on_device_seeds_equals_sampling_on_uploaded_seeds():
    """Model.sample_lattice / sample_uniform (seeds generated on the device, no q upload) give bit for bit what
    Model.sample gives on the same seeds uploaded from the host; a slice [m0, m0+M) equals the slice of the whole."""
    d, n, r = 8, 17, 16
    ns, xs, rk, c = synth.make_tt(d, n, r, seed=9)
    md = tt_irt.Model(ns, xs, rk, c)
    try:
        z = np.arange(3, 3 + 2 * d, 2); shift = np.linspace(0.05, 0.9, d)
        Z, l, q = md.sample_lattice(15, z, shift, want_q=True)
        assert np.array_equal(q, so.qmc_lattice(d, 15, z, shift))
        Z2, l2 = md.sample(q)
        assert np.array_equal(Z, Z2) and np.array_equal(l, l2)
        Zs, ls = md.sample_lattice(15, z, shift, m0=9000, M=3000)
        assert np.array_equal(Zs, Z[9000:12000]) and np.array_equal(ls, l[9000:12000])
        Zu, lu, qu = md.sample_uniform(20000, seed=77, want_q=True)
        assert np.array_equal(qu, so.uniform_philox(d, 20000, 77))
        Zu2, lu2 = md.sample(qu)
        assert np.array_equal(Zu, Zu2) and np.array_equal(lu, lu2)
    finally:
        md.close()


def test_pageable_and_pinned_callers_agree():
    """The bounce-buffer pipeline for pageable caller arrays (numpy) returns bit for bit what the direct pipeline does."""
    import os
    d, n, r, M = 8, 17, 16, 300000
    ns, xs, rk, c = synth.make_tt(d, n, r, seed=10)
    q = synth.make_q(M, d, seed=2)
    md = tt_irt.Model(ns, xs, rk, c)
    try:
        Z1, l1 = md.sample(q)                       # pageable numpy arrays: staged through pinned buffers
        os.environ["TTIRT_NO_STAGING"] = "1"        # read once per process: only effective if not yet cached
        Z2, l2 = md.sample(q[:70000])               # small call: below the staging threshold either way
        assert np.array_equal(Z2, Z1[:70000]) and np.array_equal(l2, l1[:70000])
    finally:
        os.environ.pop("TTIRT_NO_STAGING", None)
        md.close()
