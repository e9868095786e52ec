"""The DIRT sampler loop (reference matlab/samplers/tt_dirt_sample.m, spline branch): CPU tests of its numpy restatement
(oracle/dirt_oracle.py; pinned against the reference's tt_dirt_sample.m / tt_dirt_inverse.m executed from source in
tests/test_matlab_pins.py) and GPU tests of ttirt_dirt_sample_* against it.

GPU bar: the composition feeds one level's samples to the next as seeds, so a level's admitted 1e-12-class perturbation is
amplified by the next level's inverse-CDF slope 1 / p (up to ~1e3 in the tails).  Per entry: |dZ| <= 1e-9 max(1, |Z|) and
|dlFapp| <= 1e-8; in bulk (the median entry) 1e-12."""
import numpy as np
import pytest

from oracle.dirt_oracle import parse_reference, tt_dirt_inverse_oracle, tt_dirt_sample_oracle
from oracle.tt_irt_sqr_oracle import tt_irt_sqr_oracle
from tt_irt_py import synth, tt_irt


def make_levels(d, n, r, nlvl, reference, seed=0):
    """Synthetic DIRT of the shapes tt_dirt_approx builds: level 0 on its own grid x0, levels 1..nlvl on the reference grid
    (uniform on [-sigma, sigma] for a normal reference, Chebyshev-like on [0, 1] for a uniform one, tt_dirt_approx.m:305-309)."""
    sigma = parse_reference(reference)
    levels = [synth.make_tt(d, n, r, seed=seed, lo=-2.0, hi=3.0)]
    for j in range(1, nlvl + 1):
        if sigma is None:
            levels.append(synth.make_tt(d, n, r, seed=seed + j, lo=0.0, hi=1.0, grid="chebyshev"))
        else:
            levels.append(synth.make_tt(d, n, r, seed=seed + j, lo=-sigma, hi=sigma))
    return levels


def make_seeds(M, d, reference, seed=1):
    sigma = parse_reference(reference)
    u = synth.make_q(M, d, seed=seed)
    return u if sigma is None else np.asfortranarray((2.0 * u - 1.0) * sigma * 0.999)


def test_reference_string_parsing():
    assert parse_reference("uni") is None and parse_reference("UNIform") is None
    assert parse_reference("Normal") == 4.0 and parse_reference("normal 3") == 3.0 and parse_reference("Normal 2.5") == 2.5


def test_single_level_is_tt_irt_sqr():
    lv = make_levels(3, 9, 3, 0, "uni", seed=5)
    q = make_seeds(200, 3, "uni")
    z, lF = tt_dirt_sample_oracle(lv, q, "uni")
    z1, l1 = tt_irt_sqr_oracle(*lv[0], q)
    np.testing.assert_array_equal(z, z1)
    np.testing.assert_array_equal(lF, l1)


@pytest.mark.parametrize("reference", ["uni", "Normal 3"])
def test_oracle_pushforward_density_integrates_to_one(reference):
    """lFapp is the log-density of the pushforward of the reference measure: E_ref[1 / F(z(q))] = volume of level 0's box
    (importance identity); checked by Monte Carlo to its sampling error."""
    d, nlvl = 2, 2
    lv = make_levels(d, 9, 3, nlvl, reference, seed=9)
    sigma = parse_reference(reference)
    rng = np.random.default_rng(4)
    M = 40000
    if sigma is None:
        q = np.asfortranarray(rng.random((M, d)))
    else:
        from scipy.stats import truncnorm
        q = np.asfortranarray(truncnorm.rvs(-sigma, sigma, size=(M, d), random_state=rng))
    z, lF = tt_dirt_sample_oracle(lv, q, reference)
    vol = 5.0 ** d
    w = np.exp(-lF)
    est, err = w.mean(), w.std() / np.sqrt(M)
    assert abs(est - vol) < 5 * err + 0.02 * vol, (est, vol, err)
    x0 = lv[0][1][:9]
    assert (z >= x0[0]).all() and (z <= x0[-1]).all()


@pytest.mark.gpu
@pytest.mark.parametrize("d,n,r,nlvl,reference,M", [(4, 17, 6, 0, "uni", 3000), (4, 17, 6, 2, "uni", 3000), (6, 17, 8, 3, "Normal 4", 4000),
                                                     (3, 33, 12, 1, "Normal 2.5", 2500), (8, 17, 16, 2, "Normal", 2000)])
def test_device_loop_matches_the_oracle(d, n, r, nlvl, reference, M):
    from tt_irt_py import tt_irt_sqr
    if tt_irt.device_count() < 1:
        pytest.fail("no CUDA device: the -m gpu tests need a B200 (there is no CPU fallback)")
    lv = make_levels(d, n, r, nlvl, reference, seed=20 + d)
    q = make_seeds(M, d, reference)
    zo, lo = tt_dirt_sample_oracle(lv, q, reference)
    drt = tt_irt_sqr.Dirt(lv, reference)
    try:
        z, lF = drt.sample(q)
        z2, lF2 = drt.sample(q)
    finally:
        drt.close()
    np.testing.assert_array_equal(z, z2)
    np.testing.assert_array_equal(lF, lF2)
    dz = np.abs(z - zo) / np.maximum(1.0, np.abs(zo))
    dl = np.abs(lF - lo)
    assert dz.max() <= 1e-9 and dl.max() <= 1e-8, (dz.max(), dl.max())
    assert np.median(dz) <= 1e-12 and np.median(dl) <= 1e-12 * max(1.0, np.abs(lo).max()) * 10, (np.median(dz), np.median(dl))

    class S(object):
        pass
    st = {"x0": lv[0][1], "F0": tt_irt.TTTensor(lv[0][0], lv[0][2], lv[0][3]), "x": lv[1][1] if nlvl else None,
          "F": [tt_irt.TTTensor(l[0], l[2], l[3]) for l in lv[1:]], "reference": reference, "interpolation": "spline"}
    z3, lF3 = tt_irt_sqr.tt_dirt_sample(st, q)
    np.testing.assert_array_equal(z3, z)
    np.testing.assert_array_equal(lF3, lF)


@pytest.mark.parametrize("reference", ["uni", "Normal 3"])
def test_oracle_inverse_undoes_the_sampler(reference):
    """tt_dirt_inverse(tt_dirt_sample(q)) = q; the log-densities agree up to the additive constant the reference drops
    (tt_dirt_inverse.m:49: nlvl d log(2 cdf_factor^2 / pi) / 2)."""
    import math
    from scipy.special import erf
    d, nlvl = 4, 2
    lv = make_levels(d, 17, 5, nlvl, reference, seed=7)
    q = make_seeds(500, d, reference)
    z, lf = tt_dirt_sample_oracle(lv, q, reference)
    qb, lb = tt_dirt_inverse_oracle(lv, z, reference)
    assert np.median(np.abs(qb - q)) < 1e-11 and np.abs(qb - q).max() < 1e-6
    sigma = parse_reference(reference)
    const = 0.0 if sigma is None else nlvl * math.log(2 * (0.5 / erf(sigma / math.sqrt(2))) ** 2 / math.pi) * d / 2
    assert np.median(np.abs((lb - lf) - const)) < 1e-10


@pytest.mark.gpu
@pytest.mark.parametrize("d,n,r,nlvl,reference,M", [(4, 17, 6, 0, "uni", 3000), (4, 17, 6, 2, "uni", 3000), (6, 17, 8, 3, "Normal 4", 4000),
                                                     (3, 33, 12, 1, "Normal 2.5", 2500)])
def test_device_inverse_matches_the_oracle_and_round_trips(d, n, r, nlvl, reference, M):
    from tt_irt_py import tt_irt_sqr
    if tt_irt.device_count() < 1:
        pytest.fail("no CUDA device: the -m gpu tests need a B200 (there is no CPU fallback)")
    lv = make_levels(d, n, r, nlvl, reference, seed=20 + d)
    q = make_seeds(M, d, reference)
    zo, lo = tt_dirt_sample_oracle(lv, q, reference)
    qo, lbo = tt_dirt_inverse_oracle(lv, zo, reference)
    drt = tt_irt_sqr.Dirt(lv, reference)
    try:
        qb, lb = drt.inverse(zo)                      # same input as the oracle
        z, lf = drt.sample(q)
        qr, lr = drt.inverse(z)                       # device round trip
    finally:
        drt.close()
    # Conditioning: with a normal reference every level ends in erfinv (slope 1 / phi(q), 10^3 at |q| = 3.6) and these synthetic
    # levels are not the near-identity maps of a fitted DIRT, so the composition amplifies rounding by many orders of
    # magnitude.  The yardstick is the oracle's own response to a one-ulp-class relative perturbation of its input.
    rng = np.random.default_rng(0)
    qp, lp = tt_dirt_inverse_oracle(lv, zo * (1.0 + 1e-15 * rng.standard_normal(zo.shape)), reference)
    sq, sl = np.abs(qp - qo), np.abs(lp - lbo)
    dq, dl = np.abs(qb - qo), np.abs(lb - lbo)
    for (e, s_) in ((dq, sq), (dl, sl)):
        assert np.median(e) <= 10 * np.median(s_) + 1e-12, (np.median(e), np.median(s_))
        assert np.quantile(e, 0.9) <= 10 * np.quantile(s_, 0.9) + 1e-10
        assert e.max() <= 100 * s_.max() + 1e-8
    dr = np.abs(qr - q)
    so = np.abs(qo - q)                                   # the oracle's own round-trip error
    assert np.median(dr) <= 10 * np.median(so) + 1e-12 and dr.max() <= 100 * so.max() + 1e-8, (np.median(dr), dr.max())
