"""CPU tests: the numpy restatements of the reference's sampler helpers (oracle/samplers_oracle.py) against closed
forms and published known answers.  The reference ships no golden vectors for them and its Matlab cannot run here."""
import numpy as np

from oracle import samplers_oracle as so


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32 10 rounds
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = so.philox4x32_10(np.array([ctr], dtype=np.uint32), key)[0]
        assert tuple(int(x) for x in got) == want


def test_uniform_philox_range_and_sharding():
    u = so.uniform_philox(3, 1000, seed=12345)
    assert u.shape == (1000, 3) and (u >= 0).all() and (u < 1).all()
    assert abs(u.mean() - 0.5) < 0.03
    # a slice generated on its own equals the slice of the whole (counter-based)
    assert np.array_equal(so.uniform_philox(3, 100, seed=12345, m0=400), u[400:500])


def test_qmc_lattice_closed_form():
    # N = 8, z = (1, 3), no shift: exact dyadic rationals
    q = so.qmc_lattice(2, 3, [1, 3, 5], [0.0, 0.0])
    assert np.array_equal(q[:, 0], np.arange(8) / 8.0)
    assert np.array_equal(q[:, 1], (3 * np.arange(8) % 8) / 8.0)
    q = so.qmc_lattice(2, 3, [1, 3], [0.5, 0.25])
    assert np.array_equal(q[:, 0], (np.arange(8) / 8.0 + 0.5) % 1.0)
    assert ((q >= 0) & (q < 1)).all()
    assert np.array_equal(so.qmc_lattice(2, 3, [1, 3], [0.5, 0.25], m0=2, M=4), q[2:6])


def test_truncnormal_map_closed_form():
    y = so.truncnormal_map(np.array([0.0, 0.5, 1.0]), 4.0)
    np.testing.assert_allclose(y, [-4.0, 0.0, 4.0], atol=1e-9)
    y = so.truncnormal_map(np.array([0.0, 1.0]), 2.5)
    np.testing.assert_allclose(y, [-2.5, 2.5], atol=1e-12)


def test_iw_stats_closed_forms():
    M = 1000
    lfapp = np.linspace(-3.0, 1.0, M)
    # exact = proposal up to a constant: weights 1, tau 1, H 0, err1 0
    w, isstd, mx, err1, lren = so.iw_prune(lfapp + 0.7, lfapp)
    np.testing.assert_allclose(w, 1.0, atol=1e-14)
    assert isstd < 1e-14 and abs(mx - 1.0) < 1e-14 and err1 < 1e-14 and abs(lren - 0.7) < 1e-14
    assert abs(so.essinv(lfapp + 0.7, lfapp) - 1.0) < 1e-13
    assert so.hellinger(lfapp + 0.7, lfapp) < 1e-7
    # two-valued ratio: half the samples weigh 3x the others -> w = (0.5, 1.5), tau = M*sum(w^2)/sum(w)^2 = 1.25
    dF = np.where(np.arange(M) % 2 == 0, 0.0, np.log(3.0))
    w, isstd, mx, err1, lren = so.iw_prune(lfapp + dF, lfapp)
    np.testing.assert_allclose(np.sort(np.unique(np.round(w, 12))), [0.5, 1.5])
    assert abs(isstd - 0.5) < 1e-12 and abs(mx - 1.5) < 1e-12 and abs(err1 - 0.5) < 1e-12
    assert abs(so.essinv(lfapp + dF, lfapp) - 1.25) < 1e-12
    H2 = 0.5 * ((np.sqrt(0.5) - 1) ** 2 + (np.sqrt(1.5) - 1) ** 2) / 2
    assert abs(so.hellinger(lfapp + dF, lfapp) - np.sqrt(H2)) < 1e-12


def test_mcmc_prune_hand_worked():
    # log weight r = lFex - lFapp; accept i+1 iff exp(r[i+1] - r[c]) >= u[i]
    lfapp = np.zeros(6)
    lfex = np.log(np.array([1.0, 0.5, 2.0, 0.1, 0.1, 4.0]))
    u = np.array([0.6, 0.9, 0.3, 0.3, 0.99])
    # step0: 0.5/1=0.5 < 0.6 reject (c=0); step1: 2/1 >= 0.9 accept (c=2, run of 1 recorded); step2: 0.05 < 0.3 reject;
    # step3: 0.05 < 0.3 reject; step4: 4/2=2 >= 0.99 accept (run of 2 recorded)
    src, nrej, hist = so.mcmc_prune(lfex, lfapp, u)
    assert src.tolist() == [0, 0, 2, 2, 2, 5] and nrej == 3 and hist.tolist() == [1, 1]
    # the transcription of the Python reference loop (test_shock_absorber_tt.py:165-171) gives the same chain
    Z = np.arange(6.0); lPex = lfex.copy(); lPz = lfapp.copy(); n = 0
    for i in range(5):
        alpha = np.exp(lPex[i + 1] - lPex[i] + lPz[i] - lPz[i + 1])
        if alpha < u[i]:
            Z[i + 1] = Z[i]; lPex[i + 1] = lPex[i]; lPz[i + 1] = lPz[i]; n += 1
    assert Z.astype(int).tolist() == src.tolist() and n == nrej
