"""GPU tests (-m gpu): the device kernels either side of tt_irt1 DIRECTLY against the outputs of the reference's own Matlab sources
(tests/golden/matlab_*.npz: qmcnodes.m, randref.m, essinv.m, hellinger.m, iw_prune.m, mcmc_prune.m, tt_irt_sqr.m, tt_rt_sqr.m,
tt_dirt_sample.m, tt_dirt_inverse.m executed from source by oracle/mlite.py, tracemult.c compiled unmodified; generator
tests/golden/make_golden_matlab.py, CPU side tests/test_matlab_pins.py).  The other GPU tests compare the device with the numpy
restatements; these close the chain without them: reference source -> fixture -> device.

Bars: lattice seeds, uniform pass-through, Metropolis-Hastings chain and tracemult's column pick bit for bit; truncated-normal map
1e-13; importance-weight statistics 1e-12 relative; squared-density transforms by the tt_irt1 protocol (oracle/parity.py) with
the FIXTURE as the reference values and the oracle's sensitivities as the scale; DIRT loops 1e-9 / 1e-8 per entry (one level's
admitted perturbation is amplified by the next level's 1 / p) and 1e-12 in the median.
"""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden_matlab as gen  # noqa: E402

from oracle import parity  # noqa: E402
from oracle.tt_irt_sqr_oracle import tt_irt_sqr_oracle  # noqa: E402
from tt_irt_py import samplers, tt_irt, tt_irt_sqr  # noqa: E402

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(HERE, "golden")


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if tt_irt.device_count() < 1:
        pytest.fail("no CUDA device: the -m gpu tests need a B200 (there is no CPU fallback)")


def _load(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))


def test_device_seeds_against_the_reference_matlab_sources():
    g, inp = _load("matlab_helpers"), gen.helper_inputs()
    q = samplers.qmcnodes(inp["d"], inp["l"], inp["table"][:, 1].astype(np.int64), inp["shift"][:, 0])
    assert np.array_equal(q, g["qmc_Y"].T)                                            # qmcnodes.m:6-13, bit for bit
    for tag, ref in (("normal", "Normal"), ("normal3", "normal 3"), ("n2p5", "n2.5")):
        np.testing.assert_allclose(samplers.randref(ref, inp["u"]), g["randref_" + tag], rtol=1e-13, atol=1e-13)   # randref.m:22-34
    assert np.array_equal(samplers.randref("UNI", inp["u"]), g["randref_uni"])


def test_device_consumers_against_the_reference_matlab_sources():
    g, inp = _load("matlab_helpers"), gen.helper_inputs()
    lFapp = inp["lFapp"][:, 0]
    scaled, isstd, mx, err1 = samplers.iw_prune(inp["lFex"], lFapp)                   # iw_prune.m:19-29
    np.testing.assert_allclose(scaled, g["iw_lFex"], rtol=1e-12)
    np.testing.assert_allclose([isstd, mx, err1], [float(g[k].reshape(-1)[0]) for k in ("iw_isstd", "iw_max_ratio", "iw_err1")], rtol=1e-12)
    np.testing.assert_allclose(samplers.essinv(inp["lFex"][:, 0], lFapp), float(g["essinv"].reshape(-1)[0]), rtol=1e-12)        # essinv.m:12-14
    np.testing.assert_allclose(samplers.hellinger(inp["lFex"][:, 0], lFapp), float(g["hellinger"].reshape(-1)[0]), rtol=1e-12)  # hellinger.m:12-16
    y, lfe, lfa, nrej, hist, src = samplers.mcmc_prune(inp["mh_y"], inp["mh_lFex"], inp["mh_lFapp"][:, 0], inp["mh_u"])          # mcmc_prune.m:24-43
    assert nrej == int(g["mh_num_of_rejects"].reshape(-1)[0])
    assert np.array_equal(y, g["mh_y"]) and np.array_equal(lfe, g["mh_lFex"]) and np.array_equal(lfa, g["mh_lFapp"].reshape(-1))
    want = g["mh_rej_distribution"].reshape(-1).astype(np.int64)
    assert np.array_equal(hist[:want.size], want) and not hist[want.size:].any()


def test_device_tracemult_against_the_reference_mex_source():
    g = _load("matlab_tracemult")
    A, B, j, A2, j2 = gen.tracemult_inputs()
    np.testing.assert_allclose(tt_irt_sqr.tracemult(A, j[:, 0], B), g["C"], rtol=1e-13, atol=1e-13)     # tracemult.c:103-112
    assert np.array_equal(tt_irt_sqr.tracemult(A2, j2[:, 0]), g["C2"].reshape(-1))                      # :131-136


@pytest.mark.parametrize("case", gen.SQR_CASES, ids=[c[0] for c in gen.SQR_CASES])
def test_device_squared_density_transforms_against_the_reference_matlab_sources(case):
    g = _load("matlab_sqr_" + case[0])
    ns, xs, rk, c, q = gen.sqr_inputs(case)
    # the oracle supplies the per-entry sensitivities of the protocol; the reference values are the fixture's
    _, _, io, cond, gap, lsens = tt_irt_sqr_oracle(ns, xs, rk, c, q, extras=True)
    f = tt_irt.TTTensor(ns, rk, c)
    Z, lF = tt_irt_sqr.tt_irt_sqr(xs, f, q)                                           # tt_irt_sqr.m:1-208
    if case[6] == "normal":
        # signed cores: the contraction cancels, both sides sit at its noise floor (as tests/test_sqr_gpu.py's signed-core test)
        assert np.abs(Z - g["xq"]).max() < 1e-8 and np.abs(lF - g["lFapp"]).max() < 1e-8
    else:
        st, fails = parity.compare(Z, lF, None, g["xq"], g["lFapp"], None, cond, gap, lsens)
        assert not fails, (fails, st)
    q2, lF2 = tt_irt_sqr.tt_rt_sqr(xs, f, g["xq"])                                    # tt_rt_sqr.m:1-178
    tol = 1e-8 if case[6] == "normal" else 1e-11
    np.testing.assert_allclose(q2, g["rt_q"], rtol=0, atol=tol)
    np.testing.assert_allclose(lF2, g["rt_lFapp"], rtol=10 * tol, atol=10 * tol)


@pytest.mark.parametrize("case", gen.DIRT_CASES, ids=[c[0] for c in gen.DIRT_CASES])
def test_device_dirt_loops_against_the_reference_matlab_sources(case):
    g = _load("matlab_dirt_" + case[0])
    levels, q, reference = gen.dirt_inputs(case)
    drt = tt_irt_sqr.Dirt(levels, reference)
    try:
        z, lF = drt.sample(q)                                                         # tt_dirt_sample.m:17-73
        q2, lF2 = drt.inverse(g["z"])                                                 # tt_dirt_inverse.m:24-59
    finally:
        drt.close()
    dz = np.abs(z - g["z"]) / np.maximum(1.0, np.abs(g["z"]))
    dl = np.abs(lF - g["lFapp"])
    assert dz.max() <= 1e-9 and dl.max() <= 1e-8, (dz.max(), dl.max())
    assert np.median(dz) <= 1e-12 and np.median(dl) <= 1e-11, (np.median(dz), np.median(dl))
    tol = 1e-10 if reference[0].lower() == "u" else 1e-7
    assert np.abs(q2 - g["inv_q"]).max() <= tol and np.abs(lF2 - g["inv_lFapp"]).max() <= 10 * tol, (np.abs(q2 - g["inv_q"]).max(), np.abs(lF2 - g["inv_lFapp"]).max())
