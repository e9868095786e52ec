"""BASELINE.json configs[0] as the reference runs it (python/test_shock_absorber_tt.py:138-180): seeds -> tt_irt1 -> exact
density -> independence Metropolis-Hastings -> quantile of interest, on the D = 2 shock-absorber posterior compressed by
TT-SVD (ttpy's cross is absent).  The fixture tests/golden/shock_mh_D2.npz was produced by
executing the REFERENCE's own Python (density functions, MH loop, quantile) from its source file
(tests/golden/make_golden_mh.py), so it pins

  * oracle/samplers_oracle.py::mcmc_prune  (the restatement the GPU prune is tested against) to the reference's loop,
  * oracle/shock_absorber_oracle.py        (the exact density needed at test time) to the reference's functions,

and the GPU test runs the chain end to end through the C-ABI: tt_irt1 -> ttirt_mcmc_prune_host.
"""
import hashlib
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def fx():
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "shock_mh_D2.npz"), allow_pickle=False))
    from oracle import shock_absorber_oracle as S
    M, D = int(g["M"]), int(g["n"].size)
    g["q"] = S.draw_seeds(M, D, int(g["q_seed"]))
    g["u"] = S.draw_uniforms(M - 1, int(g["u_seed"]))
    assert hashlib.sha256(np.ascontiguousarray(g["q"]).tobytes()).hexdigest() == str(g["q_sha256"]), "legacy np.random stream drifted"
    assert hashlib.sha256(g["u"].tobytes()).hexdigest() == str(g["u_sha256"]), "legacy np.random stream drifted"
    return g


def test_oracle_chain_reproduces_the_reference_python(oracle_mod, fx):
    """CPU: C oracle tt_irt1 -> restated density -> restated MH loop == the reference's Python on every sample."""
    from oracle import samplers_oracle as SO, shock_absorber_oracle as S
    Z, lPz = oracle_mod.oracle_run(fx["n"], fx["xs"], fx["ranks"], fx["cores"], fx["q"])
    assert hashlib.sha256(np.asfortranarray(Z).tobytes(order="F")).hexdigest() == str(fx["Z_oracle_sha256"])
    assert np.array_equal(lPz, fx["lPz_oracle"])
    lPex = S.log_posterior(Z, fx["x"], fx["y"], fx["censind"], fx["beta_mean"], fx["beta_var"])
    fin = np.isfinite(fx["lPex_ref"])
    assert np.array_equal(np.isfinite(lPex), fin)
    np.testing.assert_allclose(lPex[fin], fx["lPex_ref"][fin], rtol=1e-13, atol=1e-11)
    # the restated prune (Matlab operation order, mcmc_prune.m:25) against the reference's Python loop (:164-171, a
    # different association of the same four terms): same accept / reject decision on all 16383 proposals
    src, nrej, hist = SO.mcmc_prune(fx["lPex_ref"], fx["lPz_oracle"], fx["u"])
    assert np.array_equal(src, fx["src_ref"])
    assert nrej == int(fx["num_of_rejects_ref"]) and int(hist.sum()) > 0
    qoi = S.quantile_of_interest(Z[src], int(fx["d_cov"])).mean()
    assert abs(qoi - float(fx["q_post_mean_ref"])) <= 1e-12 * abs(qoi)


@pytest.mark.gpu
def test_gpu_chain_end_to_end_through_the_c_abi(oracle_mod, fx):
    """GPU: the drop-in symbol tt_irt1 and ttirt_mcmc_prune_host in place of the reference's C sampler and Python loop."""
    from oracle import shock_absorber_oracle as S
    from tt_irt_py import samplers, tt_irt
    if tt_irt.device_count() < 1:
        pytest.fail("no CUDA device")
    f = tt_irt.TTTensor(fx["n"], fx["ranks"], fx["cores"])
    Z, lPz = tt_irt.tt_irt1(fx["q"], f, fx["xs"])
    Zo, lo, io, kap, gap, cond, lsens = oracle_mod.oracle_run(fx["n"], fx["xs"], fx["ranks"], fx["cores"], fx["q"], extras=True)
    stats, fails = oracle_mod.parity.compare(Z, lPz, None, Zo, lo, None, cond, gap, lsens=lsens)
    assert not fails, (fails, stats)
    lPex = S.log_posterior(Z, fx["x"], fx["y"], fx["censind"], fx["beta_mean"], fx["beta_var"])
    Zp, _, _, nrej, hist, src = samplers.mcmc_prune(Z, lPex, lPz, fx["u"])   # ttirt_mcmc_prune_host underneath
    assert np.array_equal(Zp, Z[src])
    # Z and lPz differ from the CPU chain by ~1e-13, so an accept test decided by less than that could flip; none does here
    assert np.array_equal(src, fx["src_ref"])
    assert nrej == int(fx["num_of_rejects_ref"])
    qoi = S.quantile_of_interest(Z[src], int(fx["d_cov"])).mean()
    assert abs(qoi - float(fx["q_post_mean_ref"])) <= 1e-10 * abs(qoi)
