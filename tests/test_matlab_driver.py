"""The reference's top-level Matlab driver, matlab/samplers/tt_irt_debias.m (seeds -> inverse Rosenblatt transform -> exact density
-> Metropolis-Hastings or importance-weight correction, :31-71), executed end to end without Matlab by oracle/mlite.py
(oracle/matlab_driver.py), with its sampler call served three ways:
  as written (tt_irt_lin.m, pure Matlab) | patched to the MEX gateway as install.m:160-169 does, on the reference's own C |
  the same patched call on the drop-in library (the unmodified gateway source linked against libtt_irt1_int64.so).
CPU (build container: needs /root/reference): the first two agree -- same Metropolis-Hastings chain, samples to 1e-11 -- and the
committed fixture tests/golden/matlab_debias.npz is what the second produces.  GPU: the third, i.e. the reference's driver running
on the B200 library, reproduces the fixture: identical rejection count and run-length histogram, samples and importance-weight
statistics to 1e-9 / 1e-10.  On the GPU box the driver's syntax trees come from oracle/_ref/matlab_debias.pkl (parsed here from the
reference sources, which do not exist there)."""
import os

import numpy as np
import pytest

from oracle import matlab_driver, mex_host
from tt_irt_py import tt_irt

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "matlab_debias.npz")
needs_reference = pytest.mark.skipif(not (matlab_driver.reference_available() and mex_host.gateway_available("reference")),
                                     reason="needs /root/reference and oracle/_ref (build container only)")


@needs_reference
def test_driver_as_written_and_patched_to_the_reference_mex_agree():
    compiled = matlab_driver.compile_reference()
    g = dict(np.load(GOLDEN))
    for corr in ("mcmc", "iw"):
        a = matlab_driver.run_debias("matlab", corr, compiled)
        b = matlab_driver.run_debias("reference", corr, compiled)
        for k in b:
            assert np.array_equal(b[k], g[corr + "_" + k]), (corr, k)          # the fixture is what the reference produces here
        assert np.abs(a["y"] - b["y"]).max() < 1e-11 and np.abs(a["lFex"] - b["lFex"]).max() < 1e-11
        if corr == "mcmc":
            assert float(a["bias"].reshape(-1)[0]) == float(b["bias"].reshape(-1)[0]) > 100       # same rejections ...
            assert np.array_equal(a["worst"], b["worst"])                                         # ... and run-length histogram
        else:
            np.testing.assert_allclose(a["bias"], b["bias"], rtol=1e-11)
            np.testing.assert_allclose(a["worst"], b["worst"], rtol=1e-11)


@pytest.mark.gpu
@pytest.mark.skipif(not (mex_host.gateway_available("b200") and (os.path.exists(matlab_driver.AST_FILE) or matlab_driver.reference_available())),
                    reason="oracle/_ref artefacts not built (they come from the build container)")
def test_reference_matlab_driver_runs_on_the_drop_in_library():
    if tt_irt.device_count() < 1:
        pytest.fail("no CUDA device: the -m gpu tests need a B200 (there is no CPU fallback)")
    g = dict(np.load(GOLDEN))
    compiled = matlab_driver.load_compiled()
    for corr in ("mcmc", "iw"):
        r = matlab_driver.run_debias("b200", corr, compiled)
        assert np.isfinite(r["y"]).all() and r["y"].shape == g[corr + "_y"].shape
        assert np.abs(r["y"] - g[corr + "_y"]).max() < 1e-9 and np.abs(r["lFex"] - g[corr + "_lFex"]).max() < 1e-9
        if corr == "mcmc":
            assert float(r["bias"].reshape(-1)[0]) == float(g["mcmc_bias"].reshape(-1)[0])
            assert np.array_equal(r["worst"], g["mcmc_worst"])
        else:
            np.testing.assert_allclose(r["bias"], g["iw_bias"], rtol=1e-10)
            np.testing.assert_allclose(r["worst"], g["iw_worst"], rtol=1e-10)
