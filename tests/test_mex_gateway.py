"""The Matlab side of the drop-in boundary, exercised without Matlab: the reference's MEX gateway source
(matlab/utils/tt_irt_mex.c -- what `[Z, lPz] = tt_irt_mex(f.n, cell2mat(xsf), f.r, f.core, Z)` runs, install.m:169) is compiled
UNMODIFIED against a stand-in mex.h (oracle/mexstub/) twice by oracle/Makefile:
  * with the reference's own tt_irt1_int64.c  -> oracle/_ref/libref_tt_irt_mex.so       (the reference end to end),
  * linked against tt-irt_b200/lib/libtt_irt1_int64.so instead -> ..._b200.so            (install.m:160 with its second source
    replaced by -ltt_irt1_int64: the swap INTEGRATION.md describes).
oracle/mex_host.py hands both the arguments as Matlab does (all double arrays; the gateway converts n and ttrank to mwIndex).
CPU: the reference gateway reproduces the C oracle bit for bit, the product-linked gateway loads, resolves tt_irt1 from the product
library and, without a device, NaN-fills.  GPU: the two gateways agree inside the parity protocol and the product-linked one is
bit-identical to the Python path of the same library.  The .so files are built in the build container (the gateway source lives
under /root/reference) and travel to the GPU box; where they are absent the tests skip."""
import os
import subprocess

import numpy as np
import pytest

from oracle import mex_host
from tt_irt_py import synth, tt_irt

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
needs_gateways = pytest.mark.skipif(not (mex_host.gateway_available("reference") and mex_host.gateway_available("b200")),
                                    reason="oracle/_ref/libref_tt_irt_mex*.so not built (needs /root/reference: build container only)")


def _case(seed=4, d=6, n=17, r=8, M=700):
    ns, xs, rk, c = synth.make_tt(d, n, r, seed=seed)
    return ns, xs, rk, c, synth.make_q(M, d, seed=seed + 1)


@needs_gateways
def test_reference_gateway_reproduces_the_oracle_bit_for_bit(oracle_mod):
    ns, xs, rk, c, q = _case()
    Z, l = mex_host.tt_irt_mex("reference", ns, xs, rk, c, q)
    Zo, lo = oracle_mod.oracle_run(ns, xs, rk, c, q)[:2]
    assert Z.shape == q.shape and np.array_equal(Z, Zo) and np.array_equal(l, lo)


@needs_gateways
def test_product_linked_gateway_resolves_the_drop_in_symbol():
    so = os.path.join(ROOT, "oracle", "_ref", "libref_tt_irt_mex_b200.so")
    undefined = subprocess.run(["nm", "-D", "--undefined-only", so], capture_output=True, text=True).stdout
    assert " tt_irt1" in undefined and "dgemm" not in undefined            # the gateway needs the one symbol and no BLAS
    ldd = subprocess.run(["ldd", so], capture_output=True, text=True).stdout
    assert "libtt_irt1_int64.so" in ldd and "not found" not in ldd
    if tt_irt.device_count() < 1:
        ns, xs, rk, c, q = _case(M=64)
        Z, l = mex_host.tt_irt_mex("b200", ns, xs, rk, c, q)              # no device: one stderr line, NaN-filled outputs, no crash
        assert Z.shape == q.shape and np.isnan(Z).all() and np.isnan(l).all()


@pytest.mark.gpu
@needs_gateways
def test_matlab_gateway_on_the_drop_in_library_matches_the_reference_gateway(oracle_mod):
    if tt_irt.device_count() < 1:
        pytest.fail("no CUDA device: the -m gpu tests need a B200 (there is no CPU fallback)")
    for (seed, d, n, r, M) in [(4, 6, 17, 8, 3000), (7, 5, 33, 32, 1500), (9, 4, 65, 64, 1000)]:
        ns, xs, rk, c, q = _case(seed, d, n, r, M)
        Zr, lr = mex_host.tt_irt_mex("reference", ns, xs, rk, c, q)
        Zb, lb = mex_host.tt_irt_mex("b200", ns, xs, rk, c, q)
        _, _, io, kap, gap, cond, lsens = oracle_mod.oracle_run(ns, xs, rk, c, q, extras=True)
        stats, fails = oracle_mod.parity.compare(Zb, lb, None, Zr, lr, None, cond, gap, lsens=lsens)
        assert not fails, (fails, stats)
        Zp, lp = tt_irt.run_host(ns, xs, rk, c, q)                          # the same library through its Python mirror
        assert np.array_equal(Zb, Zp) and np.array_equal(lb, lp)


# ---- the gateway INTEGRATION.md proposes for the squared-density transform (examples/tt_irt_sqr_mex.c) ----
def _build_sqr_gateway(tmp_path):
    so = str(tmp_path / "tt_irt_sqr_mex.so")
    lib64 = os.path.join(ROOT, "tt-irt_b200", "lib")
    stub = os.path.join(ROOT, "oracle", "mexstub")
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-Wl,-Bsymbolic", "-I" + stub, os.path.join(ROOT, "examples", "tt_irt_sqr_mex.c"),
           os.path.join(stub, "mexstub.c"), "-L" + lib64, "-ltt_irt1_int64", "-Wl,-rpath," + lib64, "-lm", "-o", so]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    return so


def test_sqr_gateway_example_compiles_and_links(tmp_path):
    so = _build_sqr_gateway(tmp_path)
    undefined = subprocess.run(["nm", "-D", "--undefined-only", so], capture_output=True, text=True).stdout
    assert " tt_irt_sqr" in undefined
    if tt_irt.device_count() < 1:
        ns, xs, rk, c = synth.make_tt(3, 5, 3, seed=1)
        xq, lF = mex_host.call_mex(so, [ns, xs, rk, c, synth.make_q(32, 3, seed=2)], 2)
        assert xq.shape == (32, 3) and np.isnan(xq).all() and np.isnan(lF).all()     # no device: NaN-filled, no crash


@pytest.mark.gpu
def test_sqr_gateway_example_matches_the_reference_matlab_source(tmp_path):
    """examples/tt_irt_sqr_mex.c on the B200 library against the outputs of the reference's tt_irt_sqr.m executed from source
    (tests/golden/matlab_sqr_*.npz): cores with and without boundary nodes, a marginal (q with fewer columns)."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_golden_matlab as gen
    from oracle import parity
    from oracle.tt_irt_sqr_oracle import tt_irt_sqr_oracle
    if tt_irt.device_count() < 1:
        pytest.fail("no CUDA device: the -m gpu tests need a B200 (there is no CPU fallback)")
    so = _build_sqr_gateway(tmp_path)
    for case in gen.SQR_CASES[:5]:
        g = dict(np.load(os.path.join(ROOT, "tests", "golden", "matlab_sqr_%s.npz" % case[0])))
        ns, xs, rk, c, q = gen.sqr_inputs(case)
        xq, lF = mex_host.call_mex(so, [ns, xs, rk, c, q], 2)
        _, _, io, cond, gap, lsens = tt_irt_sqr_oracle(ns, xs, rk, c, q, extras=True)
        st, fails = parity.compare(xq, lF.reshape(-1), None, g["xq"], g["lFapp"], None, cond, gap, lsens)
        assert not fails, (case[0], fails, st)
