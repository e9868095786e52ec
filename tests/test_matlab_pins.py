"""CPU tests: the numpy restatements under oracle/ of the reference's MATLAB-only routines, pinned against outputs of the
reference's own source files executed here (tests/golden/matlab_*.npz, generator tests/golden/make_golden_matlab.py).

The reference sources run under oracle/mlite.py (an interpreter for the Matlab subset they use; numpy / LAPACK numerics) and
the MEX function tracemult.c runs as the reference's own C compiled against a stand-in mex.h.  Three layers:
  1. the interpreter itself against Matlab semantics worked out by hand (indexing, growth, ranges, N-d arrays, precedence), and
     against an independent ground truth: the reference's Matlab implementation of tt_irt1's transform (tt_irt_lin.m), run by the
     interpreter, lands on the results of the reference's compiled C routine;
  2. every oracle function against the committed fixtures (always; needs nothing but the repo);
  3. when /root/reference is present (the build container): the generation is repeated and must give the committed bits, so the
     fixtures are what the reference's sources produce today and not a stale or hand-made file.
Bars: exact where the operations are exact (lattice seeds, uniform pass-through, Metropolis-Hastings decisions, tracemult's
column pick), 1e-12 relative elsewhere (a real Matlab run differs from numpy in summation order and libm by less than that).
"""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden_matlab as gen  # noqa: E402

from oracle import mlite, samplers_oracle as so  # noqa: E402
from oracle.dirt_oracle import tt_dirt_inverse_oracle, tt_dirt_sample_oracle  # noqa: E402
from oracle.tt_irt_sqr_oracle import tracemult_oracle, tt_irt_sqr_oracle, tt_rt_sqr_oracle  # noqa: E402

GOLDEN = os.path.join(HERE, "golden")


def _load(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))


needs_reference_early = pytest.mark.skipif(not gen.available(), reason="needs /root/reference and oracle/_ref/libref_tracemult.so (build container only)")


# ---------------------------------------------------------------------------------------------------------------------
# 1. the interpreter
# ---------------------------------------------------------------------------------------------------------------------
def _run(src, fname, args, nargout=1, **kw):
    ip = mlite.Interp(**kw)
    ip.load_source(src)
    return ip.call(fname, args, nargout)


def test_mlite_indexing_growth_and_ranges():
    src = """
function [a, b, c, e, g] = t(x)
a = x(2:end)' * 2;                      % end in an index, transpose, scalar product
b = zeros(2,3); b(2,:) = [1, 2, 3]; b(:,1) = [7; 8];
c = cell(2,1); c{2} = zeros(1,4); c{2}(2:3) = [5, 6];     % nested left-hand side
e = 10:-3:1;                            % [10 7 4 1]
g = zeros(1,1); g(3,1) = 4;             % growth by assignment (mcmc_prune.m:39)
end
"""
    a, b, c, e, g = _run(src, "t", [np.array([[1.0, 2.0, 3.0]])], 5)
    assert np.array_equal(a, [[4.0], [6.0]])
    assert np.array_equal(b, [[7, 0, 0], [8, 2, 3]])
    assert np.array_equal(c.a[1, 0], [[0, 5, 6, 0]])
    assert np.array_equal(e, [[10, 7, 4, 1]])
    assert np.array_equal(g, [[0], [0], [4]])


def test_mlite_precedence_logic_and_control_flow():
    src = """
function [a, b, y, k, m] = t(n)
a = -2^2 + 3*4 - 8/2/2;                 % -4 + 12 - 2
b = [1 < 2 & 3 > 4, ~(1 == 1) | 2 ~= 3, 1:3 == [1, 5, 3]];
y = 0;
for i = 1:n
  if (i == 3)
    continue
  elseif i > 5
    break
  end
  y = y + i^2;
end
k = 0;
while k < 3, k = k + 1; end
v = [4, 8, 15, 16, 23, 42];
m = v(v > 10 & v < 30);                 % logical mask keeps the orientation of a row vector
end
"""
    a, b, y, k, m = _run(src, "t", [np.array([[10.0]])], 5)
    assert float(a[0, 0]) == 6.0
    assert np.array_equal(b.astype(float), [[0, 1, 1, 0, 1]])
    assert float(y[0, 0]) == 1 + 4 + 16 + 25 and float(k[0, 0]) == 3
    assert np.array_equal(m, [[15, 16, 23]])


def test_mlite_nd_arrays_cells_and_handles():
    src = """
function [s, p, r, q, w] = t(f)
n = cellfun(@(x)size(x,2), f);          % tt_irt_sqr.m:27
s = [1; n];
A = reshape(1:24, 2, 3, 4);
p = permute(A, [1,3,2]);                % 2 x 4 x 3
r = reshape(A(:,2,:), 2, []);           % A(:,2,:) is 2 x 1 x 4
B = zeros(2,3,2); B(:,2:3,:) = A(:,1:2,3:4); B(:,1,:) = B(:,2,:) - B(:,3,:);
q = B;
w = A .* reshape([1, 10, 100, 1000], 1, 1, 4);   % implicit expansion of singleton dimensions
end
"""
    f = mlite.MCell((2, 1))
    f.a[0, 0] = np.zeros((1, 5, 3))
    f.a[1, 0] = np.zeros((3, 7))
    s, p, r, q, w = _run(src, "t", [f], 5)
    A = np.arange(1, 25, dtype=float).reshape((2, 3, 4), order="F")
    assert np.array_equal(s, [[1], [5], [7]])
    assert np.array_equal(p, np.transpose(A, (0, 2, 1)))
    assert np.array_equal(r, A[:, 1, :])
    B = np.zeros((2, 3, 2)); B[:, 1:3, :] = A[:, 0:2, 2:4]; B[:, 0, :] = B[:, 1, :] - B[:, 2, :]
    assert np.array_equal(q, B)
    assert np.array_equal(w, A * np.array([1.0, 10, 100, 1000]).reshape(1, 1, 4))


def test_mlite_strings_varargin_and_multiple_outputs():
    src = """
function [s, t] = top(name, varargin)
[s, t] = helper(name, varargin{:});
end
function [a, b] = helper(name, x, y)
a = lower(name(1)) ~= 'u';
v = double(lower(name)); v = v((v==46) | ((v>=48) & (v<=57)));     % randref.m:25-27
b = [str2double(char(v)), x + y, nargin];
end
"""
    s, t = _run(src, "top", [mlite.MStr("Normal 3.5"), 2.0, 5.0], 2)
    assert bool(s[0, 0]) and np.array_equal(t, [[3.5, 7.0, 3.0]])
    s, t = _run(src, "top", [mlite.MStr("UNI"), 1.0, 1.0], 2)
    assert not bool(s[0, 0]) and np.isnan(t[0, 0])


def test_mlite_refuses_what_it_does_not_know():
    with pytest.raises(mlite.MatlabError):
        _run("function y = t(x)\ny = svd(x);\nend\n", "t", [np.eye(2)])
    with pytest.raises(mlite.MatlabError):
        _run("function y = t(x)\ny = x(5);\nend\n", "t", [np.zeros((2, 2))])
    with pytest.raises(mlite.MatlabError):
        _run("function y = t(x)\ny = x * x;\nend\n", "t", [np.zeros((2, 3))])


@pytest.mark.parametrize("case", gen.LIN_CASES, ids=[c[0] for c in gen.LIN_CASES])
def test_mlite_runs_the_matlab_tt_irt_lin_into_the_compiled_reference_c(oracle_mod, case):
    """The interpreter against an independent ground truth: the reference ships the SAME transform twice, as Matlab
    (matlab/samplers/tt_irt_lin.m, 150 lines: cells, N-d arrays, spdiags, cumsum, logical indexing, the tracemult MEX) and as C
    (tt_irt1_int32.c).  The Matlab file executed by oracle/mlite.py must land on the compiled C routine's results -- which the C
    oracle reproduces bit for bit (tests/test_oracle.py) -- inside the parity protocol (the two differ in their root formula and
    tt_irt_lin.m clamps x_k to its cell, tt_irt_lin.m:134-150)."""
    g = _load("matlab_lin_" + case[0])
    ns, xs, rk, c, q = gen.lin_inputs(case)
    Zo, lo, io, kap, gap, cond, lsens = oracle_mod.oracle_run(ns, xs, rk, c, q, extras=True)
    stats, fails = oracle_mod.parity.compare(g["xq"], g["lFapp"], None, Zo, lo, None, cond, gap, lsens=lsens)
    assert not fails, (fails, stats)
    assert np.abs(g["xq"] - Zo).max() < 1e-11 and np.abs(g["lFapp"] - lo).max() < 1e-11


# ---------------------------------------------------------------------------------------------------------------------
# 2. oracles against the committed outputs of the reference sources
# ---------------------------------------------------------------------------------------------------------------------
def test_consumers_against_the_reference_matlab_sources():
    """essinv.m:12-14, hellinger.m:12-16, iw_prune.m:19-29."""
    g, inp = _load("matlab_helpers"), gen.helper_inputs()
    lFex, lFapp = inp["lFex"][:, 0], inp["lFapp"][:, 0]
    np.testing.assert_allclose(so.essinv(lFex, lFapp), float(g["essinv"].reshape(-1)[0]), rtol=1e-13)
    np.testing.assert_allclose(so.hellinger(lFex, lFapp), float(g["hellinger"].reshape(-1)[0]), rtol=1e-13)
    w, isstd, mx, err1, _ = so.iw_prune(lFex, lFapp)
    np.testing.assert_allclose(isstd, float(g["iw_isstd"].reshape(-1)[0]), rtol=1e-13)
    np.testing.assert_allclose(mx, float(g["iw_max_ratio"].reshape(-1)[0]), rtol=1e-13)
    np.testing.assert_allclose(err1, float(g["iw_err1"].reshape(-1)[0]), rtol=1e-13)
    np.testing.assert_allclose(inp["lFex"] * w[:, None], g["iw_lFex"], rtol=1e-13)          # lFex .* repmat(isstd, 1, 2), :28


def test_mcmc_prune_against_the_reference_matlab_source():
    """mcmc_prune.m:24-43 with a recorded rand stream: same surviving samples, rejection count and run-length histogram."""
    g, inp = _load("matlab_helpers"), gen.helper_inputs()
    src, nrej, dist = so.mcmc_prune(inp["mh_lFex"][:, 0], inp["mh_lFapp"][:, 0], inp["mh_u"])
    assert nrej == int(g["mh_num_of_rejects"].reshape(-1)[0]) and nrej > 500
    assert np.array_equal(inp["mh_y"][src], g["mh_y"])
    assert np.array_equal(inp["mh_lFex"][src], g["mh_lFex"])
    assert np.array_equal(inp["mh_lFapp"][src, 0], g["mh_lFapp"].reshape(-1))
    assert np.array_equal(dist, g["mh_rej_distribution"].reshape(-1).astype(np.int64))


def test_seeds_against_the_reference_matlab_sources():
    """qmcnodes.m:6-13 (bit-exact: every operation is exact or a single rounding) and randref.m:14-34 (rank 3)."""
    g, inp = _load("matlab_helpers"), gen.helper_inputs()
    Y = so.qmc_lattice(inp["d"], inp["l"], inp["table"][:, 1], inp["shift"][:, 0])
    assert np.array_equal(Y, g["qmc_Y"].T)                      # the reference returns d x N, tt_irt1 takes N x d
    for tag, sigma in (("normal", 4.0), ("normal3", 3.0), ("n2p5", 2.5)):
        np.testing.assert_allclose(so.truncnormal_map(inp["u"], sigma), g["randref_" + tag], rtol=1e-13, atol=1e-15)
    assert np.array_equal(g["randref_uni"], inp["u"])           # 'UNI...': the seeds pass through (:21-22)


def test_tracemult_against_the_reference_mex_source():
    """matlab/utils/tracemult.c compiled unmodified: batched product (:103-112) and column pick (:131-136)."""
    g = _load("matlab_tracemult")
    A, B, j, A2, j2 = gen.tracemult_inputs()
    np.testing.assert_allclose(tracemult_oracle(A, j[:, 0], B), g["C"], rtol=1e-14, atol=1e-14)
    assert np.array_equal(np.asarray(tracemult_oracle(A2, j2[:, 0])).reshape(-1), g["C2"].reshape(-1))


@pytest.mark.parametrize("case", gen.SQR_CASES, ids=[c[0] for c in gen.SQR_CASES])
def test_squared_density_transforms_against_the_reference_matlab_sources(case):
    """tt_irt_sqr.m:1-208 and tt_rt_sqr.m:1-178 (rank 4), the reference sources executed on seeded TTs: cores with and without
    boundary nodes, a marginal (fewer seed columns than dimensions), signed cores, seeds exactly 0 and 1.
    Z and the log-density to 1e-12 (relative to the grid span / absolute), the forward transform likewise."""
    g = _load("matlab_sqr_" + case[0])
    ns, xs, rk, c, q = gen.sqr_inputs(case)
    Z, lF = tt_irt_sqr_oracle(ns, xs, rk, c, q)       # (cores without boundary nodes: the grid is two points longer per dimension)
    assert Z.shape == g["xq"].shape
    np.testing.assert_allclose(Z, g["xq"], rtol=0, atol=1e-12 * 2.5)
    np.testing.assert_allclose(lF, g["lFapp"], rtol=1e-12, atol=1e-11)
    q2, lF2 = tt_rt_sqr_oracle(ns, xs, rk, c, g["xq"])
    np.testing.assert_allclose(q2, g["rt_q"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(lF2, g["rt_lFapp"], rtol=1e-12, atol=1e-11)
    # the reference's own round trip: tt_rt_sqr(tt_irt_sqr(q)) = q (away from the clamped seeds 0 and 1)
    np.testing.assert_allclose(g["rt_q"][2:], q[2:, :g["rt_q"].shape[1]], rtol=0, atol=1e-9)


@pytest.mark.parametrize("case", gen.DIRT_CASES, ids=[c[0] for c in gen.DIRT_CASES])
def test_dirt_loops_against_the_reference_matlab_sources(case):
    """tt_dirt_sample.m:17-73 and tt_dirt_inverse.m:24-59 (the callers of the squared-density transforms), reference sources
    executed on a synthetic two-level DIRT with a uniform and a truncated-normal reference.  The inverse under a normal
    reference goes through erfinv near its poles: 1e-9 there, 1e-12 elsewhere."""
    g = _load("matlab_dirt_" + case[0])
    levels, q, reference = gen.dirt_inputs(case)
    z, lF = tt_dirt_sample_oracle(levels, q, reference)
    np.testing.assert_allclose(z, g["z"], rtol=0, atol=5e-12)
    np.testing.assert_allclose(lF, g["lFapp"], rtol=1e-12, atol=1e-11)
    tol = 1e-12 if reference[0].lower() == "u" else 1e-9
    q2, lF2 = tt_dirt_inverse_oracle(levels, g["z"], reference)
    np.testing.assert_allclose(q2, g["inv_q"], rtol=0, atol=tol)
    np.testing.assert_allclose(lF2, g["inv_lFapp"], rtol=tol, atol=tol * 10)
    np.testing.assert_allclose(g["inv_q"], q, rtol=0, atol=1e-9)        # the reference's own round trip


@needs_reference_early
def test_mlite_runs_the_matlab_tracemult_fallback_into_the_compiled_mex():
    """A second independent ground truth for the interpreter: the reference ships tracemult twice as well, as a MEX C file
    (matlab/utils/tracemult.c, compiled unmodified here) and as its pure-Matlab fallback (matlab/utils/tracemultm.m)."""
    from oracle import mex_host
    ip = mlite.Interp()
    ip.load_file("/root/reference/matlab/utils/tracemultm.m")
    A, B, j, A2, j2 = gen.tracemult_inputs()
    np.testing.assert_allclose(ip.call("tracemultm", [A, j, B])[0], mex_host.ref_tracemult(A, j, B), rtol=1e-14, atol=1e-14)
    assert np.array_equal(ip.call("tracemultm", [A2, j2])[0], mex_host.ref_tracemult(A2, j2))


# ---------------------------------------------------------------------------------------------------------------------
# 3. provenance: with the reference present, the generation gives the committed bits
# ---------------------------------------------------------------------------------------------------------------------
needs_reference = pytest.mark.skipif(not gen.available(), reason="needs /root/reference and oracle/_ref/libref_tracemult.so (build container only)")


@needs_reference
def test_fixtures_are_what_the_reference_sources_produce_here():
    live = gen.run_helpers()
    g = _load("matlab_helpers")
    assert sorted(live) == sorted(g)
    for k in live:
        assert np.array_equal(live[k], g[k]), k
    live = gen.run_tracemult()
    g = _load("matlab_tracemult")
    for k in live:
        assert np.array_equal(live[k], g[k]), k
    for case in gen.SQR_CASES[:3]:
        live = gen.run_sqr(case)
        g = _load("matlab_sqr_" + case[0])
        for k in live:
            assert np.array_equal(live[k], g[k]), (case[0], k)
    live = gen.run_lin(gen.LIN_CASES[0])
    g = _load("matlab_lin_" + gen.LIN_CASES[0][0])
    for k in live:
        assert np.array_equal(live[k], g[k]), k
    live = gen.run_dirt(gen.DIRT_CASES[1])
    g = _load("matlab_dirt_" + gen.DIRT_CASES[1][0])
    for k in live:
        assert np.array_equal(live[k], g[k]), k
