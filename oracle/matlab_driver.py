"""TEST INFRASTRUCTURE ONLY: the reference's top-level Matlab sampler driver, matlab/samplers/tt_irt_debias.m, run end to end
without Matlab -- seeds -> inverse Rosenblatt transform -> exact density -> Metropolis-Hastings / importance-weight correction
(tt_irt_debias.m:31-71) -- by the interpreter oracle/mlite.py, in three variants of its sampler call (:44):

  "matlab"     as the file is written: tt_irt_lin.m (the Matlab implementation, with the tracemult MEX compiled from its source);
  "reference"  as matlab/install.m:160-169 patches it once the MEX file is built: tt_irt_mex(f.n, cell2mat(xsf), f.r, f.core, Z),
               served by the reference's gateway tt_irt_mex.c + the reference's tt_irt1_int64.c (oracle/_ref/libref_tt_irt_mex.so);
  "b200"       the same patched call, served by the same unmodified gateway linked against tt-irt_b200/lib/libtt_irt1_int64.so:
               the reference's own driver on the drop-in library.

The Matlab sources live under /root/reference, which does not exist on the GPU box.  `python -m oracle.matlab_driver --compile`
(oracle/Makefile) parses them HERE into the interpreter's syntax trees and leaves those, like every other artefact built from
reference sources, under oracle/_ref/ (git-ignored, travels to the GPU box): oracle/_ref/matlab_debias.pkl.
What is not the reference's: `core2cell` (TT-Toolbox, a third-party dependency the reference does not vendor) is supplied here, and
the exact density is a closed-form stand-in defined in DRIVER below (the reference's drivers pass their own model's).
"""
import os
import pickle
import sys

import numpy as np

from . import mex_host, mlite

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_M = "/root/reference/matlab/samplers"
AST_FILE = os.path.join(_HERE, "_ref", "matlab_debias.pkl")
CALL_AS_WRITTEN = "tt_irt_lin(xsf, f, Z)"                                           # tt_irt_debias.m:44
CALL_PATCHED = "tt_irt_mex(f.n, cell2mat(xsf), f.r, f.core, Z)"                     # install.m:169

# our own few lines of Matlab around the reference driver: the exact log-density (column 1) and a quantity of interest (column 2)
DRIVER = """
function [y, lFex, bias, worst] = run_debias(M, f, xsf, correction)
lFfun = @(x) [0.8*sum(x, 2) - 0.5*sum(x.^2, 2), x(:,1)];
[y, lFex, bias, tinv, worst] = tt_irt_debias(M, lFfun, f, xsf, correction);
end
"""


def reference_available():
    return os.path.isdir(REF_M) and mex_host.available()


def _parse(src):
    fs = mlite.Parser(src).parse_file()
    fs.pop("__first__")
    return fs


def compile_reference():
    """Parse the reference sources the driver needs into syntax trees: {'as_written': {...}, 'patched': {...}}."""
    out = {"as_written": {}, "patched": {}}
    for name in ("mcmc_prune", "iw_prune"):
        with open(os.path.join(REF_M, name + ".m")) as fh:
            fs = _parse(fh.read())
        out["as_written"].update(fs)
        out["patched"].update(fs)
    with open(os.path.join(REF_M, "tt_irt_lin.m")) as fh:
        out["as_written"].update(_parse(fh.read()))
    with open(os.path.join(REF_M, "tt_irt_debias.m")) as fh:
        src = fh.read()
    if src.count(CALL_AS_WRITTEN) != 1:
        raise RuntimeError("tt_irt_debias.m no longer contains the call %r exactly once" % CALL_AS_WRITTEN)
    out["as_written"].update(_parse(src))
    out["patched"].update(_parse(src.replace(CALL_AS_WRITTEN, CALL_PATCHED)))
    return out


def load_compiled():
    if os.path.exists(AST_FILE):
        with open(AST_FILE, "rb") as fh:
            return pickle.load(fh)
    if reference_available():
        return compile_reference()
    return None


def _core2cell(ip, args, nargout):
    """TT-Toolbox's core2cell (third-party, not vendored by the reference): tt_tensor -> d x 1 cell of r_k x n_k x r_{k+1} cores."""
    f = args[0]
    n = f["n"].reshape(-1).astype(int)
    r = f["r"].reshape(-1).astype(int)
    c = f["core"].reshape(-1)
    out = mlite.MCell((len(n), 1))
    off = 0
    for k in range(len(n)):
        sz = r[k] * n[k] * r[k + 1]
        out.a[k, 0] = np.asfortranarray(c[off:off + sz].reshape((r[k], n[k], r[k + 1]), order="F"))
        off += sz
    return out


def inputs(d=6, n=17, r=8, M=2000):
    """Seeded TT, seeds and the uniforms mcmc_prune draws (tt-irt_b200/tt_irt_py/synth.py generators)."""
    sys.path.insert(0, os.path.join(os.path.dirname(_HERE), "tt-irt_b200"))
    from tt_irt_py import synth
    ns, xs, rk, c = synth.make_tt(d, n, r, seed=21)
    rng = np.random.default_rng(5)
    return ns, xs, rk, c, np.asfortranarray(rng.random((M, d))), rng.random(M)


def run_debias(variant, correction, compiled=None):
    """[y, lFex, bias, worst] of the reference driver; variant 'matlab' | 'reference' | 'b200', correction 'mcmc' | 'iw'."""
    compiled = compiled or load_compiled()
    if compiled is None:
        raise RuntimeError("neither /root/reference nor oracle/_ref/matlab_debias.pkl is available")
    ns, xs, rk, c, Z, us = inputs()
    d, n = len(ns), int(ns[0])
    it = iter(us)
    ext = {"core2cell": _core2cell}
    if variant == "matlab":
        ext["tracemult"] = lambda ip, a, no: mex_host.ref_tracemult(*a)
        funcs = compiled["as_written"]
    else:
        ext["tt_irt_mex"] = lambda ip, a, no: list(mex_host.call_mex(mex_host._gateway(variant), a, 2))
        funcs = compiled["patched"]
    ip = mlite.Interp(rand_stream=lambda shape: np.array([[next(it)]]), externals=ext)
    ip.funcs.update(funcs)
    ip.load_source(DRIVER)
    xsf = mlite.MCell((d, 1))
    for k in range(d):
        xsf.a[k, 0] = np.asarray(xs[k * n:(k + 1) * n], dtype=np.float64).reshape(-1, 1)
    f = {"__class__": "tt_tensor", "d": np.array([[float(d)]]), "n": ns.astype(np.float64).reshape(-1, 1),
         "r": rk.astype(np.float64).reshape(-1, 1), "core": np.asarray(c, dtype=np.float64).reshape(-1, 1)}
    y, lFex, bias, worst = ip.call("run_debias", [Z, f, xsf, mlite.MStr(correction)], 4)
    return {"y": np.asarray(y), "lFex": np.asarray(lFex), "bias": np.asarray(bias, dtype=np.float64), "worst": np.asarray(worst, dtype=np.float64)}


if __name__ == "__main__":
    if "--compile" in sys.argv:
        os.makedirs(os.path.dirname(AST_FILE), exist_ok=True)
        with open(AST_FILE, "wb") as fh:
            pickle.dump(compile_reference(), fh)
        print("wrote", AST_FILE)
