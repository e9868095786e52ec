import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tt-irt_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
os.environ.setdefault("TTIRT_QUIET", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    if not os.path.exists(os.path.join(oracle.ORACLE_DIR, "liboracle_tt_irt1.so")):
        oracle.build()
    return oracle


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


def load_golden(golden_dir, name):
    import hashlib
    import numpy as np
    from tt_irt_py import synth
    g = dict(np.load(os.path.join(golden_dir, name + ".npz"), allow_pickle=False))
    if "cores" not in g:
        ns = g["n"]
        d = int(ns.size)
        _, _, _, c = synth.make_tt(d, int(ns[0]), int(g["ranks"][1]), seed=int(g["seed"]), lo=float(g["lo"]), hi=float(g["hi"]),
                                   grid=str(g["grid_kind"]), cores=str(g["cores_kind"]))
        if hashlib.sha256(c.tobytes()).hexdigest() != str(g["cores_sha256"]):
            raise RuntimeError("regenerated cores of golden %s do not match the recorded sha256 (numpy RNG drift)" % name)
        g["cores"] = c
    return g


GOLDEN = ["tiny_d3_n5_r3", "shock_d8_n17_r8", "shock_d8_n17_r16_cheb", "diffusion_d11_n17_r16", "signed_d6_n20_r12",
          "lorenz_d40_n33_r32_i64", "roofline_d32_n65_r64"]
