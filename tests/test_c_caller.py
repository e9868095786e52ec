"""Plain-C programs calling the drop-in symbol (examples/call_tt_irt1.c) and the squared-density transforms
(examples/call_tt_irt_sqr.c): they compile and link against both libraries here (CPU); on a B200 they run and must agree with
the Python path on the same xorshift-generated inputs."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "examples", "call_tt_irt1.c")
SO32 = os.path.join(ROOT, "tt-irt_b200", "tt_irt_py", "tt_irt1_int32.so")
LIB64 = os.path.join(ROOT, "tt-irt_b200", "lib")


SRC_SQR = os.path.join(ROOT, "examples", "call_tt_irt_sqr.c")


def _build(tmp_path, width, src=None):
    global SRC
    keep = SRC
    if src is not None:
        SRC = src
    try:
        return _build_one(tmp_path, width)
    finally:
        SRC = keep


def _build_one(tmp_path, width):
    exe = str(tmp_path / ("call%d" % width))
    if width == 32:
        cmd = ["gcc", "-O2", "-DTTIRT_INT=int", SRC, "-I" + os.path.join(ROOT, "include"), SO32, "-lm", "-o", exe]
    else:
        cmd = ["gcc", "-O2", "-DTTIRT_INT=long long", SRC, "-I" + os.path.join(ROOT, "include"), "-L" + LIB64, "-ltt_irt1_int64",
               "-Wl,-rpath," + LIB64, "-lm", "-o", exe]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    return exe


@pytest.mark.parametrize("width", [32, 64])
def test_c_caller_compiles_and_links(tmp_path, width):
    exe = _build(tmp_path, width)
    assert os.path.exists(exe)
    # without a device the library refuses loudly and NaN-fills: the program's own check fails, nothing crashes
    from tt_irt_py import tt_irt
    if tt_irt.device_count() < 1:
        env = dict(os.environ, LD_LIBRARY_PATH=os.path.dirname(SO32))
        env.pop("TTIRT_QUIET", None)
        out = subprocess.run([exe, "64"], capture_output=True, text=True, env=env)
        assert out.returncode == 1 and "sumZ=nan" in out.stdout and "no CUDA device" in out.stderr


def _xorshift_inputs(ncore, nq):
    s = 88172645463325252
    mask = (1 << 64) - 1
    out = np.empty(ncore + nq)
    for i in range(ncore + nq):
        s ^= (s << 13) & mask; s ^= s >> 7; s ^= (s << 17) & mask
        out[i] = (s >> 11) / 9007199254740992.0
    return out[:ncore], out[ncore:]


@pytest.mark.gpu
@pytest.mark.parametrize("width", [32, 64])
def test_c_caller_agrees_with_the_python_path(tmp_path, width):
    from tt_irt_py import tt_irt
    exe = _build(tmp_path, width)
    M, d, nn, r = 2048, 6, 17, 8
    out = subprocess.run([exe, str(M)], capture_output=True, text=True, env=dict(os.environ, LD_LIBRARY_PATH=os.path.dirname(SO32)))
    assert out.returncode == 0, out.stdout + out.stderr
    m = re.search(r"sumZ=(\S+) sumlPz=(\S+) outside=(\d+) launches=(\d+)", out.stdout)
    sz, sl, outside, launches = float(m.group(1)), float(m.group(2)), int(m.group(3)), int(m.group(4))
    assert outside == 0 and launches > 0
    rk = np.array([1] + [r] * (d - 1) + [1]); ns = np.full(d, nn)
    ncore = int((rk[:-1] * ns * rk[1:]).sum())
    core, qf = _xorshift_inputs(ncore, M * d)
    xs = np.tile(-1.0 + 2.0 * np.arange(nn) / (nn - 1), d)
    q = qf.reshape((M, d), order="F")
    md = tt_irt.Model(ns, xs, rk, core)
    try:
        Z, l = md.sample(q)
    finally:
        md.close()
    # same library, same inputs: the column-wise sums agree to summation-order rounding
    assert abs(Z.sum() - sz) <= 1e-9 * max(1.0, abs(sz)) and abs(l.sum() - sl) <= 1e-9 * max(1.0, abs(sl))


@pytest.mark.parametrize("width", [32, 64])
def test_sqr_c_caller_compiles_and_links(tmp_path, width):
    exe = _build(tmp_path, width, SRC_SQR)
    from tt_irt_py import tt_irt
    if tt_irt.device_count() < 1:
        env = dict(os.environ, LD_LIBRARY_PATH=os.path.dirname(SO32))
        env.pop("TTIRT_QUIET", None)
        out = subprocess.run([exe, "64"], capture_output=True, text=True, env=env)
        assert out.returncode == 1 and "sumZ=nan" in out.stdout and "no CUDA device" in out.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("width", [32, 64])
def test_sqr_c_caller_round_trips_and_agrees_with_the_python_path(tmp_path, width):
    from tt_irt_py import tt_irt_sqr
    exe = _build(tmp_path, width, SRC_SQR)
    M, d, nn, r = 2048, 5, 17, 8
    out = subprocess.run([exe, str(M)], capture_output=True, text=True, env=dict(os.environ, LD_LIBRARY_PATH=os.path.dirname(SO32)))
    assert out.returncode == 0, out.stdout + out.stderr
    m = re.search(r"sumZ=(\S+) sumlF=(\S+) outside=(\d+) roundtrip=(\S+) dlF=(\S+) launches=(\d+)", out.stdout)
    sz, sl, outside, rt, dl, launches = float(m.group(1)), float(m.group(2)), int(m.group(3)), float(m.group(4)), float(m.group(5)), int(m.group(6))
    assert outside == 0 and launches > 0 and rt < 1e-11 and dl < 1e-10
    rk = np.array([1] + [r] * (d - 1) + [1]); ns = np.full(d, nn)
    ncore = int((rk[:-1] * ns * rk[1:]).sum())
    core, qf = _xorshift_inputs(ncore, M * d)
    xs = np.tile(-1.0 + 2.0 * np.arange(nn) / (nn - 1), d)
    q = qf.reshape((M, d), order="F")
    md = tt_irt_sqr.SqrModel(ns, xs, rk, core)
    try:
        Z, l = md.sample(q)
    finally:
        md.close()
    assert abs(Z.sum() - sz) <= 1e-9 * max(1.0, abs(sz)) and abs(l.sum() - sl) <= 1e-9 * max(1.0, abs(sl))
