"""CPU tests of bench.py's contract: the reference arm runs on the host cores and prints one JSON line with the keys the
driver reads; the B200 arm refuses to run without a device (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=600, env=e)


def test_reference_arm_prints_the_contract_line():
    out = _run(["--impl", "reference", "--steps", "1", "--warmup", "1", "--shape", "6,17,8", "--log2m", "14", "--cpu-samples-per-core", "128"])
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "irt_samples_per_sec" and line["unit"] == "samples/s"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["value"] == line["value"] == line["e2e"]["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["config"]["d"] == 6 and "workload" in line["config"]


def test_reference_arm_other_ranks_exit_quietly():
    out = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"], env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_b200_arm_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    out = _run(["--steps", "1", "--warmup", "1", "--shape", "4,9,4", "--log2m", "10"])
    assert out.returncode != 0
    assert "no CUDA device" in (out.stderr + out.stdout)
