"""CPU tests of bench.py's contract: the reference arm runs without a GPU and prints the line the driver parses; the GPU arm
refuses to run without a CUDA device (there is no CPU fallback to time)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, env=e, stdout=subprocess.PIPE,
                          stderr=subprocess.PIPE, text=True, timeout=900)


def test_reference_arm_line():
    """`bench.py --impl reference`: the unmodified reference (oracle/_ref) on the host cores, one JSON line with the keys the
    driver reads; other ranks of a torchrun launch print nothing."""
    out = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert out.returncode == 0, out.stderr[-500:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "ML-KEM-768 Encaps+Decaps ops/s" and d["unit"] == "encaps+decaps pairs/s"
    assert d["higher_is_better"] is True and d["gpu_launches"] == 0 and d["value"] > 0 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] == (os.cpu_count() or 1)
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "2^22 items per GPU" in d["config"]["workload"] and "sample" in d["config"]
    other = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0", env={"RANK": "1", "WORLD_SIZE": "2"})
    assert other.returncode == 0 and other.stdout.strip() == ""


def test_gpu_arm_refuses_without_a_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    out = run_bench("--steps", "1", "--warmup", "3", "--log2-items", "10")
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)
