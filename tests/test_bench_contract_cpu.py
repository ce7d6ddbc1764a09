"""bench.py contract checks that need no GPU: the reference arm prints one well-formed JSON line, and the B200 arm
refuses to run (loudly) where there is no GPU instead of falling back to anything."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, cwd=ROOT,
                          env=e, timeout=600)


def test_reference_arm_prints_the_contract_line():
    r = _run("--impl", "reference", "--workload", "cfg1", "--steps", "1", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "pool pixels scored/s" and d["unit"] == "Gpix/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["gpu_launches"] == 0
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "Gpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("cfg1")


def test_reference_arm_other_ranks_exit_quietly():
    r = _run("--impl", "reference", "--workload", "cfg1", "--steps", "1", "--warmup", "1", "--gpus", "2", env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_b200_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = _run("--steps", "1", "--warmup", "1", "--workload", "cfg1")
    assert r.returncode != 0
    assert "GPU" in (r.stderr + r.stdout)
    assert not any(l.startswith("{") for l in r.stdout.splitlines()), "no bench line without a GPU"
