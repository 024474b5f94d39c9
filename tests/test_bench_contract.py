"""bench.py contract checks that need no GPU: the reference arm prints ONE JSON line with the contract's keys, and the native arm
fails loudly (no number, non-zero exit) when there is no CUDA device - there is no CPU fallback."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], capture_output=True, text=True, timeout=600, env=e, cwd=ROOT)


def test_reference_arm_prints_one_contract_line(ref):
    r = run("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "fitness evals/sec (CEC2014 D=100)" and d["unit"] == "evals/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("cec2014 f1-f30") and d["dtype"] == "f64" and d["vs_baseline"] is None


def test_reference_arm_other_ranks_exit_quietly(ref):
    r = run("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_native_arm_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = run("--steps", "1", "--warmup", "3", "--e2e-steps", "0", "--no-cpu-baseline", "--no-secondary")
    assert r.returncode != 0
    assert not any(ln.strip().startswith("{") and '"value"' in ln for ln in r.stdout.splitlines())
