"""bench.py contract on the CPU: the reference arm (`--impl reference`) prints ONE JSON line with the keys the driver
reads; the B200 arm refuses to run without a GPU instead of falling back."""
import json
import os
import subprocess
import sys

from conftest import ROOT

REQUIRED = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
            "dtype", "data", "config", "e2e", "cpu_baseline", "impl"]


def test_reference_arm_prints_one_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS=str(min(8, os.cpu_count() or 1)))
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, p.stdout[-2000:]
    r = json.loads(lines[0])
    for k in REQUIRED:
        assert k in r, k
    assert r["impl"] == "reference" and r["metric"] == "SFNO fwd samples/s" and r["unit"] == "samples/s"
    assert r["higher_is_better"] is True and r["vs_baseline"] is None and r["value"] > 0
    assert r["e2e"]["h2d_bytes_per_step"] == 0 and r["e2e"]["d2h_bytes_per_step"] == 0 and r["e2e"]["value"] == r["value"]
    assert r["cpu_baseline"]["kind"] in ("port", "reference") and r["cpu_baseline"]["cores"] >= 1
    assert "workload" in r["config"]


def test_b200_arm_has_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        import pytest

        pytest.skip("a GPU is present: the B200 arm would run")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1", "--no-cpu-baseline"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode != 0
    assert not any(ln.startswith("{") for ln in p.stdout.splitlines())
