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


def test_b200_arm_workload_is_built_without_the_oracle():
    """The B200 arm's model and inputs come from the product package alone (oracle/ is the checker, never part of the
    measured path), and they are the configuration the oracle-timed CPU legs use: same state_dict layout, same
    library configuration struct, same input shapes."""
    import ast
    import importlib.util

    import torch

    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    for fn in (n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("run_b200", "build_b200_case", "model_roofline")):
        for node in ast.walk(fn):
            if isinstance(node, (ast.Import, ast.ImportFrom)):
                names = [a.name for a in node.names] + [getattr(node, "module", None) or ""]
                assert not any(n.split(".")[0] == "oracle" for n in names), (fn.name, names)

    model, x, c, t = bench.build_b200_case(2, "bf16", 0, torch.device("cpu"))
    assert not model.training and model.precision == "bf16" and model.num_params == 210626560
    assert tuple(x.shape) == (2, 34, 180, 360) and tuple(c.shape) == (2, 2, 180, 360) and t.tolist() == [3.0, 3.0]
    cfg, sd, x1, c1, t1 = bench.build_case(1)
    mine = model.state_dict()
    assert list(mine) == list(sd) and all(mine[k].shape == sd[k].shape for k in sd)
    assert x1.shape[1:] == x.shape[1:] and c1.shape[1:] == c.shape[1:] and float(t1[0]) == 3.0
    assert (model.min_time, model.max_time) == (0, 5)
