"""Tensor-core engine (tcgen05 + TMA) vs CUDA-core engine on identical bf16 operands, per op and per operand
layout (K-major / MN-major SWIZZLE_128B descriptors, batched and shared operands, K / M / N tails)."""
import json
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu

OPS = {"dft": 0, "leg": 1, "dhconv": 2, "ileg": 3, "idft": 4, "conv": 5, "convb": 6}

# conv / idft epilogue bitmask: 1 bias, 2 gelu, 4 residual (+ per-row affine), 8 pos-embed, 16 dropout
CASES = [
    # op, dims, expect tensor-core engine
    ("leg", [1, 8, 16, 16, 9, 0], True),        # single tile, K = 16
    ("leg", [2, 64, 180, 180, 5, 0], True),     # ACE geometry, few m
    ("leg", [8, 256, 180, 180, 181, 0], True),  # ACE full size
    ("leg", [8, 256, 180, 180, 181, 1], True),  # ... storing only degrees l >= m (triangular)
    ("dft", [1, 8, 16, 32, 17, 0], True),
    ("dft", [2, 32, 180, 360, 181, 0], True),
    ("dft", [8, 256, 180, 360, 181, 0], True),
    ("dft", [1, 8, 18, 36, 19, 0], False),      # nlon*2 bytes not 16-byte aligned -> CUDA-core engine
    ("dhconv", [1, 16, 4, 5, 0, 0], True),
    ("dhconv", [2, 64, 20, 21, 0, 0], True),
    ("dhconv", [8, 256, 180, 181, 0, 0], True),
    ("dhconv", [8, 256, 180, 181, 1, 0], True),  # only wavenumbers m <= l (triangular)
    ("dhconv", [3, 32, 20, 21, 1, 0], True),
    ("ileg", [1, 8, 16, 16, 9, 0], True),
    ("ileg", [1, 8, 16, 16, 9, 1], True),       # X layout (residual path)
    ("ileg", [2, 64, 180, 180, 7, 0], True),
    ("ileg", [8, 256, 180, 180, 181, 0], True),
    ("ileg", [8, 256, 180, 180, 181, 2], True),  # contraction over l >= m only (triangular, Y layout)
    ("ileg", [2, 64, 180, 180, 181, 3], True),   # triangular, X layout
    ("idft", [1, 8, 16, 32, 17, 0], True),
    ("idft", [2, 16, 180, 360, 181, 7], True),
    ("idft", [8, 256, 180, 360, 181, 7], True),   # block epilogue: + bias + inner-skip -> GELU
    ("idft", [8, 256, 180, 360, 181, 0], True),   # residual path: plain inverse
    ("conv", [1, 16, 16, 288, 0, 0], True),
    ("conv", [2, 36, 256, 64800, 0, 3], True),    # fp32 output, K = 36 (tail), bias + GELU
    ("conv", [2, 292, 256, 64800, 0, 0], True),   # decoder0-like: K = 292 (tail)
    ("conv", [2, 256, 34, 64800, 0, 0], True),    # decoder1: cout = 34 rows, fp32 output
    ("conv", [2, 256, 64, 64800, 0, 5], False),   # fp32 output with a residual -> CUDA-core engine (not on the hot path)
    ("conv", [2, 64, 96, 1024, 0, 19], True),     # dropout epilogue: identical Philox masks in both engines
    ("convb", [2, 36, 256, 64800, 0, 3], True),   # encoder0
    ("convb", [2, 256, 256, 64800, 0, 8], True),  # encoder1: + pos-embed
    ("convb", [8, 256, 256, 64800, 1, 1], True),  # inner skip: per-sample folded weights + bias
    ("convb", [8, 256, 512, 64800, 1, 3], True),  # fc1 at ACE size: folded weights, bias, GELU
    ("convb", [8, 512, 256, 64800, 0, 5], True),  # fc2 at ACE size: bias + affine residual
    ("convb", [2, 512, 256, 64800, 0, 21], True), # fc2 with dropout (interpolator)
    ("convb", [2, 256, 256, 64800, 1, 15], True), # everything but dropout
    ("convb", [2, 512, 512, 4096, 0, 5], True),   # embed-512 widths (scaled configuration): inner skip / encoder
    ("convb", [1, 512, 1024, 2048, 1, 3], True),  # ... fc1 of the embed-512 model
    ("convb", [1, 16, 16, 300, 0, 0], False),     # hw not a multiple of 8 -> CUDA-core engine
]


# fp32 storage through the tensor cores as TF32 (kind::tf32) vs the CUDA-core FMA engine on TF32-exact operands: both
# compute the same products, only the fp32 summation order differs.  Same op kinds (+100 in the C entry point).
TF32_CASES = [
    ("leg", [1, 8, 16, 16, 9, 0], True),
    ("leg", [2, 64, 180, 180, 181, 1], True),      # triangular, dual-M tiles
    ("leg", [2, 64, 180, 180, 5, 0], True),
    ("dft", [1, 8, 16, 32, 17, 0], True),
    ("dft", [2, 32, 180, 360, 181, 0], True),       # K = 360: 11 full K blocks of 32 + one of 8
    ("dft", [1, 8, 18, 36, 19, 0], True),           # nlon * 4 bytes IS 16-byte aligned in fp32
    ("dhconv", [2, 64, 20, 21, 0, 0], True),
    ("dhconv", [3, 32, 20, 21, 1, 0], True),
    ("dhconv", [2, 256, 180, 181, 1, 0], True),
    ("ileg", [1, 8, 16, 16, 9, 1], True),           # MN-major A operand in 32-element atoms, X layout
    ("ileg", [2, 64, 180, 180, 181, 2], True),
    ("ileg", [2, 64, 180, 180, 181, 3], True),
    ("idft", [1, 8, 16, 32, 17, 0], True),
    ("idft", [2, 16, 180, 360, 181, 7], True),      # + bias + fp32 addend staged by TMA (two 32-column passes) -> GELU
    ("idft", [2, 64, 180, 360, 181, 0], True),
    ("convb", [2, 36, 256, 64800, 0, 3], True),     # MN-major B operand (pixels), K = 36 tail
    ("convb", [2, 256, 256, 64800, 0, 8], True),    # + pos-embed
    ("convb", [2, 256, 512, 64800, 1, 3], True),    # fc1: folded weights, bias, GELU (exact erf in this mode)
    ("convb", [2, 512, 256, 64800, 0, 5], True),    # fc2: bias + affine fp32 residual staged by TMA
    ("convb", [2, 512, 256, 64800, 0, 21], True),   # ... with dropout
    ("convb", [2, 256, 256, 64800, 1, 15], True),
    ("convb", [2, 256, 34, 64800, 0, 0], True),     # decoder1: 34 rows
    ("convb", [1, 16, 16, 296, 0, 0], True),        # single tile, N tail
    ("convb", [1, 16, 16, 300, 0, 0], False),       # hw not a multiple of 8 -> CUDA-core engine (same rule as bf16)
]


_BATCH = None


def _batch_results():
    """All cases of this file in ONE child process (one CUDA context instead of ~110): results keyed by (op code, dims,
    tc_debug).  A case that kills the child leaves the later keys missing; those cases are then re-run one by one."""
    global _BATCH
    if _BATCH is None:
        cases = [[OPS[c[0]], c[1] + [0] * (6 - len(c[1])), 0] for c in CASES]
        cases += [[OPS[c[0]] + 100, c[1] + [0] * (6 - len(c[1])), 0] for c in TF32_CASES]
        cases += [[OPS[c[0]], c[1] + [0] * (6 - len(c[1])), c[2]] for c in VARIANTS]
        env = dict(os.environ)
        env.pop("SFNO_TC_DEBUG", None)
        _BATCH = {}
        try:
            p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "tc_selftest_cli.py"), "--batch"], input=json.dumps(cases),
                               capture_output=True, text=True, timeout=900, env=env)
            for line in p.stdout.splitlines():
                if line.startswith("{"):
                    r = json.loads(line)
                    _BATCH[(r["op"], tuple(r["dims"]), r["tc_debug"])] = r
        except Exception:
            pass
    return _BATCH


def _run_case(op, dims, tc_debug=0, tf32=False):
    code = OPS[op] + (100 if tf32 else 0)
    hit = _batch_results().get((code, tuple(dims + [0] * (6 - len(dims))), tc_debug))
    if hit is not None:
        return hit
    cmd = [sys.executable, os.path.join(ROOT, "tests", "tc_selftest_cli.py"), str(code)] + [str(d) for d in dims]
    env = dict(os.environ)
    env.pop("SFNO_TC_DEBUG", None)
    if tc_debug:
        env["SFNO_TC_DEBUG"] = str(tc_debug)
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    line = p.stdout.strip().splitlines()[-1] if p.stdout.strip() else ""
    assert line, f"no output; stderr: {p.stderr[-2000:]}"
    return json.loads(line)


# engine variants selected by the correct-result tc_debug bits: 256 = single 128-row tiles where the op would use
# dual-M tiles, 128 = role-wait counters on
VARIANTS = [
    ("leg", [8, 256, 180, 180, 181, 1], 256),
    ("leg", [2, 64, 180, 180, 5, 1], 256),
    ("dhconv", [5, 32, 20, 21, 1, 0], 128),       # batch 5: eight-row boxes straddle wavenumbers
    ("convb", [2, 256, 256, 64800, 1, 15], 128),
    ("idft", [2, 16, 180, 360, 181, 7], 128),
]


@pytest.mark.parametrize("op,dims,tc_debug", VARIANTS, ids=[f"{c[0]}-{'x'.join(map(str, c[1]))}-dbg{c[2]}" for c in VARIANTS])
def test_tc_engine_variants_match_cuda_core_engine(op, dims, tc_debug):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    r = _run_case(op, dims, tc_debug)
    print(json.dumps(r))
    assert r["status"] == 0, r["error"]
    assert r["tc_used"] and r["nonfinite"] == 0 and r["max_ref"] > 0
    tol = 2e-5 if op == "conv" else 1.2e-2
    assert r["max_err"] <= tol * r["max_ref"], r
    if tc_debug & 128:
        c = r["counters"]
        assert c["ctas"] > 0 and c["epi_busy"] > 0.0 and c["cta_cycles"] > 0


@pytest.mark.parametrize("op,dims,expect_tc", CASES, ids=[f"{c[0]}-{'x'.join(map(str, c[1]))}" for c in CASES])
def test_tc_engine_matches_cuda_core_engine(op, dims, expect_tc):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    r = _run_case(op, dims)
    print(json.dumps(r))
    assert r["status"] == 0, r["error"]
    assert bool(r["tc_used"]) == expect_tc
    assert r["nonfinite"] == 0
    assert r["max_ref"] > 0
    # both engines accumulate in fp32; outputs are bf16 (1 ulp = 2^-8 relative) or fp32 (conv case)
    tol = 2e-5 if op == "conv" else 1.2e-2

    assert r["max_err"] <= tol * r["max_ref"], r


@pytest.mark.parametrize("op,dims,expect_tc", TF32_CASES, ids=[f"tf32-{c[0]}-{'x'.join(map(str, c[1]))}" for c in TF32_CASES])
def test_tf32_engine_matches_cuda_core_engine(op, dims, expect_tc):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    r = _run_case(op, dims, tf32=True)
    print(json.dumps(r))
    assert r["status"] == 0, r["error"]
    assert bool(r["tc_used"]) == expect_tc
    assert r["nonfinite"] == 0
    assert r["max_ref"] > 0
    # identical products (TF32-exact operands), fp32 accumulation in both engines, fp32 outputs (not rounded here)
    assert r["max_err"] <= 3e-5 * r["max_ref"], r
