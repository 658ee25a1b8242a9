#!/usr/bin/env python
"""Per-kernel device times of one ACE interpolator forward (70 -> 34 channels, inference dropout on), batch 8."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench_extra import _ace_pair  # noqa: E402
from spherical_dyffusion_b200.profile import profile_forward  # noqa: E402

dev = torch.device("cuda:0")
fore, ipol = _ace_pair(dev, sys.argv[1] if len(sys.argv) > 1 else "bf16")
B = 8
g = torch.Generator().manual_seed(0)
x = torch.randn(B, 68, 180, 360, generator=g).to(dev)
c = torch.randn(B, 2, 180, 360, generator=g).to(dev)
t = torch.full((B,), 2.0, device=dev)
for name, m, xin in (("interpolator", ipol, x), ("forecaster", fore, x[:, :34].contiguous())):
    if name == "interpolator":
        m.enable_inference_dropout()
    recs = profile_forward(m, xin, t, c, repeats=3)
    tot = sum(r["ms_total"] for r in recs.values())
    print(json.dumps({"model": name, "ms_total": round(tot, 3),
                      "per_kernel_ms": {k: round(v["ms_total"], 3) for k, v in sorted(recs.items(), key=lambda kv: -kv[1]["ms_total"])[:14]}}))
