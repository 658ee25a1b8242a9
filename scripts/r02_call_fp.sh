#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "resync or fingerprint or param" 2>&1 | tail -2
timeout 300 python - <<'PY'
import sys; sys.path.insert(0, ".")
import torch
from spherical_dyffusion_b200 import configs
from spherical_dyffusion_b200.profile import profile_forward
dev = torch.device("cuda:0")
m = configs.build(configs.ACE_FORECASTER, precision="bf16").to(dev)      # default param_check = "checksum"
x = torch.randn(8, 34, 180, 360, device=dev); c = torch.randn(8, 2, 180, 360, device=dev); t = torch.full((8,), 2.0, device=dev)
with torch.inference_mode():
    recs = profile_forward(m, x, t, c, repeats=3)
print("param_fingerprint ms per forward:", round(recs["param_fingerprint"]["ms_total"], 4), "total", round(sum(r["ms_total"] for r in recs.values()), 3))
PY
