"""Time one training step (forward + loss + backward through the library's adjoint kernels) of the ACE-sized forecaster
per engine precision, next to the fused inference forward.  Device timing with CUDA events; prints one JSON line each."""
import json
import sys

sys.path.insert(0, ".")
import torch

import spherical_dyffusion_b200 as sb
from spherical_dyffusion_b200 import configs

dev = torch.device("cuda:0")
for precision, B in (("bf16", 4), ("tf32", 4), ("fp32", 2)):
    m = configs.build(dict(configs.ACE_FORECASTER, loss_function="l1"), precision=precision).to(dev).train()
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, 34, 180, 360, generator=g).to(dev)
    cond = torch.randn(B, 2, 180, 360, generator=g).to(dev)
    y = torch.randn(B, 34, 180, 360, generator=g).to(dev)
    t = torch.full((B,), 1.0, device=dev)
    kw = {"time": t} if m.with_time_emb else {}

    def step():
        for p in m.parameters():
            p.grad = None
        m.get_loss(x, y, condition=cond, **kw)["loss"].backward()

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 5
    e0.record()
    for _ in range(n):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    m.eval()
    with torch.no_grad():
        for _ in range(2):
            m(x, condition=cond, **kw)
        e0.record()
        for _ in range(n):
            m(x, condition=cond, **kw)
        e1.record()
        torch.cuda.synchronize()
    print(json.dumps({"precision": precision, "batch": B, "train_step_ms": round(ms, 2), "samples_per_s": round(B / ms * 1e3, 1),
                      "fused_inference_forward_ms": round(e0.elapsed_time(e1) / n, 2),
                      "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2**30, 2)}))
    del m
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats()
