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
    # per-launch profile of one step (library launches; torch glue between two launches is booked on the later one)
    import ctypes
    from spherical_dyffusion_b200 import _lib
    from spherical_dyffusion_b200._util import stream_ptr
    L = _lib.lib()
    _lib.check(L.sfno_b200_profile_begin(stream_ptr(dev)), "profile_begin")
    step()
    names = ctypes.create_string_buffer(1 << 18)
    cap = 16384
    arr = (ctypes.c_float * cap)()
    nrec = _lib.check(L.sfno_b200_profile_end(names, len(names), arr, cap), "profile_end")
    acc = {}
    for nm, t in zip(names.value.decode().split("\n"), list(arr)[:nrec]):
        a = acc.setdefault(nm, [0.0, 0])
        a[0] += t
        a[1] += 1
    prof = {k: [round(v[0], 2), v[1]] for k, v in sorted(acc.items(), key=lambda kv: -kv[1][0])[:16]}
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
                      "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2**30, 2), "profile_ms_launches": prof}))
    del m
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats()
