#!/usr/bin/env python
"""A/B of the forked-stream dropout-mask generation: ACE interpolator forward (batch 8, inference dropout on) and one
sampling window (16 forwards, batch 8, one CUDA graph), option mask_overlap = 0 / 1.  Device time, CUDA events."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench_extra import _ace_pair  # noqa: E402
from spherical_dyffusion_b200 import _lib  # noqa: E402
from spherical_dyffusion_b200.dyffusion import DYffusion  # noqa: E402

dev = torch.device("cuda:0")
fore, ipol = _ace_pair(dev, "bf16")
B = 8
g = torch.Generator().manual_seed(0)
x = torch.randn(B, 68, 180, 360, generator=g).to(dev)
c = torch.randn(B, 2, 180, 360, generator=g).to(dev)
t = torch.full((B,), 2.0, device=dev)
ipol.enable_inference_dropout()


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def set_overlap(v):
    for m in (fore, ipol):
        if m.__dict__.get("_net") is not None:
            _lib.check(_lib.lib().sfno_net_set_option(m._net, b"mask_overlap", v), "set_option")


out = {}
with torch.inference_mode():
    ipol(x, time=t, condition=c)          # creates the native net
    fore(x[:, :34].contiguous(), time=t, condition=c)
    for v in (0, 1, 0, 1):
        set_overlap(v)
        out.setdefault(f"interpolator_forward_ms_overlap{v}", []).append(round(timed(lambda: ipol(x, time=t, condition=c)), 3))
    x0 = x[:, :34].contiguous()
    for graph in (False, True):
        for v in (0, 1):
            set_overlap(v)
            dy = DYffusion(fore, ipol, timesteps=6, forward_conditioning="none", time_encoding="dynamics", capture_graph=graph)
            out[f"window_ms_graph{int(graph)}_overlap{v}"] = round(timed(lambda: dy.sample(x0, static_condition=c), n=5), 2)
            if graph:
                dy.release_graphs()
print(json.dumps(out))
