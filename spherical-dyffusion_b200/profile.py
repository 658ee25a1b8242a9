"""Per-kernel device timing of one forward (CUDA events recorded by the library after every launch, on the launch
stream) and the roofline arithmetic used by ``bench.py``.  Algorithmic flops / bytes per launch follow SURVEY.md
section 8d (dense counts: 1 MAC = 2 flop, 1 complex MAC = 8 flop; minimum HBM traffic = operands read once + result)."""
from __future__ import annotations

import ctypes
from collections import OrderedDict

import torch

from . import _lib
from ._util import stream_ptr


def profile_forward(model, x, time, condition, repeats: int = 3):
    """Returns an OrderedDict name -> dict(ms_total, launches) averaged over `repeats` forwards."""
    L = _lib.lib()
    dev = x.device
    acc = OrderedDict()
    with torch.inference_mode(), torch.cuda.device(dev):
        model(x, time=time, condition=condition)
        torch.cuda.synchronize(dev)
        for _ in range(repeats):
            _lib.check(L.sfno_b200_profile_begin(stream_ptr(dev)), "profile_begin")
            try:
                model(x, time=time, condition=condition)
            finally:
                cap = 4096
                names = ctypes.create_string_buffer(1 << 16)
                ms = (ctypes.c_float * cap)()
                n = _lib.check(L.sfno_b200_profile_end(names, len(names), ms, cap), "profile_end")
            for nm, t in zip(names.value.decode().split("\n"), list(ms)[:n]):
                r = acc.setdefault(nm, dict(ms_total=0.0, launches=0))
                r["ms_total"] += t / repeats
                r["launches"] += 1
    for r in acc.values():
        r["launches"] //= repeats
    return acc


def algorithmic_work(model, batch: int):
    """name -> (flops per launch, min bytes per launch, kind) for the GEMM-shaped and streaming kernels."""
    e = 2 if model.precision == "bf16" else 4
    B, C, Cin, Cout = batch, model.embed_dim, model.in_chans, model.out_chans
    H, W = model.img_shape
    P, Lm, Mm = H * W, model.modes_lat, model.modes_lon
    hid = int(C * model.mlp_ratio)
    A = B * C * P * e                 # one activation tensor
    S = B * C * Lm * Mm * 2 * e       # one spectral tensor
    Fb = B * C * H * Mm * 2 * e       # longitude-spectral tensor
    ccat = C + (Cin if model.big_skip else 0)
    w = {}
    w["dft_fwd"] = (2.0 * B * C * H * W * 2 * Mm, A + Fb)
    w["legendre_fwd"] = (4.0 * B * C * Mm * Lm * H, Fb + S + Mm * Lm * H * e)
    w["legendre_inv"] = (4.0 * B * C * Mm * Lm * H, Fb + S + Mm * Lm * H * e)
    w["dft_inv"] = (2.0 * B * C * H * W * 2 * Mm, Fb + 2 * A)
    w["dhconv"] = (8.0 * B * C * C * Lm * Mm, 2 * S + 4 * C * C * Lm * e)
    w["inner_skip"] = (2.0 * B * P * C * C, 2 * A)
    w["mlp_fc1"] = (2.0 * B * P * C * hid, A + B * hid * P * e)
    w["mlp_fc2"] = (2.0 * B * P * C * hid, B * hid * P * e + 2 * A)
    w["encoder0"] = (2.0 * B * P * Cin * C, B * Cin * P * e + A)
    w["encoder1"] = (2.0 * B * P * C * C, 3 * A)
    w["decoder0"] = (2.0 * B * P * ccat * C, B * ccat * P * e + A)
    w["decoder1"] = (2.0 * B * P * C * Cout, A + B * Cout * P * 4)
    w["instance_stats0"] = (0.0, A)
    w["instance_stats1"] = (0.0, A)
    return w


def live_fraction(model) -> float:
    """Share of the (l, m) spectral grid the triangular kernels touch: degree l is live for wavenumber m iff
    l >= live_l0(m) = min(m & ~63, (lmax - 1) & ~63) (csrc/ops.cuh).  1.0 for the diagonal operator (dense transforms)."""
    if model.operator_type != "dhconv":
        return 1.0
    L, M = model.modes_lat, model.modes_lon
    last = ((L - 1) & ~63) if L > 0 else 0
    live = sum(L - min(m & ~63, last) for m in range(M))
    return live / float(L * M)


def executed_work(model, batch: int):
    """name -> (flops, bytes) the triangular kernels actually execute / move (only the live part of the spectral tensors
    and Legendre tables); other kernels are dense and not listed."""
    e = 2 if model.precision == "bf16" else 4
    B, C = batch, model.embed_dim
    H, W = model.img_shape
    Lm, Mm = model.modes_lat, model.modes_lon
    f = live_fraction(model)
    S = B * C * Lm * Mm * 2 * e
    Fb = B * C * H * Mm * 2 * e
    tab = Mm * Lm * H * e
    return {
        "legendre_fwd": (f * 4.0 * B * C * Mm * Lm * H, Fb + f * (S + tab)),
        "legendre_inv": (f * 4.0 * B * C * Mm * Lm * H, Fb + f * (S + tab)),
        "dhconv": (f * 8.0 * B * C * C * Lm * Mm, 2 * f * S + 4 * C * C * Lm * e),
    }


def _ncu_traffic(kernel_name):
    """DRAM bytes per launch of `kernel_name` from the committed ncu capture (profiles/ncu_traffic.json), else None."""
    import json
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")
    try:
        with open(path) as f:
            return float(json.load(f)["kernels"][kernel_name]["traffic_bytes"])
    except (OSError, KeyError, ValueError):
        return None


def roofline_from_profile(recs, pk, model=None, batch=None):
    recs = OrderedDict(recs)
    host_gap = recs.pop("host_before_first_launch", None)   # host time between profile_begin and the first launch
    total = sum(r["ms_total"] for r in recs.values())
    top_name, top = max(recs.items(), key=lambda kv: kv[1]["ms_total"])
    out = {"kernel": top_name, "share_of_step": top["ms_total"] / total if total else None,
           "launches_per_step": top["launches"], "ms_per_launch": top["ms_total"] / max(top["launches"], 1),
           "step_ms_profiled": total,
           "per_kernel_ms": {k: round(v["ms_total"], 4) for k, v in sorted(recs.items(), key=lambda kv: -kv[1]["ms_total"])}}
    if host_gap is not None:
        out["host_before_first_launch_ms"] = round(host_gap["ms_total"], 4)
    if model is not None and batch is not None:
        work = algorithmic_work(model, batch).get(top_name)
        if work is not None:
            flops, nbytes = work
            sec = out["ms_per_launch"] * 1e-3
            ridge = pk["bf16_tflops"] * 1e12 / (pk["hbm_gbs"] * 1e9)
            tensor_bound = model.precision == "bf16" and nbytes > 0 and flops / nbytes > ridge
            if tensor_bound:
                out.update(bound="tensor", achieved=flops / sec / 1e12, peak=pk["bf16_tflops_sustained"], unit="TFLOP/s")
            else:
                out.update(bound="hbm", achieved=nbytes / sec / 1e9, peak=pk["hbm_gbs"], unit="GB/s")
            out["frac"] = out["achieved"] / out["peak"]
            out["peak_source"] = pk["source"]
            out["algorithmic_flops_per_launch"] = flops
            out["algorithmic_bytes_per_launch"] = nbytes
            out["tflops_achieved"] = flops / sec / 1e12
            out["traffic"] = _ncu_traffic(top_name)
            out["traffic_source"] = "profiles/ncu_traffic.json (ncu --set full, dram read + write bytes of one launch)" if out["traffic"] else None
        # every GEMM-shaped kernel against its own bound (all of them sit below the ridge at ACE size -> HBM)
        table = {}
        executed = executed_work(model, batch)
        for name, (flops, nbytes) in algorithmic_work(model, batch).items():
            if name in recs and recs[name]["launches"] > 0 and nbytes > 0:
                sec = recs[name]["ms_total"] / recs[name]["launches"] * 1e-3
                table[name] = {"ms_per_launch": round(sec * 1e3, 4), "GBps": round(nbytes / sec / 1e9, 1),
                               "frac_hbm": round(nbytes / sec / 1e9 / pk["hbm_gbs"], 3),
                               "TFLOPs": round(flops / sec / 1e12, 1),
                               "frac_tensor": round(flops / sec / 1e12 / pk["bf16_tflops_sustained"], 3)}
                if name in executed:   # triangular kernels: what they execute / move, next to the dense accounting above
                    xf, xb = executed[name]
                    table[name].update({"GBps_executed": round(xb / sec / 1e9, 1), "frac_hbm_executed": round(xb / sec / 1e9 / pk["hbm_gbs"], 3),
                                        "TFLOPs_executed": round(xf / sec / 1e12, 1),
                                        "frac_tensor_executed": round(xf / sec / 1e12 / pk["bf16_tflops_sustained"], 3)})
        out["per_kernel_roofline"] = table
        out["live_fraction_of_spectral_grid"] = round(live_fraction(model), 4)
        # BASELINE.json's "SHT tensor-pipe % of peak": the transform pair (longitude DFT + Legendre, both directions) and
        # the spectral contraction as one aggregate -- dense algorithmic flops of all their launches / their summed time
        work = algorithmic_work(model, batch)
        for key, names in (("sht", ("dft_fwd", "legendre_fwd", "legendre_inv", "dft_inv")),
                           ("sht_and_spectral_conv", ("dft_fwd", "legendre_fwd", "dhconv", "legendre_inv", "dft_inv"))):
            live = [n for n in names if n in recs and recs[n]["ms_total"] > 0]
            if len(live) == len(names):
                flops = sum(work[n][0] * recs[n]["launches"] for n in live)
                xflops = sum(executed.get(n, work[n])[0] * recs[n]["launches"] for n in live)
                sec = sum(recs[n]["ms_total"] for n in live) * 1e-3
                out.setdefault("tensor_pipe", {})[key] = {
                    "TFLOPs_dense": round(flops / sec / 1e12, 1), "TFLOPs_executed": round(xflops / sec / 1e12, 1),
                    "ms_per_forward": round(sec * 1e3, 4),
                    "frac_of_bf16_sustained": round(flops / sec / 1e12 / pk["bf16_tflops_sustained"], 3),
                    "frac_of_bf16_sustained_executed": round(xflops / sec / 1e12 / pk["bf16_tflops_sustained"], 3),
                    "frac_of_bf16_burst": round(flops / sec / 1e12 / pk["bf16_tflops"], 3)}
        out["per_kernel_roofline_note"] = ("frac_hbm / frac_tensor use the DENSE algorithmic bytes / flops per launch (SURVEY 8d); the Legendre "
                                           "and dhconv kernels execute only the live part of the spectral grid (live_fraction_of_spectral_grid): "
                                           "their *_executed entries count what is actually computed / moved")
    return out
