#!/usr/bin/env python
"""Additional workloads of BASELINE.json `configs` (bench.py keeps the one-line contract for the headline metric):

    python bench_extra.py window   [--batch 8]             # config 3: one DYffusion sampling window (6 forecaster + 10 interpolator forwards)
    python bench_extra.py rollout  [--members 25 --steps 24]  # config 4: ensemble rollout with per-step statistics (torchrun for N GPUs)
    python bench_extra.py scaled   [--batch 1]             # config 5: embed 512, 12 blocks, 720x1440 forward
    python bench_extra.py sht                              # config 2: RealSHT -> InverseRealSHT round-trip sweep, both grids
    python bench_extra.py graph                            # small-batch latency: eager launches vs CUDA-graph replay

Each prints one JSON line per measurement.  Synthetic data, random-init weights, CUDA-event timing after warm-up.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

STEPS_PER_YEAR = 1460.0


def _model(config, dev, precision, seed=0, min_max_time=(0.0, 5.0)):
    """Random-init module of a spherical_dyffusion_b200.configs entry on `dev` (eval mode)."""
    from spherical_dyffusion_b200 import configs

    return configs.build(config, precision=precision, seed=seed, min_max_time=min_max_time, check_time_range=False,
                         param_check="version").to(dev).eval()


def _timeit(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def _ace_pair(dev, precision, seed=0):
    from spherical_dyffusion_b200 import configs

    fore = _model(configs.ACE_FORECASTER, dev, precision, seed=seed, min_max_time=(0.0, 5.0))
    ipol = _model(configs.ACE_INTERPOLATOR, dev, precision, seed=seed + 1, min_max_time=(1.0, 5.0))
    return fore, ipol


def run_window(args):
    from spherical_dyffusion_b200.dyffusion import DYffusion

    dev = torch.device("cuda:0")
    fore, ipol = _ace_pair(dev, args.precision)
    dy = DYffusion(fore, ipol, timesteps=6, forward_conditioning="none", time_encoding="dynamics")
    B = args.batch
    g = torch.Generator().manual_seed(0)
    x0 = torch.randn(B, 34, 180, 360, generator=g).to(dev)
    forcing = torch.randn(B, 2, 180, 360, generator=g).to(dev)
    ms = _timeit(lambda: dy.sample(x0, static_condition=forcing), args.steps, args.warmup)
    steps_per_s = 6.0 / (ms * 1e-3)  # per member
    print(json.dumps({
        "workload": "DYffusion sampling window, horizon 6, cold sampling: 6 forecaster (36->34) + 10 interpolator (70->34, dropout on) "
                    f"SFNO forwards, batch {B}, {args.precision}",
        "ms_per_window": ms, "sfno_forwards_per_s": 16.0 * B / (ms * 1e-3), "steps_per_s_per_member": steps_per_s,
        "member_sypd": steps_per_s * 86400.0 / STEPS_PER_YEAR, "batch_sypd": B * steps_per_s * 86400.0 / STEPS_PER_YEAR,
        "forwards": dy.forwards_per_window(), "data": "synthetic"}), flush=True)


def run_rollout(args):
    import torch.distributed as dist

    from spherical_dyffusion_b200.dyffusion import DYffusion
    from spherical_dyffusion_b200.ensemble import EnsembleStatistics, area_weights
    from spherical_dyffusion_b200.rollout import EnsembleRollout

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    fore, ipol = _ace_pair(dev, args.precision)
    dy = DYffusion(fore, ipol, timesteps=6, forward_conditioning="none", time_encoding="dynamics")
    stats = EnsembleStatistics(args.members)
    weights = area_weights(torch.linspace(-89.5, 89.5, 180), 360).to(dev)
    g = torch.Generator().manual_seed(0)
    ic = torch.randn(34, 180, 360, generator=g).to(dev)
    base_forcing = torch.randn(2, 180, 360, generator=g).to(dev)
    truth = torch.randn(34, 180, 360, generator=g).to(dev)

    def forcing_fn(step, n, d):
        return (base_forcing * (1.0 + 0.01 * step)).unsqueeze(0).expand(n, -1, -1, -1).contiguous()

    ro = EnsembleRollout(dy, stats, forcing_fn, truth_fn=lambda s, d: truth, weights=weights)
    ro.run(ic, n_steps=6)  # warm-up window
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    hist = ro.run(ic, n_steps=args.steps)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if rank == 0:
        steps_per_s = args.steps / (ms * 1e-3)
        print(json.dumps({
            "workload": f"{args.members}-member ensemble rollout, {args.steps} x 6-h steps, members sharded over {world} GPU(s) "
                        f"(max {len(stats.local_ids)} local), per-step mean/spread/rmse/ssr/fair-CRPS over NCCL, {args.precision}",
            "n_gpus": world, "ms_total": ms, "ensemble_steps_per_s": steps_per_s,
            "ensemble_sypd": steps_per_s * 86400.0 / STEPS_PER_YEAR, "member_steps_per_s_aggregate": steps_per_s * args.members,
            "last_crps_mean": float(hist["crps"][-1].mean()), "last_spread_mean": float(hist["spread"][-1].mean()),
            "data": "synthetic"}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_scaled(args):
    from spherical_dyffusion_b200 import configs

    dev = torch.device("cuda:0")
    # random weights generated on the GPU to avoid a 20 GB host state dict
    import spherical_dyffusion_b200 as sb

    with torch.device(dev):
        m = sb.SphericalFourierNeuralOperatorNet(**configs.SCALED_FORECASTER, precision=args.precision, check_time_range=False)
    m.set_min_max_time(0, 5)
    m = m.to(dev).eval()
    B = args.batch
    x = torch.randn(B, 34, 720, 1440, device=dev)
    c = torch.randn(B, 2, 720, 1440, device=dev)
    t = torch.full((B,), 3.0, device=dev)
    with torch.inference_mode():
        ms = _timeit(lambda: m(x, time=t, condition=c), args.steps, args.warmup)
        y = m(x, time=t, condition=c)
    flop = 66.78e12 * B
    from spherical_dyffusion_b200.profile import profile_forward
    recs = profile_forward(m, x, t, c, repeats=1)
    per_kernel = {k: [round(v["ms_total"], 3), v["launches"]] for k, v in sorted(recs.items(), key=lambda kv: -kv[1]["ms_total"])[:10]}
    print(json.dumps({
        "workload": f"scaled SFNO forward: embed 512, 12 blocks, 720x1440, lmax 720, batch {B}, {args.precision} (random-init, synthetic)",
        "params": m.num_params, "ms_per_forward": ms, "samples_per_s": B / (ms * 1e-3), "model_tflops": flop / (ms * 1e-3) / 1e12,
        "finite": bool(torch.isfinite(y).all()), "mem_GB": torch.cuda.max_memory_allocated() / 2**30,
        "per_kernel_ms_and_launches": per_kernel}), flush=True)


def run_graph(args):
    """Latency of one ACE forecaster forward at small batch: eager launches vs replay of a captured CUDA graph."""
    from spherical_dyffusion_b200 import configs

    dev = torch.device("cuda:0")
    m = _model(configs.ACE_FORECASTER, dev, args.precision)
    for B in (1, 2, 8):
        x = torch.randn(B, 34, 180, 360, device=dev)
        c = torch.randn(B, 2, 180, 360, device=dev)
        t = torch.full((B,), 3.0, device=dev)
        with torch.inference_mode():
            eager_ms = _timeit(lambda: m(x, time=t, condition=c), 20, 3)
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                m(x, time=t, condition=c)
            torch.cuda.current_stream(dev).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                y = m(x, time=t, condition=c)
            graph_ms = _timeit(graph.replay, 20, 3)
        print(json.dumps({"workload": f"ACE forecaster forward, batch {B}, {args.precision}: eager launches vs CUDA-graph replay",
                          "batch": B, "eager_ms": eager_ms, "graph_ms": graph_ms, "finite": bool(torch.isfinite(y).all())}), flush=True)


def run_sht(args):
    import spherical_dyffusion_b200 as sb

    dev = torch.device("cuda:0")
    for grid in ("legendre-gauss", "equiangular"):
        for prec in ("bf16", "fp32"):
            sht = sb.RealSHT(180, 360, lmax=180, mmax=181, grid=grid, precision=prec)
            isht = sb.InverseRealSHT(180, 360, lmax=180, mmax=181, grid=grid, precision=prec)
            for bc in (1, 16, 256, 1024, 4096):
                if prec == "fp32" and bc > 1024:
                    continue
                x = torch.randn(1, bc, 180, 360, device=dev)
                ms = _timeit(lambda: isht(sht(x)), 5, 2)
                print(json.dumps({"workload": "RealSHT -> InverseRealSHT round trip 180x360 (reference layout in/out)", "grid": grid,
                                  "precision": prec, "fields": bc, "ms": ms, "fields_per_s": bc / (ms * 1e-3)}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workload", choices=["window", "rollout", "scaled", "sht", "graph"])
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--members", type=int, default=25)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    args = ap.parse_args()
    if args.workload == "window":
        args.batch = args.batch or 8
        args.steps = args.steps or 5
        run_window(args)
    elif args.workload == "rollout":
        args.steps = args.steps or 24
        run_rollout(args)
    elif args.workload == "scaled":
        args.batch = args.batch or 1
        args.steps = args.steps or 3
        run_scaled(args)
    elif args.workload == "graph":
        run_graph(args)
    else:
        run_sht(args)


if __name__ == "__main__":
    main()
