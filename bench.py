#!/usr/bin/env python
"""bench.py -- SFNO forward throughput on B200 (metric of BASELINE.json: "SFNO fwd samples/s").

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU)
    python bench.py --impl reference --steps K --warmup W     # the reference's CPU implementation (oracle port)

A "step" is one ACE-sized SFNO forward (forecaster: 34+2 -> 34 channels, embed 256, 8 blocks, dhconv, 180x360)
on a batch of `--batch` synthetic samples per GPU with random-init weights.  `value` is whole-job samples/s with
inputs resident in HBM; `e2e` is the same through the public module call with pinned host inputs and a host read
of the output inside the timed region.  One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "SFNO fwd samples/s"
UNIT = "samples/s"
FORWARDS_PER_STEP = 16.0 / 6.0   # DYffusion window: 6 forecaster + 10 interpolator forwards per 6 output steps
STEPS_PER_YEAR = 1460.0
FLOP_PER_SAMPLE = 605.0e9        # dense algorithmic flops of one forecaster forward


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        p = json.load(open(path))
        return dict(hbm_gbs=p["hbm_gbs"], bf16_tflops=p["bf16_tflops"], bf16_tflops_sustained=p.get("bf16_tflops_sustained"),
                    source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


def workload_config(args):
    return {"workload": f"ACE SFNO forecaster forward 36->34ch embed256 x8 blocks dhconv 180x360, batch {args.batch}/GPU",
            "batch_per_gpu": args.batch, "precision": args.precision,
            "check_time_range": False,   # the drop-in default (True) adds the reference's per-forward host sync (sfnonet.py:777-782)
            "param_check": "version",    # the drop-in default ("checksum") fingerprints the parameters every forward (+0.15 ms, one sync)
            "baseline_config": "BASELINE.json configs[0] (ACE-sized SFNO, 180x360) at configs[2]'s batch 8 / bf16; configs[1,3,4] = "
                               "bench_extra.py sht / rollout / scaled",
            "l2": "per-step working set (>= 2.5 GB of activations) exceeds the 126 MB L2; no explicit flush"}


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.gpu, self.mark_idx = [], None, gpu_index, 0

    def mark(self):
        """The timed region starts here: only later samples count (earlier ones were taken under the warm-up load)."""
        self.mark_idx = len(self.rows)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        rows = self.rows[self.mark_idx:]
        during = bool(rows)
        if not rows:  # timed region shorter than the sampling period: fall back to the last sample under warm-up load
            rows = self.rows[-1:]
        sm = sorted(float(r[1]) for r in rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        smax = [float(r[2]) for r in rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            if len(r) >= 9:
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm), "sampled": "timed region" if during else "warm-up (timed region < 100 ms)"}


def bind_to_gpu_numa_node(device_index: int):
    """Run this process (and therefore first-touch its pinned staging buffers) on the CPUs of the NUMA node the GPU hangs
    off: host<->device copies from the far socket run at a fraction of the link rate.  Returns (node, previous affinity)
    or (None, None) when the topology cannot be read."""
    try:
        import torch

        pr = torch.cuda.get_device_properties(device_index)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
        if node < 0:
            return None, None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        prev = os.sched_getaffinity(0)
        allowed = cpus & prev
        if not allowed:
            return None, None
        os.sched_setaffinity(0, allowed)
        return node, prev
    except Exception:
        return None, None


def build_case(batch, seed=0):
    """Configuration, weights and inputs of the CPU legs (reference arm / cpu_baseline) -- the only users of oracle/."""
    from oracle.sfno_oracle import ACE_FORECASTER, SFNOConfig, random_state_dict

    cfg = SFNOConfig(**ACE_FORECASTER)
    sd = random_state_dict(cfg, seed=seed)
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(batch, 34, 180, 360, generator=g)
    c = torch.randn(batch, 2, 180, 360, generator=g)
    t = torch.full((batch,), 3.0)
    return cfg, sd, x, c, t


def cpu_forward_time(cfg, sd, x, c, t, steps, warmup):
    from oracle.sfno_oracle import SFNOOracle

    orc = SFNOOracle(cfg, sd)
    for _ in range(warmup):
        orc(x, time=t, condition=c)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        orc(x, time=t, condition=c)
        times.append(time.perf_counter() - t0)
    return times


def run_reference(args):
    """Reference arm: the reference's own PyTorch-CPU algorithm (oracle port; the reference needs packages that
    are not installable here) on the host cores, one sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg, sd, x, c, t = build_case(1)
    times = cpu_forward_time(cfg, sd, x, c, t, args.steps, args.warmup)
    total = sum(times)
    value = args.steps * 1.0 / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "ACE SFNO forecaster forward 36->34ch embed256 x8 blocks dhconv 180x360, batch 1 on host CPU"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{args.steps} forwards of 1 sample (oracle/sfno_oracle.py, fp32, torch CPU)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def build_b200_case(batch, precision, seed, dev):
    """The B200 arm's workload: random-init ACE forecaster (the reference's initialisers, spherical_dyffusion_b200.configs)
    on `dev`, synthetic N(0,1) host inputs [batch,34,180,360] + forcings [batch,2,180,360], time 3.  Nothing of oracle/."""
    from spherical_dyffusion_b200 import configs

    model = configs.build(configs.ACE_FORECASTER, precision=precision, seed=seed, min_max_time=(0, 5), check_time_range=False,
                          param_check="version")
    model = model.to(dev).eval()
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(batch, 34, 180, 360, generator=g)
    c = torch.randn(batch, 2, 180, 360, generator=g)
    t = torch.full((batch,), 3.0)
    return model, x, c, t


def run_b200(args):
    import torch.distributed as dist

    import spherical_dyffusion_b200 as sb
    from spherical_dyffusion_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_node, prev_affinity = bind_to_gpu_numa_node(local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)

    B = args.batch
    model, x, c, t = build_b200_case(B, args.precision, rank, dev)
    xd, cd, td = x.to(dev), c.to(dev), t.to(dev)
    x_pin, c_pin = x.pin_memory(), c.pin_memory()
    y_pin = torch.empty(B, 34, 180, 360).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        v = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
        return float(v.item())

    with torch.inference_mode():
        # ---- device-resident throughput ---------------------------------------------------------------------
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        for _ in range(max(args.warmup, 3)):
            model(xd, time=td, condition=cd)
        barrier()
        n0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        sampler.mark()
        e0.record()
        for _ in range(args.steps):
            model(xd, time=td, condition=cd)
        e1.record()
        barrier()
        ms_dev = max_over_ranks(e0.elapsed_time(e1))
        launches = _lib.launch_count() - n0
        clocks = sampler.stop() if rank == 0 else None

        # ---- end to end through the public call: pinned host -> device -> forward -> host, every step --------------
        # The copies run on their own streams and are double-buffered, so step i+1's upload and step i-1's download
        # overlap step i's forward (what a serving loop does); every step still moves all its bytes inside the timed
        # region, and the region ends only when the last result has landed in host memory.
        compute = torch.cuda.current_stream()
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
        x_buf = [torch.empty_like(xd) for _ in range(2)]
        c_buf = [torch.empty_like(cd) for _ in range(2)]
        y_pins = [y_pin, torch.empty_like(y_pin).pin_memory()]
        free_ev = [torch.cuda.Event() for _ in range(2)]
        ready_ev = [torch.cuda.Event() for _ in range(2)]
        out_ev = [torch.cuda.Event() for _ in range(2)]
        for ev in free_ev + out_ev:
            ev.record(compute)

        def e2e_step(i):
            k = i & 1
            with torch.cuda.stream(s_in):
                s_in.wait_event(free_ev[k])
                x_buf[k].copy_(x_pin, non_blocking=True)
                c_buf[k].copy_(c_pin, non_blocking=True)
                ready_ev[k].record(s_in)
            compute.wait_event(ready_ev[k])
            y = model(x_buf[k], time=td, condition=c_buf[k])
            free_ev[k].record(compute)
            done = torch.cuda.Event()
            done.record(compute)
            with torch.cuda.stream(s_out):
                s_out.wait_event(done)
                s_out.wait_event(out_ev[k])
                y_pins[k].copy_(y, non_blocking=True)
                y.record_stream(s_out)
                out_ev[k].record(s_out)

        for i in range(4):
            e2e_step(i)
        barrier()
        e0.record()
        for i in range(args.steps):
            e2e_step(i)
        compute.wait_stream(s_out)
        e1.record()
        barrier()
        ms_e2e = max_over_ranks(e0.elapsed_time(e1))

    samples = world * B * args.steps
    value = samples / (ms_dev * 1e-3)
    e2e_value = samples / (ms_e2e * 1e-3)
    pk = peaks()

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"bf16": "bf16", "tf32": "tf32", "fp32": "f32"}[args.precision], "data": "synthetic", "config": workload_config(args),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(x.numel() * 4 + c.numel() * 4),
                "d2h_bytes_per_step": int(y_pin.numel() * 4)},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "host": {"numa_node_of_gpu": numa_node, "cpus": len(os.sched_getaffinity(0))},
        "derived": {
            "steps_per_s_per_member": value / world / B / FORWARDS_PER_STEP,
            "member_sypd": value / FORWARDS_PER_STEP * 86400.0 / STEPS_PER_YEAR,
            "model_tflops": value * FLOP_PER_SAMPLE / 1e12,
            "model_flops_frac_of_bf16_peak": value / world * FLOP_PER_SAMPLE / 1e12 / pk["bf16_tflops_sustained"],
        },
    }

    # ---- BASELINE.json configs[3]: sharded 25-member ensemble rollout with the NCCL statistics gather (all ranks) ----
    if not args.no_rollout:
        del x_buf, c_buf
        torch.cuda.empty_cache()
        try:
            line["ensemble_rollout"] = run_rollout(args, dev, world, rank)
        except Exception as exc:
            line["ensemble_rollout"] = {"error": f"{type(exc).__name__}: {exc}"}

    if rank == 0:
        # ---- roofline of the dominant kernel, timed live by the library with CUDA events on the launch stream ----
        try:
            line["roofline"] = model_roofline(model, xd, td, cd, pk)
        except Exception as exc:  # keep the bench line even if the profile hook fails
            line["roofline"] = {"error": str(exc)}
        # ---- CPU baseline: the oracle port on the host cores, bounded sample ---------------------------------------
        if world == 1 and not args.no_cpu_baseline:
            if prev_affinity:
                os.sched_setaffinity(0, prev_affinity)   # the CPU baseline may use every core the process was given
            torch.set_num_threads(os.cpu_count() or 1)
            cfg1, sd1, x1, c1, t1 = build_case(1)
            times = cpu_forward_time(cfg1, sd1, x1, c1, t1, steps=2, warmup=1)
            line["cpu_baseline"] = {"value": 1.0 / min(times), "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                    "sample": "best of 2 forwards of 1 sample after 1 warm-up (oracle/sfno_oracle.py, fp32 torch CPU)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_rollout(args, dev, world, rank):
    """BASELINE.json configs[3]: a 25-member ensemble rollout sharded member-major over the ranks (1/2/4/8 GPUs), every
    rank advancing its members as one batch through DYffusion sampling windows (6 forecaster + 10 interpolator forwards
    per window, interpolator dropout live, one captured CUDA graph per window), with the per-step ensemble statistics of
    src/evaluation/metrics.py:166-246 (mean, spread, RMSE, spread-skill ratio, fair CRPS) INSIDE the timed region:
    NCCL all-to-all of grid-point slices of the members + one fused statistics kernel on the slice per 6-hour step.  The reference loops over the members
    sequentially at batch 1 (src/ace_inference/inference/loop.py:199-208)."""
    import torch.distributed as dist

    from spherical_dyffusion_b200 import configs
    from spherical_dyffusion_b200.dyffusion import DYffusion
    from spherical_dyffusion_b200.ensemble import EnsembleStatistics, area_weights, max_local_members
    from spherical_dyffusion_b200.rollout import EnsembleRollout

    members, windows, h = args.rollout_members, args.rollout_windows, 6
    kw = dict(precision=args.precision, check_time_range=False, param_check="version")
    fore = configs.build(configs.ACE_FORECASTER, seed=0, min_max_time=(0.0, 5.0), **kw).to(dev).eval()
    ipol = configs.build(configs.ACE_INTERPOLATOR, seed=1, min_max_time=(1.0, 5.0), **kw).to(dev).eval()
    dy = DYffusion(fore, ipol, timesteps=h, forward_conditioning="none", time_encoding="dynamics",
                   capture_graph=not args.no_graph, graph_outputs="view")
    stats = EnsembleStatistics(members)
    weights = area_weights(torch.linspace(-89.5, 89.5, 180), 360).to(dev)
    g = torch.Generator().manual_seed(0)   # the same initial condition, forcing and verification field on every rank
    ic = torch.randn(34, 180, 360, generator=g).to(dev)
    base_forcing = torch.randn(2, 180, 360, generator=g).to(dev)
    truth = torch.randn(34, 180, 360, generator=g).to(dev)

    def forcing_fn(step, n, d):
        return (base_forcing * (1.0 + 0.01 * step)).unsqueeze(0).expand(n, -1, -1, -1).contiguous()

    ro = EnsembleRollout(dy, stats, forcing_fn, truth_fn=lambda s, d: truth, weights=weights)
    ro.run(ic, n_steps=h)    # warm-up window (creates the nets, captures the window graph)
    ro.run(ic, n_steps=h)
    ro.time_stats = True
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    hist = ro.run(ic, n_steps=windows * h)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1), ro.stats_ms()], device=dev, dtype=torch.float64)
    t_min = t.clone()
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(t_min, op=dist.ReduceOp.MIN)
    # the statistics step of a rank with fewer members starts early and then waits inside the collective for the rank with
    # the most members: the MAX over ranks is exchange + kernel + that wait, the MIN is what the critical-path rank pays
    ms, ms_stats, ms_stats_min = float(t[0]), float(t[1]), float(t_min[1])
    steps = windows * h
    steps_per_s = steps / (ms * 1e-3)
    local = max_local_members(members, world)
    out = {
        "workload": f"{members}-member ensemble rollout, {windows} windows = {steps} x 6-h steps, member-major over {world} GPU(s), "
                    f"interpolator dropout on, per-step mean/spread/rmse/ssr/fair-CRPS (all-to-all of grid-point slices + fused kernel on the slice + all-gather of the result maps) inside the timed region",
        "members": members, "windows": windows, "steps": steps, "max_local_members": local,
        "ideal_speedup_vs_1gpu": members / local, "ms_total": ms, "ensemble_steps_per_s": steps_per_s,
        "ensemble_sypd": steps_per_s * 86400.0 / STEPS_PER_YEAR, "member_steps_per_s_aggregate": steps_per_s * members,
        "statistics_ms_total": ms_stats, "statistics_share": ms_stats / ms,
        "statistics_ms_critical_rank": ms_stats_min, "statistics_share_critical_rank": ms_stats_min / ms,
        "statistics_bytes_received_per_rank_per_step": int(members * 34 * 180 * 360 * 4 // world) if world > 1 else 0,
        "window_graph": bool(not args.no_graph), "forwards_per_window": dy.forwards_per_window(),
        "last_crps_mean": float(hist["crps"][-1].mean()), "last_spread_mean": float(hist["spread"][-1].mean()),
    }
    dy.release_graphs()
    del ro, dy, fore, ipol
    torch.cuda.empty_cache()
    return out


def model_roofline(model, xd, td, cd, pk):
    """Per-kernel device times of one forward (CUDA events inside the library) -> roofline of the dominant kernel."""
    from spherical_dyffusion_b200.profile import profile_forward, roofline_from_profile

    recs = profile_forward(model, xd, td, cd)
    return roofline_from_profile(recs, pk, model=model, batch=xd.shape[0])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32", "tf32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-rollout", action="store_true", help="skip the ensemble-rollout leg (configs[3])")
    ap.add_argument("--no-graph", action="store_true", help="ensemble rollout: eager windows instead of one CUDA graph per window")
    ap.add_argument("--rollout-members", type=int, default=25)
    ap.add_argument("--rollout-windows", type=int, default=10)
    args = ap.parse_args()
    if args.impl == "reference":
        args.steps = args.steps if args.steps is not None else 3
        args.warmup = args.warmup if args.warmup is not None else 1
        args.steps = min(args.steps, 8)  # bounded: one step is ~5-10 s of CPU work
        args.warmup = min(args.warmup, 2)
        run_reference(args)
    else:
        args.steps = args.steps if args.steps is not None else 100   # ~1.2 s timed region: >= 10 clock samples
        args.warmup = args.warmup if args.warmup is not None else 5
        run_b200(args)


if __name__ == "__main__":
    main()
