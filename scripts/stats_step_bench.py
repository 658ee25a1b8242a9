"""Micro-benchmark of one ensemble-statistics step (25 members, 34 x 180 x 360) on N GPUs under torchrun:
(a) all-gather of the members + fused kernel on all grid points on every rank, (b) all-to-all of grid-point slices +
fused kernel on the slice + all-gather of the three result maps.  Device time, barrier before every iteration, max over ranks."""
import json
import os
import sys

sys.path.insert(0, ".")
import torch
import torch.distributed as dist

from spherical_dyffusion_b200.ensemble import EnsembleStatistics, area_weights

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
E = 25
stats = EnsembleStatistics(E)
g = torch.Generator().manual_seed(rank)
local = torch.randn(len(stats.local_ids), 34, 180, 360, generator=g).to(dev)
truth = torch.randn(34, 180, 360, generator=torch.Generator().manual_seed(99)).to(dev)
weights = area_weights(torch.linspace(-89.5, 89.5, 180), 360).to(dev)


def gather_variant():
    members, rows = stats.gather_padded(local)
    flat = members.reshape(members.shape[0], -1)
    mean, var, crps = stats.ops.stats(flat, truth.reshape(-1), rows)
    out = stats._scalars(mean.reshape(truth.shape), var.reshape(truth.shape), crps.reshape(truth.shape), truth, weights)
    return out


def sliced_variant():
    return stats.step(local, truth=truth, weights=weights)


res = {}
ref = None
for name, fn in (("all_gather", gather_variant), ("all_to_all_slices", sliced_variant)):
    for _ in range(3):
        out = fn()
    if ref is None:
        ref = out
    else:
        for k in ("spread", "rmse", "ssr", "crps"):
            assert torch.allclose(out[k], ref[k], rtol=2e-5, atol=1e-7), (name, k)
    ms = []
    for _ in range(20):
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    t = torch.tensor(sorted(ms)[len(ms) // 2], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res[name] = round(float(t), 3)
if rank == 0:
    print(json.dumps({"n_gpus": world, "members": E, "ms_per_statistics_step_median_max_over_ranks": res}))
dist.destroy_process_group()
