"""Ensemble sharding and ensemble statistics across GPUs.

Members never interact inside the SFNO or the sampler, so rank r simply owns members {r, r+G, ...} and keeps their
state resident on its GPU.  The only exchange is for the statistics of ``src/evaluation/metrics.py``:

* mean / variance (``ensemble_spread`` :166-175, ``spread_skill_ratio`` :178-196): ``all_reduce(SUM)`` of
  [sum x, sum x^2] accumulated locally by ``sfno_ensemble_accumulate``;
* fair CRPS (``crps_ensemble`` :199-246): ``all_gather`` of the members, then the sorted-form kernel
  ``sfno_ensemble_crps`` (no [E, E, ...] tensor is ever materialised).

Collectives are plain ``torch.distributed`` (NCCL over NVLink on the GPU box, Gloo in the CPU tests).  The local
arithmetic runs through the C ABI on the GPU; ``ops`` exists so the CPU tests can exercise the sharding / collective
plumbing with a torch stand-in -- the default and only product implementation is ``CudaEnsembleOps``.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.distributed as dist

from . import _lib
from ._util import stream_ptr


def member_shard(n_members: int, world_size: int, rank: int) -> List[int]:
    """Member-major partition: rank r owns members r, r+G, r+2G, ... (max local count = ceil(E/G))."""
    return list(range(rank, n_members, world_size))


def max_local_members(n_members: int, world_size: int) -> int:
    return (n_members + world_size - 1) // world_size


class CudaEnsembleOps:
    """Local statistics kernels of libsfno_b200 (fp32, flattened [members, n])."""

    def accumulate(self, members: torch.Tensor, sums: torch.Tensor) -> None:
        assert members.is_cuda and members.dtype == torch.float32 and members.is_contiguous()
        E, n = members.shape
        _lib.check(_lib.lib().sfno_ensemble_accumulate(members.data_ptr(), E, n, sums.data_ptr(), stream_ptr(members.device)),
                   "sfno_ensemble_accumulate")

    def finalize(self, sums: torch.Tensor, total_members: int):
        n = sums.shape[1]
        mean = torch.empty(n, dtype=torch.float32, device=sums.device)
        var = torch.empty(n, dtype=torch.float32, device=sums.device)
        _lib.check(_lib.lib().sfno_ensemble_finalize(sums.data_ptr(), total_members, n, mean.data_ptr(), var.data_ptr(),
                                                     stream_ptr(sums.device)), "sfno_ensemble_finalize")
        return mean, var

    def crps(self, members: torch.Tensor, truth: torch.Tensor) -> torch.Tensor:
        E, n = members.shape
        out = torch.empty(n, dtype=torch.float32, device=members.device)
        _lib.check(_lib.lib().sfno_ensemble_crps(members.data_ptr(), truth.data_ptr(), E, n, out.data_ptr(),
                                                 stream_ptr(members.device)), "sfno_ensemble_crps")
        return out


def weighted_mean(x: torch.Tensor, weights: Optional[torch.Tensor]) -> torch.Tensor:
    """``metrics.py:32-57`` over the two spatial dims of x [..., H, W]."""
    if weights is None:
        return x.mean(dim=(-2, -1))
    return (x * weights).sum(dim=(-2, -1)) / weights.expand(x.shape).sum(dim=(-2, -1))


def area_weights(lats_deg: torch.Tensor, num_lon: int) -> torch.Tensor:
    """``spherical_area_weights`` (``metrics.py:15-29``)."""
    w = torch.cos(torch.deg2rad(lats_deg)).repeat(num_lon, 1).t()
    return w / w.sum()


class EnsembleStatistics:
    """Per-step ensemble statistics of members sharded over the ranks of ``group``."""

    def __init__(self, n_members: int, group=None, ops=None):
        self.n_members = n_members
        self.group = group
        self.ops = ops if ops is not None else CudaEnsembleOps()
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        self.local_ids = member_shard(n_members, self.world, self.rank)

    def mean_var(self, local_members: torch.Tensor):
        """local_members [E_local, ...] fp32 -> (mean [...], unbiased variance [...]) over ALL members."""
        shape = local_members.shape[1:]
        flat = local_members.reshape(local_members.shape[0], -1).contiguous()
        n = flat.shape[1]
        sums = torch.zeros(2, n, dtype=torch.float32, device=flat.device)
        if flat.shape[0] > 0:
            self.ops.accumulate(flat, sums)
        if self.world > 1:
            dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=self.group)
        mean, var = self.ops.finalize(sums, self.n_members)
        return mean.reshape(shape), var.reshape(shape)

    def gather_members(self, local_members: torch.Tensor) -> torch.Tensor:
        """All members [E, ...] in member order on every rank (uneven shards are padded for the collective)."""
        if self.world == 1:
            return local_members
        k = max_local_members(self.n_members, self.world)
        pad = torch.zeros(k, *local_members.shape[1:], dtype=local_members.dtype, device=local_members.device)
        pad[: local_members.shape[0]] = local_members
        out = torch.empty(self.world * k, *pad.shape[1:], dtype=pad.dtype, device=pad.device)
        dist.all_gather_into_tensor(out, pad, group=self.group)  # concatenation along dim 0 (NCCL and Gloo)
        out = out.view(self.world, k, *pad.shape[1:])
        members = torch.empty(self.n_members, *local_members.shape[1:], dtype=pad.dtype, device=pad.device)
        for r in range(self.world):
            ids = member_shard(self.n_members, self.world, r)
            if ids:
                members[ids] = out[r, : len(ids)]
        return members

    def step(self, local_members: torch.Tensor, truth: Optional[torch.Tensor] = None,
             weights: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        """Statistics the reference records per time step (``aggregators/timestepwise.py:131-177``):
        ensemble mean, spread = sqrt(weighted mean variance), and with a truth field: RMSE of the mean, spread-skill
        ratio (with the sqrt((E+1)/E) correction) and the fair CRPS.  Spatial dims are the last two."""
        mean, var = self.mean_var(local_members)
        out = {"mean": mean, "var": var, "spread": torch.sqrt(weighted_mean(var, weights))}
        if truth is not None:
            E = self.n_members
            rmse = torch.sqrt(weighted_mean((mean - truth) ** 2, weights))
            out["rmse"] = rmse
            out["ssr"] = out["spread"] * ((E + 1) / E) ** 0.5 / rmse
            members = self.gather_members(local_members)
            flat = members.reshape(E, -1).contiguous()
            crps = self.ops.crps(flat, truth.reshape(-1).contiguous()).reshape(truth.shape)
            out["crps"] = weighted_mean(crps, weights)
        return out
