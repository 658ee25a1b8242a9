"""Ensemble sharding and ensemble statistics across GPUs.

Members never interact inside the SFNO or the sampler, so rank r simply owns members {r, r+G, ...} and keeps their
state resident on its GPU.  The only exchange is for the statistics of ``src/evaluation/metrics.py``:

* mean / variance (``ensemble_spread`` :166-175, ``spread_skill_ratio`` :178-196) without a truth field:
  ``all_reduce(SUM)`` of the local sums (a common pivot), then of the moments shifted by it
  (``sfno_ensemble_local_sum`` / ``_shifted_moments``) -- the reference's two-pass ``var``, never sum(x^2) - E mean^2;
* with a truth field (fair CRPS, ``crps_ensemble`` :199-246) every member of a grid point has to meet on one rank: an
  ``all_to_all`` hands rank r the grid-point slice r of every member, ONE fused kernel (``sfno_ensemble_stats``: mean,
  two-pass variance, sorted-form CRPS; no [E, E, ...] tensor is ever materialised) runs on that slice, and the three
  result maps are all-gathered.  Per rank that moves E n / G + 3 n values and computes n / G points, against E n and n
  for an all-gather of the members with replicated statistics.

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

    def local_sum(self, members: torch.Tensor) -> torch.Tensor:
        """Sum over this rank's members [n]; members may be empty."""
        assert members.is_cuda and members.dtype == torch.float32 and members.is_contiguous()
        E, n = members.shape
        s = torch.empty(n, dtype=torch.float32, device=members.device)
        _lib.check(_lib.lib().sfno_ensemble_local_sum(members.data_ptr() if E else None, E, n, s.data_ptr(),
                                                      stream_ptr(members.device)), "sfno_ensemble_local_sum")
        return s

    def shifted_moments(self, members: torch.Tensor, sum_global: torch.Tensor, total_members: int) -> torch.Tensor:
        """[2, n]: sum (x - p), sum (x - p)^2 of this rank's members about the common pivot p = sum_global / E."""
        E, n = members.shape
        mom = torch.empty(2, n, dtype=torch.float32, device=members.device)
        _lib.check(_lib.lib().sfno_ensemble_shifted_moments(members.data_ptr() if E else None, E, n, sum_global.data_ptr(),
                                                            total_members, mom.data_ptr(), stream_ptr(members.device)),
                   "sfno_ensemble_shifted_moments")
        return mom

    def finalize(self, sum_global, moments, total_members: int):
        n = sum_global.numel()
        mean = torch.empty(n, dtype=torch.float32, device=sum_global.device)
        var = torch.empty(n, dtype=torch.float32, device=sum_global.device)
        _lib.check(_lib.lib().sfno_ensemble_finalize(sum_global.data_ptr(), moments.data_ptr(), total_members, n, mean.data_ptr(),
                                                     var.data_ptr(), stream_ptr(sum_global.device)), "sfno_ensemble_finalize")
        return mean, var

    def stats(self, members: torch.Tensor, truth: Optional[torch.Tensor], rows: Optional[torch.Tensor] = None):
        """All members [E, n] (after the gather) -> (mean, var, crps or None) in ONE pass over the members.  ``rows``
        (int32 [E], device): the live rows of a padded gather buffer [>= E, n]; the statistics do not depend on the member
        order, so the buffer is used as it arrives."""
        assert members.is_cuda and members.dtype == torch.float32 and members.is_contiguous()
        n = members.shape[1]
        E = members.shape[0] if rows is None else int(rows.numel())
        mean = torch.empty(n, dtype=torch.float32, device=members.device)
        var = torch.empty(n, dtype=torch.float32, device=members.device)
        crps = torch.empty(n, dtype=torch.float32, device=members.device) if truth is not None else None
        _lib.check(_lib.lib().sfno_ensemble_stats_rows(members.data_ptr(), rows.data_ptr() if rows is not None else None,
                                                       truth.data_ptr() if truth is not None else None, E, n, mean.data_ptr(),
                                                       var.data_ptr(), crps.data_ptr() if crps is not None else None,
                                                       stream_ptr(members.device)), "sfno_ensemble_stats_rows")
        return mean, var, crps

    def crps(self, members: torch.Tensor, truth: torch.Tensor) -> torch.Tensor:
        return self.stats(members, truth)[2]


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
        self._rows: Dict[str, torch.Tensor] = {}     # live rows of the padded gather buffer, per device
        self._sendbuf: Dict[tuple, torch.Tensor] = {}

    def mean_var(self, local_members: torch.Tensor):
        """local_members [E_local, ...] fp32 -> (mean [...], unbiased variance [...]) over ALL members, without gathering
        them: all-reduce of the sums (-> common pivot), then of the shifted first / second moments about it."""
        shape = local_members.shape[1:]
        flat = local_members.reshape(local_members.shape[0], -1).contiguous()
        s_glob = self.ops.local_sum(flat)
        if self.world > 1:
            dist.all_reduce(s_glob, op=dist.ReduceOp.SUM, group=self.group)   # -> the common pivot p = sum / E
        mom = self.ops.shifted_moments(flat, s_glob, self.n_members)
        if self.world > 1:
            dist.all_reduce(mom, op=dist.ReduceOp.SUM, group=self.group)
        mean, var = self.ops.finalize(s_glob, mom, self.n_members)
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

    def gather_padded(self, local_members: torch.Tensor):
        """All-gather WITHOUT compaction or re-ordering: returns (buffer [world * k, ...] with k = max local members, int32
        row indices of the live members).  Uneven shards leave padding rows at the end of a rank's block; they are
        never read.  One copy of the local members into the send slot, one collective."""
        if self.world == 1:
            return local_members, None
        k = max_local_members(self.n_members, self.world)
        dev = local_members.device
        key = (str(dev), tuple(local_members.shape[1:]))
        send = self._sendbuf.get(key)
        if send is None:
            send = self._sendbuf[key] = torch.zeros(k, *local_members.shape[1:], dtype=local_members.dtype, device=dev)
        send[: local_members.shape[0]].copy_(local_members)
        out = torch.empty(self.world * k, *local_members.shape[1:], dtype=local_members.dtype, device=dev)
        dist.all_gather_into_tensor(out, send, group=self.group)
        rows = self._rows.get(str(dev))
        if rows is None:
            live = [r * k + j for r in range(self.world) for j in range(len(member_shard(self.n_members, self.world, r)))]
            rows = self._rows[str(dev)] = torch.tensor(live, dtype=torch.int32).to(dev)
        return out, rows

    # ---- statistics with a verification field: every rank owns one SLICE of the grid points -------------------------------
    def _slice_len(self, n: int) -> int:
        return (-(-n // self.world) + 3) // 4 * 4      # ceil(n / world), rounded up to 16 bytes

    def exchange_slices(self, local_members: torch.Tensor):
        """All-to-all: rank r receives, from every rank, that rank's members restricted to grid points
        [r * ns, (r + 1) * ns) of the flattened field.  Returns (buffer [world * k, ns], int32 live rows, lo, hi) with
        [lo, hi) this rank's slice of the n grid points.  Each rank receives E * n / world values instead of the E * n of an
        all-gather, and the fused statistics kernel then runs on n / world points per rank instead of on all n everywhere."""
        W, k = self.world, max_local_members(self.n_members, self.world)
        El = local_members.shape[0]
        flat = local_members.reshape(El, -1)
        n = flat.shape[1]
        ns = self._slice_len(n)
        dev = flat.device
        key = (str(dev), "slices", n)
        send = self._sendbuf.get(key)
        if send is None:
            send = self._sendbuf[key] = torch.zeros(W, k, ns, dtype=flat.dtype, device=dev)   # pad rows / columns stay zero
        if W * ns == n:
            send[:, :El].copy_(flat.view(El, W, ns).transpose(0, 1))                           # one strided copy
        else:
            for r in range(W):
                lo_r, hi_r = min(r * ns, n), min((r + 1) * ns, n)
                if hi_r > lo_r:
                    send[r, :El, : hi_r - lo_r].copy_(flat[:, lo_r:hi_r])
        recv = torch.empty(W, k, ns, dtype=flat.dtype, device=dev)
        dist.all_to_all_single(recv, send, group=self.group)
        rows = self._rows.get(str(dev))
        if rows is None:
            live = [r * k + j for r in range(W) for j in range(len(member_shard(self.n_members, W, r)))]
            rows = self._rows[str(dev)] = torch.tensor(live, dtype=torch.int32).to(dev)
        return recv.view(W * k, ns), rows, min(self.rank * ns, n), min((self.rank + 1) * ns, n)

    def _sliced_step(self, local_members, truth, weights):
        """World > 1, with a truth field: all-to-all of grid-point slices, ONE fused kernel on this rank's slice, then an
        all-gather of the three result maps (3 n values; the members exchanged were E n / world).  Measured on 4 B200s, 25
        members of 34 x 180 x 360: 0.51 ms per step against 0.76 ms for all-gathering the members and computing the
        statistics of every grid point on every rank (profiles/r02_o_stats_step_4gpu.json); reducing only per-channel
        sums instead of gathering the maps was tried and is slower (1.24 ms: the segmented reduction costs more than the
        26 MB gather)."""
        E, W = self.n_members, self.world
        shape = truth.shape
        n = truth.numel()
        dev = truth.device
        members, rows, lo, hi = self.exchange_slices(local_members)
        ns = members.shape[1]
        tflat = truth.reshape(-1)
        if hi - lo == ns:
            tslice = tflat[lo:hi]
        else:
            tslice = torch.zeros(ns, dtype=tflat.dtype, device=dev)
            tslice[: hi - lo] = tflat[lo:hi]
        mean_s, var_s, crps_s = self.ops.stats(members, tslice.contiguous(), rows)
        mine = torch.stack((mean_s, var_s, crps_s))                                            # [3, ns]
        every = torch.empty(W * 3, ns, dtype=mine.dtype, device=dev)
        dist.all_gather_into_tensor(every, mine, group=self.group)            # concatenation along dim 0 (NCCL and Gloo)
        every = every.view(W, 3, ns)
        mean, var, crps = (every[:, i].reshape(-1)[:n].reshape(shape) for i in range(3))
        return self._scalars(mean, var, crps, truth, weights)

    def _scalars(self, mean, var, crps, truth, weights):
        E = self.n_members
        out = {"mean": mean, "var": var, "spread": torch.sqrt(weighted_mean(var, weights))}
        rmse = torch.sqrt(weighted_mean((mean - truth) ** 2, weights))
        out["rmse"] = rmse
        out["ssr"] = out["spread"] * ((E + 1) / E) ** 0.5 / rmse
        out["crps"] = weighted_mean(crps, weights)
        return out

    def step(self, local_members: torch.Tensor, truth: Optional[torch.Tensor] = None,
             weights: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        """Statistics the reference records per time step (``aggregators/timestepwise.py:131-177``):
        ensemble mean, spread = sqrt(weighted mean variance), and with a truth field: RMSE of the mean, spread-skill
        ratio (with the sqrt((E+1)/E) correction) and the fair CRPS.  Spatial dims are the last two."""
        if truth is None:
            mean, var = self.mean_var(local_members)
            return {"mean": mean, "var": var, "spread": torch.sqrt(weighted_mean(var, weights))}
        # with a verification field the fair CRPS needs every member of a grid point on one rank
        if self.world > 1:
            return self._sliced_step(local_members, truth, weights)
        flat = local_members.reshape(local_members.shape[0], -1).contiguous()
        mean, var, crps = self.ops.stats(flat, truth.reshape(-1).contiguous(), None)
        return self._scalars(mean.reshape(truth.shape), var.reshape(truth.shape), crps.reshape(truth.shape), truth, weights)
