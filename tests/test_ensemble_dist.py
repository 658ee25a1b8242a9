"""Multi-rank host logic on CPU: member sharding and the statistics collectives (world_size 2 and 3, gloo).
The local arithmetic is replaced by a torch stand-in (the product's CudaEnsembleOps needs a GPU); what is under
test is the sharding, padding, all-reduce / all-gather plumbing and the metric definitions of metrics.py."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from spherical_dyffusion_b200.ensemble import EnsembleStatistics, area_weights, max_local_members, member_shard


class TorchOps:
    """torch stand-in with the interface of CudaEnsembleOps (two-pass / Chan moments, pairwise CRPS)."""

    def local_sum(self, members):
        return members.sum(0)

    def shifted_moments(self, members, sum_global, total_members):
        d = members - sum_global / total_members
        return torch.stack([d.sum(0), (d * d).sum(0)])

    def finalize(self, sum_global, mom, total):
        return sum_global / total + mom[0] / total, (mom[1] - mom[0] ** 2 / total) / (total - 1)

    def crps(self, members, truth):
        E = members.shape[0]
        skill = (members - truth).abs().mean(0)
        spread = (members[None] - members[:, None]).abs().sum((0, 1)) / (E * (E - 1))
        return skill - 0.5 * spread

    def stats(self, members, truth, rows=None):
        if rows is not None:
            members = members[rows.long()]
        return members.mean(0), members.var(dim=0), (self.crps(members, truth) if truth is not None else None)


def test_member_shard_partition():
    for E in (1, 5, 25, 32):
        for G in (1, 2, 4, 8):
            shards = [member_shard(E, G, r) for r in range(G)]
            assert sorted(sum(shards, [])) == list(range(E))
            assert max(len(s) for s in shards) == max_local_members(E, G)
    assert [len(member_shard(25, 8, r)) for r in range(8)] == [4, 3, 3, 3, 3, 3, 3, 3]


def _reference_metrics(members, truth, weights):
    # metrics.py:166-246 restated with torch for the comparison
    def wmean(x):
        return (x * weights).sum((-2, -1)) / weights.expand(x.shape).sum((-2, -1))
    E = members.shape[0]
    mean = members.mean(0)
    spread = torch.sqrt(wmean(members.var(dim=0)))
    rmse = torch.sqrt(wmean((mean - truth) ** 2))
    skill = (members - truth).abs().mean(0)
    sp = (members[None] - members[:, None]).abs().sum((0, 1)) / (E * (E - 1))
    return dict(mean=mean, spread=spread, rmse=rmse, ssr=spread * ((E + 1) / E) ** 0.5 / rmse, crps=wmean(skill - 0.5 * sp))


def _worker(rank, world, port, E, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        members = torch.randn(E, 3, 6, 12, generator=g) * 2 + 1
        truth = torch.randn(3, 6, 12, generator=g)
        weights = area_weights(torch.linspace(-80, 80, 6), 12)
        stats = EnsembleStatistics(E, ops=TorchOps())
        local = members[stats.local_ids]
        out = stats.step(local, truth=truth, weights=weights)
        ref = _reference_metrics(members, truth, weights)
        ok = all(torch.allclose(out[k], ref[k], rtol=1e-4, atol=1e-5) for k in ref)
        # a field whose point count (3 * 5 * 7 = 105) does not divide into equal 16-byte-aligned slices: 56 + 49
        m2 = torch.randn(E, 3, 5, 7, generator=g) - 0.5
        t2 = torch.randn(3, 5, 7, generator=g)
        ref2 = _reference_metrics(m2, t2, torch.ones(5, 7))
        o2 = stats.step(m2[stats.local_ids], truth=t2, weights=None)
        ok = ok and all(torch.allclose(o2[k], ref2[k], rtol=1e-4, atol=1e-5) for k in ref2)
        buf, rows, lo, hi = stats.exchange_slices(local)
        ns = buf.shape[1]
        ok = ok and (lo, hi) == (rank * ns, min((rank + 1) * ns, 216))
        ok = ok and torch.equal(buf[rows.long()].sort(dim=0).values[:, : hi - lo], members.reshape(E, -1)[:, lo:hi].sort(dim=0).values)
        # without a truth field: no gather, two all-reduces (sums -> pivot, shifted moments); a large offset must not hurt
        big = local + 1.0e5
        nt = stats.step(big, weights=weights)
        ok = ok and torch.allclose(nt["var"], (members + 1.0e5).var(dim=0), rtol=2e-3, atol=1e-4)
        ok = ok and torch.allclose(nt["mean"], (members + 1.0e5).mean(0), rtol=1e-6)
        gathered = stats.gather_members(local)
        ok = ok and torch.equal(gathered, members)
        results[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("E,world", [(5, 2), (8, 2), (7, 3)])
def test_statistics_multi_rank_gloo(E, world):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, port, E, results), nprocs=world, join=True)
    assert dict(results) == {r: True for r in range(world)}


def test_statistics_single_process_matches_metrics():
    g = torch.Generator().manual_seed(1)
    members = torch.randn(7, 2, 5, 10, generator=g)
    truth = torch.randn(2, 5, 10, generator=g)
    weights = area_weights(torch.linspace(-60, 60, 5), 10)
    out = EnsembleStatistics(7, ops=TorchOps()).step(members, truth=truth, weights=weights)
    ref = _reference_metrics(members, truth, weights)
    for k in ref:
        assert torch.allclose(out[k], ref[k], rtol=1e-4, atol=1e-5), k


# ---- pinned against the reference's own metric functions (tests/golden/make_golden_metrics.py) -------------------------
METRIC_CASES = ["e2", "e5_ties", "e8", "e25"]


def _metrics_fixture():
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ensemble_metrics.pt")
    return torch.load(path, map_location="cpu", weights_only=False)


class SortedFormOps(TorchOps):
    """The arithmetic of ``ensemble_crps_kernel`` (csrc/ensemble.cu) restated with torch: the pairwise term through the
    order statistics, sum_{i<j} |x_i - x_j| = sum_k (2k - E + 1) x_(k)."""

    def crps(self, members, truth):
        E = members.shape[0]
        skill = (members - truth).abs().sum(0) / E
        xs = members.sort(dim=0).values
        coef = (2 * torch.arange(E, dtype=members.dtype) - E + 1).reshape(E, *([1] * (members.dim() - 1)))
        return skill - (coef * xs).sum(0) / (E * (E - 1))


@pytest.mark.parametrize("case", METRIC_CASES)
@pytest.mark.parametrize("ops", [TorchOps, SortedFormOps])
def test_statistics_match_reference_metric_functions(case, ops):
    fx = _metrics_fixture()[case]
    E = fx["spec"]["E"]
    ref = fx["ref"]
    weights = area_weights(fx["lats"], fx["members"].shape[-1])
    assert torch.equal(weights, ref["weights"])
    out = EnsembleStatistics(E, ops=ops()).step(fx["members"], truth=fx["truth"], weights=weights)
    for k in ("mean", "spread", "rmse", "ssr", "crps"):
        assert torch.allclose(out[k], ref[k], rtol=2e-5, atol=1e-6), (k, out[k], ref[k])
    unweighted = EnsembleStatistics(E, ops=ops()).step(fx["members"], truth=fx["truth"], weights=None)
    assert torch.allclose(unweighted["crps"], ref["crps_unweighted"], rtol=2e-5, atol=1e-6)
    flat = fx["members"].reshape(E, -1)
    pointwise = ops().crps(flat, fx["truth"].reshape(-1)).reshape(fx["truth"].shape)
    assert torch.allclose(pointwise, ref["crps_pointwise"], rtol=1e-4, atol=2e-6)


def _fixture_worker(rank, world, port, case, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        fx = _metrics_fixture()[case]
        stats = EnsembleStatistics(fx["spec"]["E"], ops=SortedFormOps())
        out = stats.step(fx["members"][stats.local_ids], truth=fx["truth"], weights=fx["ref"]["weights"])
        results[rank] = all(bool(torch.allclose(out[k], fx["ref"][k], rtol=2e-5, atol=1e-6))
                            for k in ("mean", "spread", "rmse", "ssr", "crps"))
    finally:
        dist.destroy_process_group()


def test_sharded_statistics_match_reference_metric_functions_two_ranks():
    """25 members over two ranks (13 + 12, uneven shards) reproduce the reference's numbers."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_fixture_worker, args=(2, port, "e25", results), nprocs=2, join=True)
    assert dict(results) == {0: True, 1: True}
