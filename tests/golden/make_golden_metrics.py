"""Generate ``tests/golden/ensemble_metrics.pt`` from the reference's OWN metric functions.

Run in the build container only (needs ``/root/reference``):

    python tests/golden/make_golden_metrics.py

``/root/reference/src/evaluation/metrics.py`` imports nothing but numpy / torch / typing_extensions, so the file is
loaded unchanged (by path, without the package around it).  For a few seeded ensembles the fixture stores the members,
the truth field, the latitudes and what the reference computes from them: ``spherical_area_weights`` (:15-29),
the ensemble mean, ``ensemble_spread`` (:166-175), ``root_mean_squared_error`` of the mean, ``spread_skill_ratio``
(:178-196) and ``crps_ensemble`` (:199-246; reduced and per grid point), all reduced over the two spatial dims.
"""
from __future__ import annotations

import importlib.util
import os

import torch

OUT = os.path.dirname(os.path.abspath(__file__))
REF_METRICS = "/root/reference/src/evaluation/metrics.py"

# E members, C variables on an H x W grid; offset / scale of the members; some cases carry duplicated members (ties in
# the sorted-form CRPS) and a two-member ensemble (smallest E the fair CRPS is defined for)
CASES = {
    "e2": dict(E=2, C=2, H=6, W=12, seed=0, loc=0.0, scale=1.0),
    "e5_ties": dict(E=5, C=3, H=8, W=16, seed=1, loc=1.0, scale=2.0, duplicate=(0, 3)),
    "e8": dict(E=8, C=3, H=10, W=20, seed=2, loc=-0.5, scale=0.3),
    "e25": dict(E=25, C=4, H=12, W=24, seed=3, loc=1.0, scale=2.0),
}


def load_reference_metrics():
    spec = importlib.util.spec_from_file_location("ref_metrics", REF_METRICS)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    m = load_reference_metrics()
    out = {}
    for name, c in CASES.items():
        g = torch.Generator().manual_seed(c["seed"])
        E, C, H, W = c["E"], c["C"], c["H"], c["W"]
        members = torch.randn(E, C, H, W, generator=g) * c["scale"] + c["loc"]
        if "duplicate" in c:
            a, b = c["duplicate"]
            members[b] = members[a]
        truth = torch.randn(C, H, W, generator=g) * c["scale"] + c["loc"]
        lats = torch.linspace(-90 + 90 / H, 90 - 90 / H, H)
        weights = m.spherical_area_weights(lats, W)
        dim = (-2, -1)
        ref = dict(
            weights=weights,
            mean=members.mean(dim=0),
            spread=m.ensemble_spread(members, weights=weights, dim=dim),
            rmse=m.root_mean_squared_error(truth, members.mean(dim=0), weights=weights, dim=dim),
            ssr=m.spread_skill_ratio(truth, members, weights=weights, dim=dim),
            crps=m.crps_ensemble(truth, members, weights=weights, dim=dim),
            crps_pointwise=m.crps_ensemble(truth, members, reduction="none"),
            crps_unweighted=m.crps_ensemble(truth, members, dim=dim),
        )
        out[name] = dict(spec=c, members=members, truth=truth, lats=lats, ref=ref)
        print(name, {k: tuple(v.shape) for k, v in ref.items()}, "crps", [round(float(x), 5) for x in ref["crps"]])
    path = os.path.join(OUT, "ensemble_metrics.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
