import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench_extra import _ace_pair
from spherical_dyffusion_b200.dyffusion import DYffusion
from spherical_dyffusion_b200.ensemble import EnsembleStatistics, area_weights
dev = torch.device("cuda:0")
fore, ipol = _ace_pair(dev, "bf16")
dy = DYffusion(fore, ipol, timesteps=6)
g = torch.Generator().manual_seed(0)
ic = torch.randn(34, 180, 360, generator=g).to(dev)
forc = torch.randn(2, 180, 360, generator=g).to(dev)
truth = torch.randn(34, 180, 360, generator=g).to(dev)
for E in (8, 25):
    state = ic.unsqueeze(0).expand(E, -1, -1, -1).contiguous()
    f = forc.unsqueeze(0).expand(E, -1, -1, -1).contiguous()
    with torch.inference_mode():
        y = fore(state, time=torch.zeros(E, device=dev), static_condition=f)
        print(E, "forecaster finite", bool(torch.isfinite(y).all()), float(y.float().std()))
        xi = torch.cat([state, y], 1)
        with ipol.inference_dropout_scope(True):
            z = ipol(xi, time=torch.full((E,), 1.0, device=dev), static_condition=f)
        print(E, "interpolator(dropout) finite", bool(torch.isfinite(z).all()), float(z.float().std()), "member diff", float((z[0]-z[1]).abs().max()))
        z2 = ipol(xi, time=torch.full((E,), 1.0, device=dev), static_condition=f)
        print(E, "interpolator(no dropout) finite", bool(torch.isfinite(z2).all()))
        preds = dy.sample(state, static_condition=f)
        for k, v in preds.items():
            print(E, k, "finite", bool(torch.isfinite(v).all()), float(v.float().std()))
        st = EnsembleStatistics(E).step(preds["t6_preds"].float(), truth=truth, weights=area_weights(torch.linspace(-89.5, 89.5, 180), 360).to(dev))
        print(E, {k: float(v.mean()) for k, v in st.items()})
