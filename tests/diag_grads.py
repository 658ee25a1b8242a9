"""Per-parameter gradient errors of the differentiable forward vs an fp64 oracle (and the fp32 CPU oracle's own error)."""
import sys
sys.path.insert(0, "tests"); sys.path.insert(0, ".")  # run from the repository root: python tests/diag_grads.py fp32
import torch
from oracle.sfno_oracle import SFNOOracle, perturb_affine_and_biases, random_state_dict, rel_l2
import spherical_dyffusion_b200 as sb
import test_gpu_backward as t

dev = torch.device("cuda:0")
for case in sys.argv[2:] or sorted(t.NET_CASES):
    cfg = t.NET_CASES[case]
    precision = sys.argv[1]
    sd = perturb_affine_and_biases(random_state_dict(cfg, seed=3, spectral_gain=4.0), seed=4)
    if cfg.normalization_layer == "none":
        sd = {k: v for k, v in sd.items() if ".norm0." not in k and ".norm1." not in k}
    g = torch.Generator().manual_seed(11)
    B = 3
    H, W = cfg.spatial_shape
    x = torch.randn(B, cfg.num_input_channels, H, W, generator=g)
    cond = torch.randn(B, cfg.num_conditional_channels, H, W, generator=g) if cfg.num_conditional_channels else None
    target = torch.randn(B, cfg.num_output_channels, H, W, generator=g)
    time = torch.tensor([1.0, 2.0, 5.0]) if cfg.with_time_emb else None
    loss = "l1" if "time" in case else "mse"
    _, gx32, gp32, _ = t._oracle_grads(cfg, sd, x, target, time, cond, loss)
    o = SFNOOracle(cfg, sd, dtype=torch.float64)
    o.sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in o.sd.items()}
    xr = x.double().requires_grad_(True)
    out = o._forward(xr, time=time, condition=cond)
    l = (out - target.double()).abs().mean() if loss == "l1" else (out - target.double()).square().mean()
    l.backward()
    gp64 = {k: v.grad for k, v in o.sd.items() if torch.is_tensor(v) and v.requires_grad}
    m = sb.SphericalFourierNeuralOperatorNet(
        num_input_channels=cfg.num_input_channels, num_output_channels=cfg.num_output_channels,
        num_output_channels_raw=cfg.num_output_channels, num_conditional_channels=cfg.num_conditional_channels,
        spatial_shape_in=cfg.spatial_shape, spatial_shape_out=cfg.spatial_shape, precision=precision, loss_function=loss,
        **cfg.model_kwargs())
    m.load_state_dict(sd, strict=True)
    if cfg.with_time_emb:
        m.set_min_max_time(cfg.min_time, cfg.max_time)
    m = m.to(dev).train()
    xd = x.to(dev).requires_grad_(True)
    kw = {"time": time.to(dev)} if time is not None else {}
    ld = m.get_loss(xd, target.to(dev), condition=None if cond is None else cond.to(dev), **kw)
    ld["loss"].backward()
    print(f"== {case} {precision}")
    print(f"{'input':40s} norm {float(xr.grad.norm()):.3e} gpu {rel_l2(xd.grad, xr.grad):.3e} cpu32 {rel_l2(gx32, xr.grad):.3e}")
    for name, p in m.named_parameters():
        r = gp64[name].reshape(p.shape)
        n = float(r.norm())
        eg = float((p.grad.cpu().double() - r).norm())
        ec = float((gp32[name].reshape(p.shape).double() - r).norm())
        print(f"{name:40s} norm {n:.3e} numel {p.numel():6d} gpu abs {eg:.3e} rel {eg / max(n, 1e-300):.3e} | cpu32 abs {ec:.3e} rel {ec / max(n, 1e-300):.3e}")
