"""Backward pass on the GPU: the autograd formulas registered for the custom ops (``ops.py``) against CPU
autograd through the oracle (``oracle/harmonics.py``, ``oracle/sfno_oracle.py``) on identical seeded inputs.  fp32 mode
gradients are held to the same <= 1e-4 relative-L2 bar as the forward; the tensor-core modes to their forward bounds."""
import math

import pytest
import torch

from oracle import harmonics as oh
from oracle.sfno_oracle import SFNOConfig, SFNOOracle, dhconv_contract, diagonal_contract, instance_norm, perturb_affine_and_biases, \
    random_state_dict, rel_l2

import spherical_dyffusion_b200 as sb
from spherical_dyffusion_b200 import _lib

pytestmark = pytest.mark.gpu

GRAD_TOL = 1e-4          # fp32 engine, any gradient
BF16_GRAD_BOUND = 8e-3   # adjoint transforms on the bf16 tensor-core engine (forward op bound: 5.7e-3; two roundings of the cotangent)
TF32_GRAD_BOUND = 1.2e-3


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def _leaf(t, dev=None):
    t = t.detach().clone()
    if dev is not None:
        t = t.to(dev)
    return t.requires_grad_(True)


# ---------------------------------------------------------------------------------------------------------------------------
# transforms
# ---------------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision,tol", [("fp32", GRAD_TOL), ("tf32", TF32_GRAD_BOUND), ("bf16", BF16_GRAD_BOUND)])
@pytest.mark.parametrize("grid,nlat,nlon,lmax,mmax,lead", [
    ("legendre-gauss", 12, 24, 12, 13, (2, 3)),
    ("equiangular", 16, 32, 16, 17, (5,)),
    ("legendre-gauss", 33, 64, 20, 21, (1, 7)),       # truncated modes
    ("equiangular", 180, 360, 180, 181, (1, 16)),     # ACE grid
])
def test_sht_gradients_match_oracle_autograd(dev, precision, tol, grid, nlat, nlon, lmax, mmax, lead):
    g = torch.Generator().manual_seed(nlat + 3 * len(lead))
    x = torch.randn(*lead, nlat, nlon, generator=g)
    cot_X = torch.randn(*lead, lmax, mmax, 2, generator=g)          # cotangent of the coefficient pairs
    coeffs = torch.randn(*lead, lmax, mmax, 2, generator=g)
    cot_x = torch.randn(*lead, nlat, nlon, generator=g)
    o_sht = oh.RealSHT(nlat, nlon, lmax=lmax, mmax=mmax, grid=grid).float()
    o_isht = oh.InverseRealSHT(nlat, nlon, lmax=lmax, mmax=mmax, grid=grid).float()
    xr = _leaf(x)
    (torch.view_as_real(o_sht(xr)) * cot_X).sum().backward()
    cr = _leaf(coeffs)
    (o_isht(torch.view_as_complex(cr)) * cot_x).sum().backward()

    sht = sb.RealSHT(nlat, nlon, lmax=lmax, mmax=mmax, grid=grid, precision=precision)
    isht = sb.InverseRealSHT(nlat, nlon, lmax=lmax, mmax=mmax, grid=grid, precision=precision)
    xd = _leaf(x, dev)
    X = torch.ops.sfno_b200.sht_forward(sht._plan(dev).value, xd, lmax, mmax)
    (X * cot_X.to(dev)).sum().backward()
    cd = _leaf(coeffs, dev)
    y = torch.ops.sfno_b200.sht_inverse(isht._plan(dev).value, cd, nlat, nlon)
    (y * cot_x.to(dev)).sum().backward()
    e_f, e_i = rel_l2(xd.grad, xr.grad), rel_l2(cd.grad, cr.grad)
    print(f"SHT gradients {precision} {grid} {nlat}x{nlon}: d/dx forward {e_f:.3e}, d/dcoeffs inverse {e_i:.3e}")
    assert e_f < tol and e_i < tol


def test_sht_module_call_is_differentiable(dev):
    """The drop-in modules (complex in / out) backpropagate too: view_as_complex / view_as_real are torch views."""
    nlat, nlon = 16, 32
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 3, nlat, nlon, generator=g)
    o_sht = oh.RealSHT(nlat, nlon, grid="equiangular").float()
    o_isht = oh.InverseRealSHT(nlat, nlon, grid="equiangular").float()
    xr = _leaf(x)
    o_isht(o_sht(xr)).square().sum().backward()
    sht, isht = sb.RealSHT(nlat, nlon, grid="equiangular").float(), sb.InverseRealSHT(nlat, nlon, grid="equiangular").float()
    xd = _leaf(x, dev)
    isht(sht(xd)).square().sum().backward()
    assert rel_l2(xd.grad, xr.grad) < GRAD_TOL


# ---------------------------------------------------------------------------------------------------------------------------
# contraction, 1x1 convolution, InstanceNorm
# ---------------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("op", ["dhconv", "diagonal"])
@pytest.mark.parametrize("B,Ci,Co,L,M", [(3, 20, 12, 9, 11), (2, 64, 64, 45, 46), (1, 5, 7, 3, 4)])
def test_spectral_contract_gradients(dev, op, B, Ci, Co, L, M):
    g = torch.Generator().manual_seed(B * 100 + Ci)
    x = torch.randn(B, Ci, L, M, 2, generator=g)
    w = torch.randn(*((Ci, Co, L, 2) if op == "dhconv" else (Ci, Co, L, M, 2)), generator=g)
    cot = torch.randn(B, Co, L, M, 2, generator=g)
    xr, wr = _leaf(x.double()), _leaf(w.double())
    fn = dhconv_contract if op == "dhconv" else diagonal_contract
    (torch.view_as_real(fn(torch.view_as_complex(xr), wr)) * cot.double()).sum().backward()
    xd, wd = _leaf(x, dev), _leaf(w, dev)
    y = torch.ops.sfno_b200.spectral_contract(_lib.SFNO_OP[op], xd, wd)
    (y * cot.to(dev)).sum().backward()
    assert rel_l2(xd.grad, xr.grad) < 2e-6 and rel_l2(wd.grad, wr.grad) < 2e-6


@pytest.mark.parametrize("precision,tol", [("fp32", GRAD_TOL), ("tf32", 2e-3), ("bf16", 1.2e-2)])
@pytest.mark.parametrize("op,cin,cout", [("dhconv", 6, 4), ("dhconv", 64, 64), ("diagonal", 5, 5)])
@pytest.mark.parametrize("grid_in,grid_out,nlat,nlon", [("equiangular", "equiangular", 16, 32), ("equiangular", "legendre-gauss", 24, 48),
                                                         ("legendre-gauss", "legendre-gauss", 90, 180)])
def test_fused_spectral_conv_gradients(dev, precision, tol, op, cin, cout, grid_in, grid_out, nlat, nlon):
    """``SpectralConvS2.forward`` (one library call forward, one backward) against autograd through the oracle's transforms
    and contraction (s2convolutions.py:158-193), with and without the resampled residual (``scale_residual``)."""
    lmax, mmax = nlat, nlon // 2 + 1
    B = 3
    g = torch.Generator().manual_seed(nlat + cin)
    x = torch.randn(B, cin, nlat, nlon, generator=g)
    w = torch.randn(*((cin, cout, lmax, 2) if op == "dhconv" else (cin, cout, lmax, mmax, 2)), generator=g) / math.sqrt(cin)
    b = torch.randn(1, cout, 1, 1, generator=g)
    cy, cr = torch.randn(B, cout, nlat, nlon, generator=g), torch.randn(B, cin, nlat, nlon, generator=g)
    scale_residual = grid_in != grid_out
    o_sht = oh.RealSHT(nlat, nlon, lmax=lmax, mmax=mmax, grid=grid_in).double()
    o_isht = oh.InverseRealSHT(nlat, nlon, lmax=lmax, mmax=mmax, grid=grid_out).double()
    xr, wr, br = _leaf(x.double()), _leaf(w.double()), _leaf(b.double())
    X = o_sht(xr)
    Y = (dhconv_contract if op == "dhconv" else diagonal_contract)(X, wr)
    y_ref = o_isht(Y) + br
    loss = (y_ref * cy.double()).sum()
    if scale_residual:
        loss = loss + (o_isht(X) * cr.double()).sum()
    loss.backward()

    sht = sb.RealSHT(nlat, nlon, lmax=lmax, mmax=mmax, grid=grid_in, precision=precision)
    isht = sb.InverseRealSHT(nlat, nlon, lmax=lmax, mmax=mmax, grid=grid_out, precision=precision)
    conv = sb.SpectralConvS2(sht, isht, cin, cout, operator_type=op, bias=True).to(dev)
    with torch.no_grad():
        conv.weight.copy_(w)
        conv.bias.copy_(b)
    assert bool(conv.scale_residual) == scale_residual
    xd = _leaf(x, dev)
    y, res = conv(xd)
    assert rel_l2(y.detach(), y_ref.detach()) < tol
    loss_d = (y * cy.to(dev)).sum()
    if scale_residual:
        loss_d = loss_d + (res * cr.to(dev)).sum()
    else:
        assert res is xd
    loss_d.backward()
    errs = {"x": rel_l2(xd.grad, xr.grad), "weight": rel_l2(conv.weight.grad, wr.grad), "bias": rel_l2(conv.bias.grad, br.grad)}
    print(f"fused spectral conv gradients {precision} {op} {cin}->{cout} {grid_in}->{grid_out} {nlat}x{nlon}: " +
          ", ".join(f"{k} {v:.3e}" for k, v in errs.items()))
    assert max(errs.values()) < tol, errs


@pytest.mark.parametrize("B,cin,cout,H,W,bias,res", [(2, 5, 16, 12, 24, True, False), (3, 36, 40, 18, 36, False, True),
                                                      (2, 130, 34, 180, 360, True, True), (8, 64, 128, 90, 180, True, False)])
def test_conv1x1_gradients(dev, B, cin, cout, H, W, bias, res):
    g = torch.Generator().manual_seed(cin + cout)
    x = torch.randn(B, cin, H, W, generator=g)
    w = torch.randn(cout, cin, 1, 1, generator=g) / math.sqrt(cin)
    b = torch.randn(cout, generator=g) if bias else None
    r = torch.randn(B, cout, H, W, generator=g) if res else None
    cot = torch.randn(B, cout, H, W, generator=g)
    leaves_r = [_leaf(t.double()) if t is not None else None for t in (x, w, b, r)]
    y = torch.nn.functional.conv2d(leaves_r[0], leaves_r[1], leaves_r[2])
    if res:
        y = y + leaves_r[3]
    (y * cot.double()).sum().backward()
    leaves_d = [_leaf(t, dev) if t is not None else None for t in (x, w, b, r)]
    yd = torch.ops.sfno_b200.conv1x1(*leaves_d, 0)
    (yd * cot.to(dev)).sum().backward()
    for name, a, ref in zip(("x", "weight", "bias", "residual"), leaves_d, leaves_r):
        if a is not None:
            assert a.grad.shape == a.shape
            e = rel_l2(a.grad, ref.grad)
            assert e < 5e-6, (name, e)


@pytest.mark.parametrize("precision,tol", [("fp32", 5e-6), ("tf32", 1.5e-3), ("bf16", 8e-3)])
@pytest.mark.parametrize("B,cin,cout,H,W", [(2, 64, 96, 30, 64), (3, 128, 136, 90, 180), (2, 36, 40, 18, 36)])
def test_conv1x1_ex_gradients_on_each_engine(dev, precision, tol, B, cin, cout, H, W):
    """Forward, data gradient (forward op, transposed weight) and weight gradient (split-K GEMM over the pixels with the K
    chunk as a TMA dimension) on each engine; 36 -> 40 channels falls back to the CUDA-core engine for the weight gradient
    (cin % 8 != 0)."""
    g = torch.Generator().manual_seed(8)
    x = torch.randn(B, cin, H, W, generator=g)
    w = torch.randn(cout, cin, 1, 1, generator=g) / math.sqrt(cin)
    b = torch.randn(cout, generator=g)
    cot = torch.randn(B, cout, H, W, generator=g)
    xr, wr, br = _leaf(x.double()), _leaf(w.double()), _leaf(b.double())
    (torch.nn.functional.conv2d(xr, wr, br) * cot.double()).sum().backward()
    xd, wd, bd = _leaf(x, dev), _leaf(w, dev), _leaf(b, dev)
    y = torch.ops.sfno_b200.conv1x1_ex(xd, wd, bd, None, 0, 0.0, 0, 0, _lib.SFNO_PREC[precision])
    (y * cot.to(dev)).sum().backward()
    errs = {"x": rel_l2(xd.grad, xr.grad), "weight": rel_l2(wd.grad, wr.grad), "bias": rel_l2(bd.grad, br.grad)}
    print(f"conv1x1_ex gradients {precision}: " + ", ".join(f"{k} {v:.3e}" for k, v in errs.items()))
    assert max(errs.values()) < tol, errs
    y2 = torch.ops.sfno_b200.conv1x1_ex(xd, wd, bd, None, _lib.SFNO_ACT["gelu"], 0.0, 0, 0, _lib.SFNO_PREC[precision])
    with pytest.raises(NotImplementedError):
        y2.sum().backward()


def test_conv1x1_with_fused_activation_refuses_backward(dev):
    x = torch.randn(1, 4, 8, 16, device=dev, requires_grad=True)
    w = torch.randn(4, 4, device=dev)
    y = torch.ops.sfno_b200.conv1x1(x, w, None, None, _lib.SFNO_ACT["gelu"])
    with pytest.raises(NotImplementedError):
        y.sum().backward()


@pytest.mark.parametrize("affine,with_time", [(True, True), (True, False), (False, True), (False, False)])
def test_instance_norm_gradients(dev, affine, with_time):
    g = torch.Generator().manual_seed(6)
    B, C, H, W = 3, 10, 18, 36
    x = 3.0 + 2.0 * torch.randn(B, C, H, W, generator=g)
    gamma, beta = (torch.randn(C, generator=g), torch.randn(C, generator=g)) if affine else (None, None)
    scale, shift = (0.3 * torch.randn(B, C, generator=g), torch.randn(B, C, generator=g)) if with_time else (None, None)
    cot = torch.randn(B, C, H, W, generator=g)
    ref_leaves = [_leaf(t.double()) if t is not None else None for t in (x, gamma, beta, scale, shift)]
    xr, gr, br, sr, hr = ref_leaves
    y = instance_norm(xr, gr, br) if affine else instance_norm(xr, torch.ones(C, dtype=torch.float64), torch.zeros(C, dtype=torch.float64))
    if with_time:
        y = y * (sr[:, :, None, None] + 1) + hr[:, :, None, None]
    (y * cot.double()).sum().backward()
    dev_leaves = [_leaf(t, dev) if t is not None else None for t in (x, gamma, beta, scale, shift)]
    yd = torch.ops.sfno_b200.instance_norm(*dev_leaves, 1e-6)
    (yd * cot.to(dev)).sum().backward()
    for name, a, ref in zip(("x", "gamma", "beta", "scale", "shift"), dev_leaves, ref_leaves):
        if a is not None:
            e = rel_l2(a.grad, ref.grad)
            assert e < 2e-5, (name, e)


# ---------------------------------------------------------------------------------------------------------------------------
# whole network: loss gradient w.r.t. every parameter and the input
# ---------------------------------------------------------------------------------------------------------------------------
NET_CASES = {
    "dhconv_time_12x24": SFNOConfig(spatial_shape=(12, 24), num_input_channels=3, num_output_channels=2, num_conditional_channels=2,
                                    embed_dim=16, num_layers=3, operator_type="dhconv", with_time_emb=True, min_time=0.0, max_time=6.0),
    "diagonal_lg_16x32": SFNOConfig(spatial_shape=(16, 32), num_input_channels=4, num_output_channels=4, embed_dim=12, num_layers=2,
                                        operator_type="diagonal", data_grid="legendre-gauss", with_time_emb=False),
    "dhconv_nonorm_18x36": SFNOConfig(spatial_shape=(18, 36), num_input_channels=2, num_output_channels=3, embed_dim=8, num_layers=2,
                                      operator_type="dhconv", normalization_layer="none", big_skip=False, pos_embed=False,
                                      with_time_emb=False),
    "dhconv_time_after_rescaled_16x32": SFNOConfig(spatial_shape=(16, 32), num_input_channels=2, num_output_channels=2, embed_dim=8,
                                                   num_layers=2, operator_type="dhconv", with_time_emb=True, time_rescale=True,
                                                   time_scale_shift_before_filter=False, min_time=1.0, max_time=5.0),
}


def _oracle_grads(cfg, sd, x, target, time, condition, loss, dtype=torch.float32):
    o = SFNOOracle(cfg, sd, dtype=dtype)
    o.sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in o.sd.items()}
    xr = _leaf(x.to(dtype))
    out = o._forward(xr, time=time, condition=condition)
    l = (out - target.to(dtype)).abs().mean() if loss == "l1" else (out - target.to(dtype)).square().mean()
    l.backward()
    return float(l.detach()), xr.grad, {k: v.grad for k, v in o.sd.items() if torch.is_tensor(v) and v.requires_grad}, out.detach()


@pytest.mark.parametrize("case", sorted(NET_CASES))
@pytest.mark.parametrize("precision,tol", [("fp32", GRAD_TOL), ("tf32", 1.8e-3), ("bf16", 1.35e-2)])
def test_net_loss_gradients_match_oracle_autograd(dev, case, precision, tol):
    """d loss / d (every parameter, input) of ``get_loss`` vs autograd through the oracle evaluated in fp64.

    fp32 engine: EVERY parameter's gradient is within 1e-4 of the fp64 gradient, relative to its own norm -- or, where the
    gradient cancels to (nearly) zero in exact arithmetic (a bias in front of an InstanceNorm: mlp.fwd.2.bias of every block
    but the last has a true gradient of 1e-20), no further from it than 4x the error of fp32 CPU autograd through the same
    algorithm.  Tensor-core engines (transforms in tf32 / bf16, everything else fp32): the whole gradient vector is within
    the mode's forward bound (tests/test_gpu_parity.py); the worst single parameter is printed, not bounded -- small
    gradients that are differences of large path contributions inherit the absolute error of the large ones."""
    cfg = NET_CASES[case]
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
    l_ref, gx_ref, gp_ref, out_ref = _oracle_grads(cfg, sd, x, target, time, cond, loss, torch.float64)
    _, gx_32, gp_32, _ = _oracle_grads(cfg, sd, x, target, time, cond, loss, torch.float32)

    m = sb.SphericalFourierNeuralOperatorNet(
        num_input_channels=cfg.num_input_channels, num_output_channels=cfg.num_output_channels,
        num_output_channels_raw=cfg.num_output_channels, num_conditional_channels=cfg.num_conditional_channels,
        spatial_shape_in=cfg.spatial_shape, spatial_shape_out=cfg.spatial_shape, precision=precision, loss_function=loss,
        **cfg.model_kwargs())
    m.load_state_dict(sd, strict=True)
    if cfg.with_time_emb:
        m.set_min_max_time(cfg.min_time, cfg.max_time)
    m = m.to(dev).train()
    xd = _leaf(x, dev)
    kwargs = {}
    if time is not None:
        kwargs["time"] = time.to(dev)
    loss_dict, preds = m.get_loss(xd, target.to(dev), condition=None if cond is None else cond.to(dev), return_predictions=True, **kwargs)
    assert rel_l2(preds.detach(), out_ref) < tol
    loss_dict["loss"].backward()
    assert abs(float(loss_dict["loss"]) - l_ref) <= (1e-5 if precision == "fp32" else tol) * abs(l_ref)
    e_in = rel_l2(xd.grad, gx_ref)
    worst = ("input", e_in)
    num = den = 0.0
    for name, p in m.named_parameters():
        assert p.grad is not None, f"{name} has no gradient"
        assert p.grad.shape == p.shape
        ref = gp_ref[name].reshape(p.shape)
        err = float((p.grad.cpu().double() - ref).norm())
        err32 = float((gp_32[name].reshape(p.shape).double() - ref).norm())
        n = float(ref.norm())
        num, den = num + err * err, den + n * n
        if precision == "fp32":
            assert err <= tol * n + 4.0 * err32, (name, err, n, err32)
        if n > 0 and err / n > worst[1] and n > 1e-12:
            worst = (name, err / n)
    e_all = math.sqrt(num / den)
    print(f"{case} {precision}: loss {float(loss_dict['loss']):.6f} (oracle {l_ref:.6f}); gradient rel-L2: input {e_in:.3e}, "
          f"all parameters {e_all:.3e}, worst single parameter {worst[0]} {worst[1]:.3e}")
    assert e_in < tol and e_all < tol


def test_training_step_reduces_loss_and_inference_path_sees_new_weights(dev):
    """A few optimizer steps through the library's backward lower the loss, and the fused inference executor
    (``sfno_net_forward``) picks the updated parameters up (version-counter check of ``sync_parameters``)."""
    cfg = NET_CASES["dhconv_time_12x24"]
    sd = random_state_dict(cfg, seed=5, spectral_gain=4.0)
    m = sb.SphericalFourierNeuralOperatorNet(
        num_input_channels=cfg.num_input_channels, num_output_channels=cfg.num_output_channels,
        num_output_channels_raw=cfg.num_output_channels, num_conditional_channels=cfg.num_conditional_channels,
        spatial_shape_in=cfg.spatial_shape, spatial_shape_out=cfg.spatial_shape, precision="fp32", loss_function="mse",
        **cfg.model_kwargs())
    m.load_state_dict(sd, strict=True)
    m.set_min_max_time(cfg.min_time, cfg.max_time)
    m = m.to(dev).train()
    g = torch.Generator().manual_seed(1)
    x = torch.randn(4, cfg.num_input_channels, 12, 24, generator=g).to(dev)
    cond = torch.randn(4, cfg.num_conditional_channels, 12, 24, generator=g).to(dev)
    target = 0.5 * x[:, :cfg.num_output_channels] - 0.25 * cond      # learnable through the big skip
    time = torch.tensor([1.0, 2.0, 3.0, 4.0], device=dev)
    opt = torch.optim.Adam(m.parameters(), lr=1e-2)
    losses = []
    for _ in range(30):
        opt.zero_grad()
        l = m.get_loss(x, target, condition=cond, time=time)["loss"]
        l.backward()
        opt.step()
        losses.append(float(l))
    assert losses[-1] < 0.5 * losses[0], losses
    m.eval()
    with torch.no_grad():
        fused = m(x, time=time, condition=cond)
    with torch.enable_grad():
        piecewise = m(x, time=time, condition=cond)      # parameters require grad -> differentiable path
    assert rel_l2(fused, piecewise.detach()) < 1e-5
