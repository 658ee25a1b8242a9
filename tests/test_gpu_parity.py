"""GPU parity tests (run with ``-m gpu`` on the B200 box).  Everything goes through the C ABI
(``libsfno_b200.so`` via the drop-in modules) and is compared with the CPU oracle / the reference-generated
golden fixtures.  Tolerances: fp32 mode <= 1e-4 relative L2 (north_star); bf16 mode <= BF16_BOUND."""
import math

import pytest
import torch

from conftest import golden_cases
from oracle import harmonics as oh
from oracle.sfno_oracle import (ACE_FORECASTER, SFNOConfig, SFNOOracle, dhconv_contract, diagonal_contract, instance_norm,
                                perturb_affine_and_biases, random_state_dict, rel_l2)

import spherical_dyffusion_b200 as sb
from spherical_dyffusion_b200 import _lib
from spherical_dyffusion_b200._util import stream_ptr, workspace

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-4      # north_star: <= 1e-4 relative L2 in fp32
BF16_BOUND = 1.35e-2  # stated bf16 bound on the end-to-end forward: 1.5 x the largest measured value (8.93e-3, profiles/r02_b_pytest_gpu.log)
BF16_OP_BOUND = 5.7e-3  # single transform / op in bf16: 1.5 x measured (3.75e-3)
TF32_BOUND = 1.8e-3   # tf32 mode (fp32 storage, kind::tf32 tensor-core MMA), end-to-end forward: 1.5 x measured (1.21e-3; ACE 1.09e-3)
TF32_OP_BOUND = 8e-4  # single transform in tf32: 1.5 x measured (5.25e-4)


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def module_from_cfg(cfg: SFNOConfig, sd, dev, precision="fp32"):
    m = sb.SphericalFourierNeuralOperatorNet(
        num_input_channels=cfg.num_input_channels, num_output_channels=cfg.num_output_channels,
        num_output_channels_raw=cfg.num_output_channels, num_conditional_channels=cfg.num_conditional_channels,
        spatial_shape_in=cfg.spatial_shape, spatial_shape_out=cfg.spatial_shape, precision=precision, **cfg.model_kwargs())
    m.load_state_dict(sd, strict=True)
    if cfg.with_time_emb:
        m.set_min_max_time(cfg.min_time, cfg.max_time)
    return m.to(dev).eval()


# ---------------------------------------------------------------------------------------------------------
# SHT pair
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("grid", ["legendre-gauss", "equiangular"])
@pytest.mark.parametrize("nlat,nlon,lead", [(12, 24, (2, 3)), (18, 36, (5,)), (33, 64, (1, 7)), (180, 360, (1, 8)), (180, 360, (37,))])
def test_sht_forward_inverse_match_oracle(dev, grid, nlat, nlon, lead):
    lmax, mmax = nlat, nlon // 2 + 1
    g = torch.Generator().manual_seed(nlat * 7 + len(lead))
    x = torch.randn(*lead, nlat, nlon, generator=g)
    o_sht = oh.RealSHT(nlat, nlon, lmax=lmax, mmax=mmax, grid=grid).float()
    o_isht = oh.InverseRealSHT(nlat, nlon, lmax=lmax, mmax=mmax, grid=grid).float()
    X_ref = o_sht(x)
    sht = sb.RealSHT(nlat, nlon, lmax=lmax, mmax=mmax, grid=grid).float()
    isht = sb.InverseRealSHT(nlat, nlon, lmax=lmax, mmax=mmax, grid=grid).float()
    X = sht(x.to(dev))
    assert X.shape == X_ref.shape and X.dtype == torch.complex64
    assert rel_l2(X, X_ref) < 1e-5
    # inverse on the oracle's coefficients, with non-zero Im at m = 0 / Nyquist to exercise the C2R rule
    Xp = X_ref.clone()
    Xp[..., 0] += 0.5j
    Xp[..., -1] += 0.25j
    xr_ref = o_isht(Xp)
    xr = isht(Xp.to(dev))
    assert xr.shape == xr_ref.shape
    assert rel_l2(xr, xr_ref) < 1e-5


def test_sht_round_trip_properties_full_size(dev):
    """Size-independent properties at 180x360: linearity, LG round trip is a projection."""
    nlat, nlon = 180, 360
    sht = sb.RealSHT(nlat, nlon, lmax=180, mmax=181, grid="legendre-gauss")
    isht = sb.InverseRealSHT(nlat, nlon, lmax=180, mmax=181, grid="legendre-gauss")
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 64, nlat, nlon, generator=g).to(dev)
    y = torch.randn(1, 64, nlat, nlon, generator=g).to(dev)
    assert rel_l2(sht(2.0 * x - 3.0 * y), 2.0 * sht(x) - 3.0 * sht(y)) < 1e-5
    xb = isht(sht(x))
    assert rel_l2(isht(sht(xb)), xb) < 1e-5
    # empty leading dimension
    assert sht(torch.zeros(0, 3, nlat, nlon, device=dev)).shape == (0, 3, 180, 181)


@pytest.mark.parametrize("grid", ["legendre-gauss", "equiangular"])
def test_sht_bf16_bound(dev, grid):
    nlat, nlon = 180, 360
    x = torch.randn(1, 16, nlat, nlon, generator=torch.Generator().manual_seed(3))
    X_ref = oh.RealSHT(nlat, nlon, lmax=180, mmax=181, grid=grid).float()(x)
    X = sb.RealSHT(nlat, nlon, lmax=180, mmax=181, grid=grid, precision="bf16")(x.to(dev))
    e_fwd = rel_l2(X, X_ref)
    xr_ref = oh.InverseRealSHT(nlat, nlon, lmax=180, mmax=181, grid=grid).float()(X_ref)
    xr = sb.InverseRealSHT(nlat, nlon, lmax=180, mmax=181, grid=grid, precision="bf16")(X_ref.to(dev))
    e_inv = rel_l2(xr, xr_ref)
    print(f"bf16 SHT rel-L2: forward {e_fwd:.3e} inverse {e_inv:.3e} ({grid})")
    assert e_fwd < BF16_OP_BOUND and e_inv < BF16_OP_BOUND


# ---------------------------------------------------------------------------------------------------------
# op-level: contraction, InstanceNorm, conv1x1
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("op", ["dhconv", "diagonal"])
def test_spectral_contract_matches_oracle(dev, op):
    g = torch.Generator().manual_seed(5)
    B, Ci, Co, L, M = 3, 20, 12, 9, 11
    x = torch.randn(B, Ci, L, M, 2, generator=g)
    w = torch.randn(*((Ci, Co, L, 2) if op == "dhconv" else (Ci, Co, L, M, 2)), generator=g)
    xc = torch.view_as_complex(x)
    ref = dhconv_contract(xc, w) if op == "dhconv" else diagonal_contract(xc, w)
    out = torch.empty(B, Co, L, M, 2, device=dev)
    xd, wd = x.to(dev), w.to(dev)
    _lib.check(sb.lib().sfno_spectral_contract(_lib.SFNO_OP[op], xd.data_ptr(), wd.data_ptr(), out.data_ptr(), B, Ci, Co, L, M,
                                               stream_ptr(dev)))
    assert rel_l2(torch.view_as_complex(out), ref) < 1e-6


@pytest.mark.parametrize("with_time", [False, True])
def test_instance_norm_matches_oracle(dev, with_time):
    g = torch.Generator().manual_seed(6)
    B, C, H, W = 3, 10, 18, 36
    x = 3.0 + 2.0 * torch.randn(B, C, H, W, generator=g)
    gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g)
    ref = instance_norm(x, gamma, beta)
    scale = shift = None
    if with_time:
        scale, shift = 0.3 * torch.randn(B, C, generator=g), torch.randn(B, C, generator=g)
        ref = ref * (scale[:, :, None, None] + 1) + shift[:, :, None, None]
    L = sb.lib()
    xd, y = x.to(dev), torch.empty(B, C, H, W, device=dev)
    gd, bd = gamma.to(dev), beta.to(dev)
    sd_, sh_ = (scale.to(dev), shift.to(dev)) if with_time else (None, None)
    ws = workspace(dev, L.sfno_instance_norm_workspace_bytes(B, C), "t")
    _lib.check(L.sfno_instance_norm(xd.data_ptr(), y.data_ptr(), gd.data_ptr(), bd.data_ptr(),
                                    sd_.data_ptr() if with_time else None, sh_.data_ptr() if with_time else None,
                                    B, C, H * W, 1e-6, ws.data_ptr(), ws.numel(), stream_ptr(dev)))
    assert rel_l2(y, ref) < 1e-5
    y_op = torch.ops.sfno_b200.instance_norm(xd, gd, bd, sd_, sh_, 1e-6)   # the same call through the custom-op layer
    assert torch.equal(y_op, y)


@pytest.mark.parametrize("cin,cout,hw,act", [(5, 16, 288, "gelu"), (36, 256, 1000, "none"), (130, 34, 64800, "gelu")])
def test_conv1x1_matches_torch(dev, cin, cout, hw, act):
    g = torch.Generator().manual_seed(cin)
    B = 2
    x = torch.randn(B, cin, hw, generator=g)
    w = torch.randn(cout, cin, generator=g) / math.sqrt(cin)
    b = torch.randn(cout, generator=g)
    r = torch.randn(B, cout, hw, generator=g)
    ref = torch.einsum("oc,bcp->bop", w.double(), x.double()) + b.double()[None, :, None]
    if act == "gelu":
        ref = torch.nn.functional.gelu(ref)
    ref = ref + r.double()
    xd, wd, bd, rd = x.to(dev), w.to(dev), b.to(dev), r.to(dev)
    y = torch.empty(B, cout, hw, device=dev)
    _lib.check(sb.lib().sfno_conv1x1(xd.data_ptr(), wd.data_ptr(), bd.data_ptr(), rd.data_ptr(), y.data_ptr(), B, cin, cout, hw,
                                     _lib.SFNO_ACT[act], stream_ptr(dev)))
    assert rel_l2(y, ref) < 2e-6
    y_op = torch.ops.sfno_b200.conv1x1(xd.reshape(B, cin, 1, hw), wd, bd, rd.reshape(B, cout, 1, hw), _lib.SFNO_ACT[act])
    assert torch.equal(y_op.reshape(B, cout, hw), y)   # the same call through the custom-op layer


def test_abi_error_statuses(dev):
    L = sb.lib()
    with pytest.raises(_lib.SfnoLibraryError, match="invalid argument"):
        _lib.check(L.sfno_conv1x1(None, None, None, None, None, 1, 1, 1, 1, 0, None))
    sht = sb.RealSHT(12, 24, grid="equiangular")
    plan = sht._plan(dev)
    x = torch.zeros(1, 12, 24, device=dev)
    out = torch.zeros(1, 12, 13, 2, device=dev)
    ws = torch.zeros(16, dtype=torch.uint8, device=dev)
    with pytest.raises(_lib.SfnoLibraryError, match="workspace too small"):
        _lib.check(L.sfno_sht_forward(plan, x.data_ptr(), out.data_ptr(), 1, ws.data_ptr(), ws.numel(), stream_ptr(dev)))
    with pytest.raises(AssertionError):  # torch_harmonics asserts on shape (SURVEY 8b error conventions)
        sht(torch.zeros(1, 13, 24, device=dev))


# ---------------------------------------------------------------------------------------------------------
# whole network vs the reference-generated goldens
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", golden_cases())
def test_net_matches_reference_golden_fp32(dev, case, load_golden):
    fx = load_golden(case)
    cfg = SFNOConfig(**fx["cfg"])
    m = module_from_cfg(cfg, fx["state_dict"], dev)
    cond = fx["condition"].to(dev) if fx["condition"] is not None else None
    time = fx["time"].to(dev) if fx["time"] is not None else None
    with torch.inference_mode():
        out, t_repr = m(fx["inputs"].to(dev), time=time, condition=cond, return_time_emb=True)
        errs = {}
        for i in range(cfg.num_layers):
            act = m.debug_activation(i, fx["inputs"].to(dev), time=time, condition=cond)
            errs[i] = rel_l2(act, fx["taps"][f"blocks.{i}.out"])
    e = rel_l2(out, fx["output"])
    print(f"{case}: out rel-L2 {e:.3e}; per-block {', '.join(f'{v:.2e}' for v in errs.values())}")
    if fx["t_repr"] is not None:
        assert rel_l2(t_repr, fx["t_repr"]) < 1e-5
    for i, v in errs.items():
        assert v < FP32_TOL, f"block {i}: {v}"
    assert out.shape == fx["output"].shape
    assert e < FP32_TOL


@pytest.mark.parametrize("case", ["sfno_dhconv_12x24", "sfno_dhconv_18x36_lg", "sfno_dhconv_16x32_variants",
                                  "sfno_dhconv_24x48_interp", "sfno_dhconv_24x48_lg_b1"])
def test_net_golden_bf16_bound(dev, case, load_golden):
    fx = load_golden(case)
    cfg = SFNOConfig(**fx["cfg"])
    m = module_from_cfg(cfg, fx["state_dict"], dev, precision="bf16")
    cond = fx["condition"].to(dev) if fx["condition"] is not None else None
    time = fx["time"].to(dev) if fx["time"] is not None else None
    with torch.inference_mode():
        out = m(fx["inputs"].to(dev), time=time, condition=cond)
    e = rel_l2(out, fx["output"])
    print(f"{case}: bf16 out rel-L2 {e:.3e}")
    assert e < BF16_BOUND


def test_net_parameter_resync_after_inplace_update(dev, load_golden):
    """EMA swaps / load_state_dict mutate parameters in place (ema.py:54-91): packed copies must follow."""
    fx = load_golden("sfno_dhconv_12x24")
    cfg = SFNOConfig(**fx["cfg"])
    m = module_from_cfg(cfg, fx["state_dict"], dev)
    x, c, t = fx["inputs"].to(dev), fx["condition"].to(dev), fx["time"].to(dev)
    with torch.inference_mode():
        y0 = m(x, time=t, condition=c).clone()
    with torch.no_grad():
        for p in m.parameters():
            p.mul_(1.5)
    sd2 = {k: v * 1.5 for k, v in fx["state_dict"].items()}
    ref = SFNOOracle(cfg, sd2)(fx["inputs"], time=fx["time"], condition=fx["condition"])
    with torch.inference_mode():
        y1 = m(x, time=t, condition=c)
    assert rel_l2(y1, ref) < FP32_TOL
    assert rel_l2(y1, y0) > 1e-2
    m.load_state_dict(fx["state_dict"])
    with torch.inference_mode():
        assert rel_l2(m(x, time=t, condition=c), fx["output"]) < FP32_TOL
    # the reference's EMA swap writes through .data (ema.py:54-91 copy_to / restore): autograd's version counter does
    # not move, the device-side fingerprint (param_check="checksum", the default) must catch it
    versions = [p._version for p in m.parameters()]
    for p in m.parameters():
        p.data.copy_(p.data * 1.5)
    assert versions == [p._version for p in m.parameters()]
    with torch.inference_mode():
        assert rel_l2(m(x, time=t, condition=c), ref) < FP32_TOL
    # param_check="version" trusts the counter: stale until invalidate_parameters() is called
    m.param_check = "version"
    for p in m.parameters():
        p.data.copy_(p.data / 1.5)
    with torch.inference_mode():
        assert rel_l2(m(x, time=t, condition=c), ref) < FP32_TOL          # still the x1.5 weights
        m.invalidate_parameters()
        assert rel_l2(m(x, time=t, condition=c), fx["output"]) < FP32_TOL


def test_net_time_assert_and_static_condition(dev, load_golden):
    fx = load_golden("sfno_dhconv_12x24")
    cfg = SFNOConfig(**fx["cfg"])
    m = module_from_cfg(cfg, fx["state_dict"], dev)
    x, c = fx["inputs"].to(dev), fx["condition"].to(dev)
    with torch.inference_mode():
        with pytest.raises(AssertionError):  # sfnonet.py:780-782
            m(x, time=torch.tensor([1.0, 7.0], device=dev), condition=c)
        y = m(x, time=fx["time"].to(dev), static_condition=c)  # _base_model.py:175-177
    assert rel_l2(y, fx["output"]) < FP32_TOL


def test_net_inference_dropout_statistics(dev):
    """Dropout / DropPath are live at inference inside the interpolator (dyffusion.py:226-235): masks change per
    call, are reproducible for a fixed Philox key, and the ensemble mean approaches the deterministic output."""
    cfg = SFNOConfig(num_input_channels=4, num_output_channels=4, num_conditional_channels=0, spatial_shape=(16, 32),
                     embed_dim=32, num_layers=3, dropout_mlp=0.1, drop_path_rate=0.1, with_time_emb=False)
    sd = perturb_affine_and_biases(random_state_dict(cfg, seed=11))
    m = module_from_cfg(cfg, sd, dev)
    x = torch.randn(64, 4, 16, 32, generator=torch.Generator().manual_seed(1)).to(dev)
    with torch.inference_mode():
        y_det = m(x)
        with m.inference_dropout_scope(condition=True):
            m.seed_dropout(0, 100 * 4096)
            y1 = m(x)
            y2 = m(x)
            m.seed_dropout(0, 100 * 4096)
            y1b = m(x)
        y_det2 = m(x)
    assert torch.equal(y_det, y_det2)
    assert torch.equal(y1, y1b)
    assert rel_l2(y1, y2) > 1e-3
    assert rel_l2(y1, y_det) > 1e-3
    ref = SFNOOracle(cfg, sd)(x.cpu())
    assert rel_l2(y_det, ref) < FP32_TOL


# ---------------------------------------------------------------------------------------------------------
# ACE-sized forward vs the travelling oracle
# ---------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def ace_case():
    cfg = SFNOConfig(**ACE_FORECASTER)
    sd = perturb_affine_and_biases(random_state_dict(cfg, seed=0, spectral_gain=256.0))
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 34, 180, 360, generator=g)
    c = torch.randn(1, 2, 180, 360, generator=g)
    t = torch.tensor([3.0])
    torch.set_num_threads(max(1, torch.get_num_threads()))
    ref = SFNOOracle(cfg, sd)(x, time=t, condition=c)
    return cfg, sd, x, c, t, ref


def test_ace_forward_fp32(dev, ace_case):
    cfg, sd, x, c, t, ref = ace_case
    m = module_from_cfg(cfg, sd, dev)
    assert m.num_params == sum(v.numel() for v in sd.values())
    with torch.inference_mode():
        y = m(x.to(dev), time=t.to(dev), condition=c.to(dev))
    e = rel_l2(y, ref)
    print(f"ACE-sized forward fp32 rel-L2 vs oracle: {e:.3e}")
    assert e < FP32_TOL
    # batch independence: rows of a batch are independent samples (InstanceNorm is per sample)
    with torch.inference_mode():
        xb = torch.cat((x, 0.5 * x.flip(-1)), 0).to(dev)
        cb = torch.cat((c, c), 0).to(dev)
        yb = m(xb, time=torch.tensor([3.0, 1.0], device=dev), condition=cb)
    assert rel_l2(yb[:1], y) < 1e-5


def test_ace_forward_bf16_bound(dev, ace_case):
    cfg, sd, x, c, t, ref = ace_case
    m = module_from_cfg(cfg, sd, dev, precision="bf16")
    with torch.inference_mode():
        y = m(x.to(dev), time=t.to(dev), condition=c.to(dev))
    e = rel_l2(y, ref)
    print(f"ACE-sized forward bf16 rel-L2 vs oracle: {e:.3e}")
    assert e < BF16_BOUND


# ---------------------------------------------------------------------------------------------------------
# ensemble statistics kernels (metrics.py:166-175,199-246)
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("E", [2, 7, 25])
@pytest.mark.parametrize("n", [34 * 60 * 12, 1003])
def test_ensemble_statistics(dev, E, n):
    """Shifted moments of two member shards about the common pivot (the two-rank path on one GPU), the fused one-pass
    statistics kernel, and both at a large offset (|mean| >> spread: surface pressure) and an n % 4 != 0 tail."""
    from spherical_dyffusion_b200.ensemble import CudaEnsembleOps

    g = torch.Generator().manual_seed(E)
    ops = CudaEnsembleOps()
    for offset, var_tol in ((1.0, 1e-5), (1.0e5, 2e-3)):
        mem = torch.randn(E, n, generator=g) * 2 + offset
        truth = torch.randn(n, generator=g) + offset
        md, td = mem.to(dev), truth.to(dev)
        ref_var = mem.double().var(0).float()
        a, b = md[: E // 2].contiguous(), md[E // 2:].contiguous()
        s_glob = ops.local_sum(a) + ops.local_sum(b)
        mom = ops.shifted_moments(a, s_glob, E) + ops.shifted_moments(b, s_glob, E)
        mean, var = ops.finalize(s_glob, mom, E)
        assert rel_l2(mean, mem.double().mean(0).float()) < 1e-6
        assert rel_l2(var, ref_var) < var_tol
        mean2, var2, crps = ops.stats(md, td)
        assert rel_l2(mean2, mem.double().mean(0).float()) < 1e-6
        assert rel_l2(var2, ref_var) < var_tol
        skill = (mem - truth).abs().mean(0)
        spread = (mem[None] - mem[:, None]).abs().sum((0, 1)) / (E * (E - 1))
        assert rel_l2(crps, skill - 0.5 * spread) < (1e-5 if offset == 1.0 else 2e-2)
        # the raw-moment form this replaces loses everything at the large offset
        raw = ((mem * mem).sum(0) - E * mem.mean(0) ** 2) / (E - 1)
        if offset > 1.0:
            assert rel_l2(raw, ref_var) > 10 * rel_l2(var, ref_var)


def test_cold_update_kernel(dev):
    """x_s + (x_next - x_cur) of dyffusion.py:519 in one launch (vector path, scalar tail, unaligned views)."""
    g = torch.Generator().manual_seed(3)
    for shape in ((2, 34, 12, 24), (1, 3, 5, 7)):
        a, b, c = (torch.randn(*shape, generator=g).to(dev) for _ in range(3))
        out = torch.ops.sfno_b200.cold_update(a, b, c)
        assert torch.equal(out, a + (b - c))
    a, b, c = (torch.randn(1001, generator=g).to(dev)[1:] for _ in range(3))
    assert torch.equal(torch.ops.sfno_b200.cold_update(a, b, c), a + (b - c))


@pytest.mark.parametrize("case", ["e2", "e5_ties", "e8", "e25"])
def test_ensemble_statistics_match_reference_metric_functions(dev, case):
    """The CUDA statistics kernels against numbers produced by the reference's own metrics.py
    (tests/golden/make_golden_metrics.py): mean, spread, RMSE, spread-skill ratio, fair CRPS (reduced and per point)."""
    import os

    from spherical_dyffusion_b200.ensemble import CudaEnsembleOps, EnsembleStatistics, area_weights

    fx = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ensemble_metrics.pt"),
                    map_location="cpu", weights_only=False)[case]
    E, ref = fx["spec"]["E"], fx["ref"]
    members, truth = fx["members"].to(dev), fx["truth"].to(dev)
    weights = area_weights(fx["lats"], members.shape[-1])
    assert torch.equal(weights, ref["weights"])
    out = EnsembleStatistics(E).step(members, truth=truth, weights=weights.to(dev))
    assert rel_l2(out["mean"], ref["mean"]) < 1e-6
    for k in ("spread", "rmse", "ssr", "crps"):
        assert rel_l2(out[k], ref[k]) < 2e-5, k
    pointwise = CudaEnsembleOps().crps(members.reshape(E, -1).contiguous(), truth.reshape(-1).contiguous())
    assert rel_l2(pointwise.reshape(truth.shape), ref["crps_pointwise"]) < 1e-5


def test_ensemble_statistics_module_and_rollout(dev):
    """EnsembleStatistics with the CUDA kernels (single rank) against the metric definitions, and a tiny rollout."""
    from spherical_dyffusion_b200.dyffusion import DYffusion
    from spherical_dyffusion_b200.ensemble import EnsembleStatistics, area_weights
    from spherical_dyffusion_b200.rollout import EnsembleRollout

    g = torch.Generator().manual_seed(2)
    E, C, H, W = 6, 4, 16, 32
    members = (torch.randn(E, C, H, W, generator=g) * 2 + 1).to(dev)
    truth = torch.randn(C, H, W, generator=g).to(dev)
    weights = area_weights(torch.linspace(-84, 84, H), W).to(dev)
    out = EnsembleStatistics(E).step(members, truth=truth, weights=weights)
    wm = lambda x: (x * weights).sum((-2, -1)) / weights.expand(x.shape).sum((-2, -1))
    mean = members.mean(0)
    assert rel_l2(out["spread"], torch.sqrt(wm(members.var(dim=0)))) < 1e-5
    assert rel_l2(out["rmse"], torch.sqrt(wm((mean - truth) ** 2))) < 1e-5
    skill = (members - truth).abs().mean(0)
    sp = (members[None] - members[:, None]).abs().sum((0, 1)) / (E * (E - 1))
    assert rel_l2(out["crps"], wm(skill - 0.5 * sp)) < 1e-5

    hcfg = dict(spatial_shape=(H, W), embed_dim=16, num_layers=2)
    fcfg = SFNOConfig(num_input_channels=C, num_output_channels=C, num_conditional_channels=2, min_time=0.0, max_time=2.0, **hcfg)
    icfg = SFNOConfig(num_input_channels=2 * C, num_output_channels=C, num_conditional_channels=2, min_time=1.0, max_time=2.0,
                      dropout_mlp=0.1, drop_path_rate=0.1, **hcfg)
    fore = module_from_cfg(fcfg, perturb_affine_and_biases(random_state_dict(fcfg, seed=5)), dev)
    ipol = module_from_cfg(icfg, perturb_affine_and_biases(random_state_dict(icfg, seed=6)), dev)
    dy = DYffusion(fore, ipol, timesteps=3)
    stats = EnsembleStatistics(5)
    ro = EnsembleRollout(dy, stats, forcing_fn=lambda s, n, d: torch.zeros(n, 2, H, W, device=d) + 0.01 * s,
                         truth_fn=lambda s, d: torch.zeros(C, H, W, device=d), weights=weights)
    hist = ro.run(torch.randn(C, H, W, generator=g).to(dev), n_steps=7)
    assert len(hist["crps"]) == 7 and len(hist["spread"]) == 7
    assert all(torch.isfinite(v).all() for v in hist["crps"])
    assert float(hist["spread"][-1].mean()) > 0.0  # members diverge through the interpolator's dropout stream


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_forward_does_not_depend_on_workspace_contents(dev, precision, load_golden):
    """The scratch arena is shared between nets and never cleared: poisoning it with NaN bit patterns before a
    forward must not change the result (every scratch entry that is read is written in the same forward)."""
    from spherical_dyffusion_b200 import _util

    cfg = SFNOConfig(num_input_channels=6, num_output_channels=6, num_conditional_channels=2, spatial_shape=(96, 192),
                     embed_dim=64, num_layers=3)
    sd = perturb_affine_and_biases(random_state_dict(cfg, seed=3, spectral_gain=64.0))
    m = module_from_cfg(cfg, sd, dev, precision)
    g = torch.Generator().manual_seed(4)
    x = torch.randn(3, 6, 96, 192, generator=g).to(dev)
    c = torch.randn(3, 2, 96, 192, generator=g).to(dev)
    t = torch.tensor([0.0, 2.0, 5.0], device=dev)
    with torch.inference_mode():
        y0 = m(x, time=t, condition=c).clone()
        for buf in _util._workspaces.values():
            buf.fill_(0xFF)  # 0xFFFF = bf16 NaN, 0xFFFFFFFF = fp32 NaN
        y1 = m(x, time=t, condition=c)
    assert torch.isfinite(y1).all()
    assert rel_l2(y1, y0) < 1e-6
    ref = SFNOOracle(cfg, sd)(x.cpu(), time=t.cpu(), condition=c.cpu())
    assert rel_l2(y1, ref) < (FP32_TOL if precision == "fp32" else BF16_BOUND)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_forward_is_cuda_graph_capturable(dev, precision):
    """The forward enqueues kernels only (no synchronisation, no allocation inside the library, tensor maps passed by
    value), so a user can capture it in a CUDA graph; replays with new inputs in the static buffers reproduce the
    eager results."""
    cfg = SFNOConfig(num_input_channels=6, num_output_channels=6, num_conditional_channels=2, spatial_shape=(32, 64),
                     embed_dim=32, num_layers=2)
    sd = perturb_affine_and_biases(random_state_dict(cfg, seed=5, spectral_gain=16.0))
    m = module_from_cfg(cfg, sd, dev, precision)
    g = torch.Generator().manual_seed(6)
    xs = [torch.randn(2, 6, 32, 64, generator=g).to(dev) for _ in range(2)]
    cs = [torch.randn(2, 2, 32, 64, generator=g).to(dev) for _ in range(2)]
    ts = [torch.tensor([0.0, 3.0], device=dev), torch.tensor([5.0, 1.0], device=dev)]
    with torch.inference_mode():
        eager = [m(x, time=t, condition=c).clone() for x, t, c in zip(xs, ts, cs)]
        sx, sc, st = xs[0].clone(), cs[0].clone(), ts[0].clone()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            m(sx, time=st, condition=sc)  # warm-up on the capture stream (weights packed, workspace allocated)
        torch.cuda.current_stream(dev).wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            sy = m(sx, time=st, condition=sc)
        for x, t, c, ref in zip(xs, ts, cs, eager):
            sx.copy_(x); sc.copy_(c); st.copy_(t)
            graph.replay()
            torch.cuda.synchronize(dev)
            assert rel_l2(sy, ref) < 1e-6


def test_net_embed512_matches_oracle(dev):
    """embed 512 / hidden 1024 (the scaled configuration's widths) on a small grid, both precisions."""
    cfg = SFNOConfig(num_input_channels=4, num_output_channels=4, num_conditional_channels=2, spatial_shape=(16, 32),
                     embed_dim=512, num_layers=2)
    sd = perturb_affine_and_biases(random_state_dict(cfg, seed=7, spectral_gain=256.0))
    g = torch.Generator().manual_seed(8)
    x = torch.randn(2, 4, 16, 32, generator=g)
    c = torch.randn(2, 2, 16, 32, generator=g)
    t = torch.tensor([1.0, 4.0])
    ref = SFNOOracle(cfg, sd)(x, time=t, condition=c)
    for precision, tol in (("fp32", FP32_TOL), ("bf16", BF16_BOUND)):
        m = module_from_cfg(cfg, sd, dev, precision)
        with torch.inference_mode():
            y = m(x.to(dev), time=t.to(dev), condition=c.to(dev))
        assert rel_l2(y, ref) < tol, precision


# ---------------------------------------------------------------------------------------------------------
# tf32 mode: fp32 storage, tcgen05 kind::tf32 (VERDICT r1 "missing" 1: an fp32-grade tensor-core mode)
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("grid", ["legendre-gauss", "equiangular"])
def test_sht_tf32_bound(dev, grid):
    nlat, nlon = 180, 360
    x = torch.randn(1, 16, nlat, nlon, generator=torch.Generator().manual_seed(3))
    X_ref = oh.RealSHT(nlat, nlon, lmax=180, mmax=181, grid=grid).float()(x)
    X = sb.RealSHT(nlat, nlon, lmax=180, mmax=181, grid=grid, precision="tf32")(x.to(dev))
    e_fwd = rel_l2(X, X_ref)
    xr_ref = oh.InverseRealSHT(nlat, nlon, lmax=180, mmax=181, grid=grid).float()(X_ref)
    xr = sb.InverseRealSHT(nlat, nlon, lmax=180, mmax=181, grid=grid, precision="tf32")(X_ref.to(dev))
    e_inv = rel_l2(xr, xr_ref)
    print(f"tf32 SHT rel-L2: forward {e_fwd:.3e} inverse {e_inv:.3e} ({grid})")
    assert e_fwd < TF32_OP_BOUND and e_inv < TF32_OP_BOUND


@pytest.mark.parametrize("case", ["sfno_dhconv_12x24", "sfno_dhconv_18x36_lg", "sfno_dhconv_16x32_variants", "sfno_dhconv_24x48_interp"])
def test_net_forward_tf32_matches_golden(dev, load_golden, case):
    fx = load_golden(case)
    cfg = SFNOConfig(**fx["cfg"])
    m = module_from_cfg(cfg, fx["state_dict"], dev, "tf32")
    cond = fx["condition"].to(dev) if fx["condition"] is not None else None
    time = fx["time"].to(dev) if fx["time"] is not None else None
    with torch.inference_mode():
        y = m(fx["inputs"].to(dev), time=time, condition=cond)
    e = rel_l2(y, fx["output"])
    print(f"{case}: tf32 out rel-L2 {e:.3e}")
    assert e < TF32_BOUND


def test_ace_forward_tf32(dev, ace_case):
    """ACE-sized forward in tf32 mode against the fp32 oracle, next to an EMULATION of what the reference computes under
    torch_matmul_precision "high" (config_utils.py:310-313): the oracle with matmul operands rounded to TF32."""
    cfg, sd, x, c, t, ref = ace_case
    m = module_from_cfg(cfg, sd, dev, "tf32")
    with torch.inference_mode():
        y = m(x.to(dev), time=t.to(dev), condition=c.to(dev))
    e = rel_l2(y, ref)
    emu = rel_l2(SFNOOracle(cfg, sd, tf32_matmul=True)(x, time=t, condition=c), ref)
    print(f"ACE-sized forward tf32 rel-L2 vs fp32 oracle: {e:.3e}; oracle with TF32-rounded matmul operands: {emu:.3e}")
    assert e < TF32_BOUND


# ---------------------------------------------------------------------------------------------------------
# rollout step glue: normalise + pack / prescribe + denormalise against the reference's own classes
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["mask_int", "interp", "mask_zero"])
def test_step_glue_matches_reference_classes(dev, case):
    """tests/golden/rollout_glue.pt was produced by the reference's StandardNormalizer / Packer / Prescriber
    (tests/golden/make_golden_rollout.py); the two fused kernels reproduce it."""
    import os

    from spherical_dyffusion_b200.rollout import StepGlue

    fx = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rollout_glue.pt"), map_location="cpu",
                    weights_only=False)[case]
    c = fx["spec"]
    glue = StepGlue(c["names"], c["out"], fx["means"], fx["stds"], prescribed_name=c["prescribed"], mask_name=c["mask"],
                    mask_value=c["mask_value"], interpolate=c["interpolate"])
    packed = glue.normalize_pack({k: v.to(dev) for k, v in fx["data"].items()})
    assert packed.shape == fx["packed_norm"].shape
    assert torch.allclose(packed.cpu(), fx["packed_norm"], rtol=1e-6, atol=1e-6)
    p = c["prescribed"]
    target_norm = ((fx["target"][p] - fx["means"][p]) / fx["stds"][p]).to(dev)
    gen = fx["gen"].to(dev).clone()
    out, den = glue.finish(gen, target_norm, fx["target"][c["mask"]].to(dev))
    assert out.data_ptr() == gen.data_ptr()                       # in place: it seeds the next window
    assert torch.allclose(out.cpu(), fx["gen_prescribed"], rtol=1e-6, atol=1e-6)
    assert torch.allclose(den.cpu(), fx["gen_denorm"], rtol=1e-6, atol=1e-4)
    # without a prescriber: denormalise only
    plain = StepGlue(c["names"], c["out"], fx["means"], fx["stds"])
    g2 = fx["gen"].to(dev).clone()
    _, den2 = plain.finish(g2)
    mean = torch.tensor([fx["means"][n] for n in c["out"]]).view(1, -1, 1, 1)
    std = torch.tensor([fx["stds"][n] for n in c["out"]]).view(1, -1, 1, 1)
    assert torch.equal(g2.cpu(), fx["gen"])
    assert torch.allclose(den2.cpu(), fx["gen"] * std + mean, rtol=1e-6, atol=1e-4)


# ---------------------------------------------------------------------------------------------------------
# sub-module boundary (SURVEY 8b "must export"): fused spectral convolution and precision-selectable conv1x1
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("operator_type", ["dhconv", "diagonal"])
@pytest.mark.parametrize("precision,tol", [("fp32", 1e-5), ("tf32", 6e-4), ("bf16", 4.2e-3)])   # 1.5 x measured (4.0e-4 / 2.76e-3)
@pytest.mark.parametrize("grids", [("equiangular", "legendre-gauss"), ("legendre-gauss", "legendre-gauss")])
def test_spectral_conv_module_matches_reference_forward(dev, operator_type, precision, tol, grids):
    """The reference-shaped SpectralConvS2.forward(x) -> (y, residual) (s2convolutions.py:158-193) through the fused
    sfno_spectral_conv entry point in every precision, against the CPU restatement (oracle transforms + contraction):
    the sub-module drop-in of INTEGRATION.md section 2b IS the B200-native path."""
    from spherical_dyffusion_b200.sfnonet import SpectralConvS2

    nlat, nlon, C, B = 36, 72, 32, 3
    lmax, mmax = nlat, nlon // 2 + 1
    g = torch.Generator().manual_seed(11)
    fwd = sb.RealSHT(nlat, nlon, lmax=lmax, mmax=mmax, grid=grids[0], precision=precision)
    inv = sb.InverseRealSHT(nlat, nlon, lmax=lmax, mmax=mmax, grid=grids[1], precision=precision)
    conv = SpectralConvS2(fwd, inv, C, C, operator_type=operator_type, bias=True).to(dev)
    with torch.no_grad():
        conv.weight.copy_(torch.randn(conv.weight.shape, generator=g) * 0.1)
        conv.bias.copy_(torch.randn(conv.bias.shape, generator=g))
    x = torch.randn(B, C, nlat, nlon, generator=g)
    o_fwd = oh.RealSHT(nlat, nlon, lmax=lmax, mmax=mmax, grid=grids[0]).float()
    o_inv = oh.InverseRealSHT(nlat, nlon, lmax=lmax, mmax=mmax, grid=grids[1]).float()
    X = o_fwd(x)
    w = conv.weight.detach().cpu()
    Y = dhconv_contract(X, w) if operator_type == "dhconv" else diagonal_contract(X, w)
    y_ref = o_inv(Y) + conv.bias.detach().cpu()
    with torch.inference_mode():
        y, residual = conv(x.to(dev))
    e = rel_l2(y, y_ref)
    print(f"SpectralConvS2[{operator_type}, {precision}, {grids[0]}->{grids[1]}]: y rel-L2 {e:.3e}")
    assert y.shape == y_ref.shape and e < tol
    if grids[0] != grids[1]:     # scale_residual: residual = inverse(forward(x)) (s2convolutions.py:166-169)
        assert conv.scale_residual and rel_l2(residual, o_inv(X)) < 1.7 * tol   # two transforms, no contraction in between (bf16: 4.6e-3)
    else:
        assert residual.data_ptr() == x.to(dev).data_ptr() or torch.equal(residual.cpu(), x)
    # in-place weight update -> re-packed
    with torch.no_grad():
        conv.weight.mul_(2.0)
    with torch.inference_mode():
        y2, _ = conv(x.to(dev))
    assert rel_l2(y2 - conv.bias, 2.0 * (y_ref - conv.bias.detach().cpu())) < 2 * tol


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-5), ("tf32", 2e-4), ("bf16", 2.4e-3)])   # 1.5 x measured (1.22e-4 / 1.6e-3)
def test_conv1x1_ex_precisions_and_dropout(dev, precision, tol):
    """Precision-selectable 1x1 convolution with the fused epilogue (bias -> GELU -> dropout -> + residual) against torch;
    the dropout mask is a function of the Philox key only, identical on every engine."""
    g = torch.Generator().manual_seed(5)
    B, Ci, Co, H, W = 2, 64, 96, 24, 48
    x = torch.randn(B, Ci, H, W, generator=g)
    w = torch.randn(Co, Ci, 1, 1, generator=g) * 0.1
    b = torch.randn(Co, generator=g)
    r = torch.randn(B, Co, H, W, generator=g)
    prec = _lib.SFNO_PREC[precision]
    ref = torch.nn.functional.gelu(torch.nn.functional.conv2d(x, w, b)) + r
    y = torch.ops.sfno_b200.conv1x1_ex(x.to(dev), w.to(dev), b.to(dev), r.to(dev), 1, 0.0, 0, 0, prec)
    e = rel_l2(y, ref)
    print(f"conv1x1_ex[{precision}]: rel-L2 {e:.3e}")
    assert e < tol
    p = 0.25
    yd = torch.ops.sfno_b200.conv1x1_ex(x.to(dev), w.to(dev), b.to(dev), None, 1, p, 7, 3, prec)
    y0 = torch.ops.sfno_b200.conv1x1_ex(x.to(dev), w.to(dev), b.to(dev), None, 1, 0.0, 0, 0, prec)
    kept = yd != 0
    frac = kept.float().mean().item()
    assert abs(frac - (1 - p)) < 0.01, frac
    assert rel_l2(yd[kept], y0[kept] / (1 - p)) < 1e-5                  # kept elements are scaled by 1 / keep
    mask_fp32 = torch.ops.sfno_b200.conv1x1_ex(x.to(dev), w.to(dev), b.to(dev), None, 1, p, 7, 3, _lib.SFNO_PREC["fp32"]) != 0
    assert (kept != mask_fp32).float().mean().item() < 1e-3             # same mask on every engine (ties of exact zeros aside)
    yd2 = torch.ops.sfno_b200.conv1x1_ex(x.to(dev), w.to(dev), b.to(dev), None, 1, p, 7, 4, prec)
    assert ((yd2 != 0) != kept).float().mean().item() > 0.2            # another offset, another mask


def test_spectral_conv_180x360_triangular_bf16_vs_oracle(dev):
    """The triangular tensor-core kernels (Legendre stores only live degrees, dhconv visits only live wavenumbers, inverse
    Legendre contracts only over live degrees) at the ACE grid against the CPU oracle -- not against the other engine."""
    from spherical_dyffusion_b200.sfnonet import SpectralConvS2

    nlat, nlon, C, B = 180, 360, 64, 2
    g = torch.Generator().manual_seed(21)
    x = torch.randn(B, C, nlat, nlon, generator=g)
    X = oh.RealSHT(nlat, nlon, lmax=180, mmax=181, grid="equiangular").float()(x)
    o_inv = oh.InverseRealSHT(nlat, nlon, lmax=180, mmax=181, grid="legendre-gauss").float()
    for precision, tol in (("bf16", 4.2e-3), ("tf32", 6e-4)):
        fwd = sb.RealSHT(nlat, nlon, lmax=180, mmax=181, grid="equiangular", precision=precision)
        inv = sb.InverseRealSHT(nlat, nlon, lmax=180, mmax=181, grid="legendre-gauss", precision=precision)
        conv = SpectralConvS2(fwd, inv, C, C, operator_type="dhconv", bias=True).to(dev)
        with torch.no_grad():
            conv.weight.copy_(torch.randn(conv.weight.shape, generator=g) * 0.1)
            conv.bias.copy_(torch.randn(conv.bias.shape, generator=g))
        y_ref = o_inv(dhconv_contract(X, conv.weight.detach().cpu())) + conv.bias.detach().cpu()
        with torch.inference_mode():
            y, residual = conv(x.to(dev))
        e, er = rel_l2(y, y_ref), rel_l2(residual, o_inv(X))
        print(f"SpectralConvS2 180x360 dhconv {precision}: y rel-L2 {e:.3e}, residual {er:.3e}")
        assert e < tol and er < 1.7 * tol


def test_mlp_dropout_with_injected_masks(dev):
    """layers.py:73-80 with inference dropout: fc1 -> GELU -> Dropout(p) -> fc2 -> Dropout(p).  The masks the library
    draws are read back from its outputs (an element is dropped iff it is exactly zero) and INJECTED into a torch
    restatement: positions, scaling by 1 / (1 - p) and the order activation -> dropout must then agree to fp32 accuracy."""
    g = torch.Generator().manual_seed(9)
    B, C, hid, H, W, p = 2, 32, 64, 16, 32, 0.1
    x = torch.randn(B, C, H, W, generator=g)
    w1, b1 = torch.randn(hid, C, 1, 1, generator=g) * 0.2, torch.randn(hid, generator=g) * 0.1
    w2, b2 = torch.randn(C, hid, 1, 1, generator=g) * 0.2, torch.randn(C, generator=g) * 0.1
    fp32 = _lib.SFNO_PREC["fp32"]
    h = torch.ops.sfno_b200.conv1x1_ex(x.to(dev), w1.to(dev), b1.to(dev), None, 1, p, 5, 0, fp32)
    y = torch.ops.sfno_b200.conv1x1_ex(h, w2.to(dev), b2.to(dev), None, 0, p, 5, 1, fp32)
    m1, m2 = (h != 0).float().cpu(), (y != 0).float().cpu()
    assert abs(m1.mean().item() - (1 - p)) < 0.01 and abs(m2.mean().item() - (1 - p)) < 0.02
    F = torch.nn.functional
    keep = 1.0 - round(p * 65536) / 65536.0     # the library decides on 16 random bits per element: p is a multiple of 2^-16
    h_ref = F.gelu(F.conv2d(x, w1, b1)) * m1 / keep
    y_ref = F.conv2d(h_ref, w2, b2) * m2 / keep
    assert rel_l2(h, h_ref) < 2e-6 and rel_l2(y, y_ref) < 4e-6


def test_net_dropout_masks_identical_on_both_engines(dev):
    """Inside the net the tensor-core MLP reads its keep bits from masks generated by dropout_mask_kernel; the CUDA-core
    engine evaluates the same Philox blocks inline.  Same key -> same masks: the two bf16 runs differ by rounding noise
    only, while another key changes the output by an order of magnitude more."""
    cfg = SFNOConfig(num_input_channels=6, num_output_channels=6, num_conditional_channels=0, spatial_shape=(32, 64),
                     embed_dim=64, num_layers=3, dropout_mlp=0.2, drop_path_rate=0.0, with_time_emb=False)
    sd = perturb_affine_and_biases(random_state_dict(cfg, seed=4, spectral_gain=16.0))
    m = module_from_cfg(cfg, sd, dev, "bf16")
    x = torch.randn(4, 6, 32, 64, generator=torch.Generator().manual_seed(2)).to(dev)
    with torch.inference_mode(), m.inference_dropout_scope(condition=True):
        m.seed_dropout(3)
        y_tc = m(x).clone()
        m.seed_dropout(4)
        y_other = m(x).clone()
        _lib.set_option("force_simt", 1)
        try:
            m.seed_dropout(3)
            y_simt = m(x).clone()
        finally:
            _lib.set_option("force_simt", 0)
    same, other = rel_l2(y_tc, y_simt), rel_l2(y_tc, y_other)
    print(f"dropout masks: tensor-core (mask kernel) vs CUDA-core (inline Philox), same key {same:.3e}; other key {other:.3e}")
    assert same < 2e-2 and other > 5 * same


# ---------------------------------------------------------------------------------------------------------
# ragged shapes: odd batches, channel counts that break the 32-row store boxes, grids whose tiles have tails
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision,tol", [("fp32", FP32_TOL), ("tf32", TF32_BOUND), ("bf16", BF16_BOUND)])
@pytest.mark.parametrize("B,embed,nlat,nlon,grid", [
    (1, 32, 24, 48, "equiangular"),
    (3, 32, 24, 48, "legendre-gauss"),
    (5, 24, 24, 48, "equiangular"),        # embed % 32 != 0: eight-row boxes in the inverse Legendre transform
    (7, 40, 33, 64, "equiangular"),        # odd nlat, 33 degrees, rows of a tile straddle samples
    (2, 16, 45, 96, "legendre-gauss"),     # nlon / 2 + 1 = 49 wavenumbers
])
def test_net_ragged_shapes_match_oracle(dev, precision, tol, B, embed, nlat, nlon, grid):
    cfg = SFNOConfig(spatial_shape=(nlat, nlon), num_input_channels=5, num_output_channels=3, num_conditional_channels=2, embed_dim=embed,
                     num_layers=2, operator_type="dhconv", with_time_emb=True, data_grid=grid, min_time=0.0, max_time=5.0)
    sd = perturb_affine_and_biases(random_state_dict(cfg, seed=B, spectral_gain=float(embed * embed) / 4.0), seed=B + 1)
    g = torch.Generator().manual_seed(100 + B)
    x = torch.randn(B, 5, nlat, nlon, generator=g)
    c = torch.randn(B, 2, nlat, nlon, generator=g)
    t = torch.rand(B, generator=g) * 5.0
    ref = SFNOOracle(cfg, sd)(x, time=t, condition=c)
    m = module_from_cfg(cfg, sd, dev, precision=precision)
    with torch.inference_mode():
        y = m(x.to(dev), time=t.to(dev), condition=c.to(dev))
        y1 = m(x[-1:].to(dev), time=t[-1:].to(dev), condition=c[-1:].to(dev))     # the last sample alone
    e = rel_l2(y, ref)
    print(f"ragged {precision} B={B} embed={embed} {nlat}x{nlon} {grid}: rel-L2 {e:.3e}; last sample alone vs in batch {rel_l2(y1, y[-1:]):.3e}")
    assert torch.isfinite(y).all()
    assert e < tol
    assert rel_l2(y1, y[-1:]) < (1e-5 if precision == "fp32" else 1.4 * tol)
