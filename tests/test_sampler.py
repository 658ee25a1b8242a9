"""DYffusion sampling window (caller of the hot path): the loop of spherical_dyffusion_b200.dyffusion against the
window produced by the reference's own sampler (fixtures from tests/golden/make_golden.py).
CPU: loop logic with the oracle standing in for the networks.  GPU: the same loop over the B200 SFNO modules."""
import contextlib
import os

import pytest
import torch

from conftest import GOLDEN_DIR
from oracle.sfno_oracle import SFNOConfig, SFNOOracle, rel_l2

from spherical_dyffusion_b200.dyffusion import DYffusion

SAMPLER_CASES = sorted(f[:-3] for f in os.listdir(GOLDEN_DIR) if f.startswith("dyffusion_") and f.endswith(".pt"))


class OracleNet:
    """Test adapter: the CPU oracle behind the two methods the sampler calls."""

    def __init__(self, cfg, sd):
        self.oracle = SFNOOracle(cfg, sd)
        self.with_time_emb, self.min_time = True, cfg.min_time

    def predict_forward(self, x, time=None, condition=None, static_condition=None, **kw):
        return self.oracle(x, time=time, condition=condition, static_condition=static_condition)

    def inference_dropout_scope(self, condition, context=None):
        return contextlib.nullcontext()


def _load(name):
    return torch.load(os.path.join(GOLDEN_DIR, name + ".pt"), map_location="cpu", weights_only=False)


def _cfg(d):
    d = dict(d)
    d["spatial_shape"] = tuple(d["spatial_shape"])
    return SFNOConfig(**d)


@pytest.mark.parametrize("case", SAMPLER_CASES)
def test_sampler_loop_matches_reference_window_cpu(case):
    fx = _load(case)
    spec = fx["spec"]
    fore = OracleNet(_cfg(fx["forecaster_cfg"]), fx["forecaster_sd"])
    ipol = OracleNet(_cfg(fx["interpolator_cfg"]), fx["interpolator_sd"])
    dy = DYffusion(fore, ipol, timesteps=spec["horizon"], forward_conditioning=spec["forward_conditioning"],
                   time_encoding="dynamics", enable_interpolator_dropout=False, sampling_type=spec.get("sampling_type", "cold"),
                   refine_intermediate_predictions=spec.get("refine", False),
                   additional_interpolation_steps=spec.get("additional_interpolation_steps", 0))
    preds = dy.sample(fx["x0"], **fx["kwargs"])
    assert sorted(k for k in preds if k.endswith("_preds")) == sorted(fx["preds"])
    for k, ref in fx["preds"].items():
        assert rel_l2(preds[k], ref) < 5e-6, k
    counts = dy.forwards_per_window()
    h = spec["horizon"]
    add = spec.get("additional_interpolation_steps", 0)
    assert counts["forecaster"] == h + add   # one forecaster call per diffusion step (dynamical + artificial)


def test_forward_counts_for_ace_window():
    dy = DYffusion(OracleNetStub(), OracleNetStub(), timesteps=6)
    assert dy.forwards_per_window() == {"forecaster": 6, "interpolator": 10}


class OracleNetStub:
    with_time_emb = False

    def predict_forward(self, *a, **k):
        raise AssertionError("not called")


def _b200_module(cfg, sd, dev, precision="fp32"):
    import spherical_dyffusion_b200 as sb

    m = sb.SphericalFourierNeuralOperatorNet(
        num_input_channels=cfg.num_input_channels, num_output_channels=cfg.num_output_channels,
        num_output_channels_raw=cfg.num_output_channels, num_conditional_channels=cfg.num_conditional_channels,
        spatial_shape_in=cfg.spatial_shape, spatial_shape_out=cfg.spatial_shape, precision=precision, **cfg.model_kwargs())
    m.load_state_dict(sd, strict=True)
    m.set_min_max_time(cfg.min_time, cfg.max_time)
    return m.to(dev).eval()


@pytest.mark.gpu
@pytest.mark.parametrize("case", SAMPLER_CASES)
@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16", 5e-2)])
def test_sampler_window_on_b200_matches_reference(case, precision, tol):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    dev = torch.device("cuda:0")
    fx = _load(case)
    spec = fx["spec"]
    fore = _b200_module(_cfg(fx["forecaster_cfg"]), fx["forecaster_sd"], dev, precision)
    ipol = _b200_module(_cfg(fx["interpolator_cfg"]), fx["interpolator_sd"], dev, precision)
    dy = DYffusion(fore, ipol, timesteps=spec["horizon"], forward_conditioning=spec["forward_conditioning"],
                   time_encoding="dynamics", enable_interpolator_dropout=False, sampling_type=spec.get("sampling_type", "cold"),
                   refine_intermediate_predictions=spec.get("refine", False),
                   additional_interpolation_steps=spec.get("additional_interpolation_steps", 0))
    kwargs = {k: v.to(dev) for k, v in fx["kwargs"].items()}
    preds = dy.sample(fx["x0"].to(dev), **kwargs)
    errs = {k: rel_l2(preds[k], ref) for k, ref in fx["preds"].items()}
    print(f"{case}[{precision}]: " + ", ".join(f"{k} {v:.2e}" for k, v in sorted(errs.items())))
    for k, v in errs.items():
        assert v < tol, (k, v)
