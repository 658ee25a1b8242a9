"""Drop-in surface check, build container only (needs /root/reference; skipped on the GPU box, which has no reference
tree and -- here -- there is no GPU): the REFERENCE's own ``DYffusion`` sampler (``src/diffusion/dyffusion.py``, imported
unchanged through ``oracle/ref_shim.py``) runs over THIS package's ``SphericalFourierNeuralOperatorNet`` modules and
reproduces the committed reference windows.  Everything the reference sampler and ``BaseDiffusion.__init__`` touch on a
model -- constructor bookkeeping, ``hparams``, channel / shape attributes, ``predict_forward``, ``inference_dropout_scope``,
``set_min_max_time``, ``state_dict`` loading -- is this package's code; only the arithmetic of ``forward`` is delegated to
the CPU oracle, because the product's forward needs a GPU (the same modules run that arithmetic on the B200 in
``tests/test_sampler.py::test_sampler_window_on_b200_matches_reference``).  INTEGRATION.md section 1 cites this test."""
import os
import sys

import pytest
import torch

from conftest import GOLDEN_DIR, ROOT
from oracle import ref_shim
from oracle.sfno_oracle import SFNOConfig, SFNOOracle, rel_l2

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="needs the reference tree (/root/reference)")

CASES = ["dyffusion_window_12x24_h4", "dyffusion_window_12x24_h3_dyn", "dyffusion_window_12x24_h3_arinit"]


def _cfg(d):
    d = dict(d)
    d["spatial_shape"] = tuple(d["spatial_shape"])
    return SFNOConfig(**d)


def _our_module_with_oracle_arithmetic(cfg, sd):
    import spherical_dyffusion_b200 as sb

    m = sb.SphericalFourierNeuralOperatorNet(
        num_input_channels=cfg.num_input_channels, num_output_channels=cfg.num_output_channels,
        num_output_channels_raw=cfg.num_output_channels, num_conditional_channels=cfg.num_conditional_channels,
        spatial_shape_in=cfg.spatial_shape, spatial_shape_out=cfg.spatial_shape, **cfg.model_kwargs())
    m.load_state_dict(sd, strict=True)
    m.set_min_max_time(cfg.min_time, cfg.max_time)
    m.eval()
    oracle = SFNOOracle(cfg, {k: v for k, v in m.state_dict().items()})

    def forward(inputs, time=None, condition=None, static_condition=None, return_time_emb=False, **kw):
        m._input_parts(inputs, condition, static_condition)      # this package's validation of the call
        return oracle(inputs, time=time, condition=condition, static_condition=static_condition)

    m.forward = forward
    return m


@pytest.mark.parametrize("case", CASES)
def test_reference_sampler_runs_over_this_packages_modules(case):
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from make_golden import _InterpolatorHandle     # the duck-typed InterpolationExperiment the fixtures were made with

    ref_shim.install()
    from src.diffusion.dyffusion import DYffusion as RefDYffusion

    fx = torch.load(os.path.join(GOLDEN_DIR, case + ".pt"), map_location="cpu", weights_only=False)
    spec = fx["spec"]
    fore = _our_module_with_oracle_arithmetic(_cfg(fx["forecaster_cfg"]), fx["forecaster_sd"])
    ipol = _our_module_with_oracle_arithmetic(_cfg(fx["interpolator_cfg"]), fx["interpolator_sd"])
    dy = RefDYffusion(model=fore, timesteps=spec["horizon"], interpolator=_InterpolatorHandle(ipol, spec["horizon"]),
                      interpolator_local_checkpoint_path=None, forward_conditioning=spec["forward_conditioning"],
                      time_encoding=spec.get("time_encoding", "dynamics"), enable_interpolator_dropout=False,
                      sampling_type=spec.get("sampling_type", "cold"),
                      use_cold_sampling_for_last_step=spec.get("use_cold_sampling_for_last_step", True),
                      use_cold_sampling_for_init_of_ar_step=spec.get("use_cold_sampling_for_init_of_ar_step"))
    with torch.inference_mode():
        preds = dy.sample(fx["x0"], **fx["kwargs"])
    got = {k: v for k, v in preds.items() if k.endswith("_preds") or k == "preds_autoregressive_init"}
    assert sorted(got) == sorted(fx["preds"])
    for k, ref in fx["preds"].items():
        assert rel_l2(got[k], ref) < 5e-6, k
