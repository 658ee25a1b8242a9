"""Generate the committed golden fixtures from the reference's OWN Python classes.

Run in the build container only (needs ``/root/reference``):

    python tests/golden/make_golden.py

For each case it instantiates the reference ``SphericalFourierNeuralOperatorNet``
(``/root/reference/src/models/sfno/sfnonet.py:340``) through ``oracle/ref_shim.py`` with the
reference's initialisers, perturbs biases / norm affines so they are exercised, runs
``forward(inputs, time=, condition=)`` under ``torch.inference_mode()`` and stores config, weights,
inputs, output and a few hooked intermediates (outputs of the reference's own sub-modules) in
``tests/golden/<case>.pt``.  Cases are small (fixtures of a few hundred kB) so they can be committed;
full-size parity is covered by the travelling oracle, itself pinned by these fixtures.
"""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from oracle.sfno_oracle import SFNOConfig  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # tiny dhconv model on a 12x24 grid, both block kinds (first/last with scale_residual, middle without)
    "sfno_dhconv_12x24": dict(
        cfg=dict(num_input_channels=3, num_output_channels=3, num_conditional_channels=2, spatial_shape=(12, 24),
                 embed_dim=16, num_layers=3, operator_type="dhconv", data_grid="equiangular"),
        batch=2, seed=0, times=[1.0, 3.0]),
    # odd-ish sizes: nlat not a multiple of 8, mmax = nlon/2+1 even, different channel counts, LG data grid
    "sfno_dhconv_18x36_lg": dict(
        cfg=dict(num_input_channels=5, num_output_channels=4, num_conditional_channels=1, spatial_shape=(18, 36),
                 embed_dim=24, num_layers=2, operator_type="dhconv", data_grid="legendre-gauss", mlp_ratio=2.0),
        batch=3, seed=1, times=[0.0, 2.5, 5.0]),
    # time shift after the filter, no big skip, no pos-embed, 4 layers
    "sfno_dhconv_16x32_variants": dict(
        cfg=dict(num_input_channels=4, num_output_channels=4, num_conditional_channels=0, spatial_shape=(16, 32),
                 embed_dim=32, num_layers=4, operator_type="dhconv", data_grid="equiangular",
                 time_scale_shift_before_filter=False, big_skip=False, pos_embed=False),
        batch=1, seed=2, times=[4.0]),
    # no time embedding at all
    "sfno_dhconv_12x24_notime": dict(
        cfg=dict(num_input_channels=2, num_output_channels=2, num_conditional_channels=0, spatial_shape=(12, 24),
                 embed_dim=16, num_layers=2, operator_type="dhconv", data_grid="equiangular", with_time_emb=False),
        batch=2, seed=3, times=None),
    # interpolator-shaped call: two stacked snapshots + forcing as input, odd batch (dhconv rows (m,b) with B = 3),
    # embed 40 (2C = 80: partial K block), 3 blocks; only the block outputs are kept as taps (fixture size)
    "sfno_dhconv_24x48_interp": dict(
        cfg=dict(num_input_channels=6, num_output_channels=3, num_conditional_channels=2, spatial_shape=(24, 48),
                 embed_dim=40, num_layers=3, operator_type="dhconv", data_grid="equiangular"),
        batch=3, seed=5, times=[0.0, 2.0, 5.0], taps="out"),
    # batch 1, Gaussian data grid, embed 64, two blocks (both with the SHT round-trip residual)
    "sfno_dhconv_24x48_lg_b1": dict(
        cfg=dict(num_input_channels=4, num_output_channels=4, num_conditional_channels=1, spatial_shape=(24, 48),
                 embed_dim=64, num_layers=2, operator_type="dhconv", data_grid="legendre-gauss"),
        batch=1, seed=6, times=[3.0], taps="out"),
    # interpolator-style module in eval mode: dropout layers exist (state-dict key mlp.fwd.3, layers.py:76-80) but are
    # inactive, DropPath modules likewise; mlp_ratio 1
    "sfno_dhconv_16x32_dropout_eval": dict(
        cfg=dict(num_input_channels=4, num_output_channels=2, num_conditional_channels=1, spatial_shape=(16, 32),
                 embed_dim=24, num_layers=3, operator_type="dhconv", data_grid="equiangular", mlp_ratio=1.0,
                 dropout_mlp=0.1, drop_path_rate=0.1),
        batch=2, seed=7, times=[1.0, 2.0], taps="out"),
    # the 'diagonal' operator (ctor default of the reference)
    "sfno_diagonal_12x24": dict(
        cfg=dict(num_input_channels=3, num_output_channels=3, num_conditional_channels=2, spatial_shape=(12, 24),
                 embed_dim=16, num_layers=2, operator_type="diagonal", data_grid="equiangular"),
        batch=2, seed=4, times=[2.0, 5.0]),
}


def perturb(model, seed):
    g = torch.Generator().manual_seed(1000 + seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.endswith(".bias") or ("norm" in name and name.endswith(".weight")):
                p.add_(0.1 * torch.randn(p.shape, generator=g))
            if name.endswith("filter.filter.weight"):
                # make the spectral branch visible end to end
                p.mul_(float(p.shape[0]))


def make_case(name, spec):
    cfg = SFNOConfig(**spec["cfg"])
    model = ref_shim.build_reference_sfno(
        num_input_channels=cfg.num_input_channels, num_output_channels=cfg.num_output_channels,
        num_conditional_channels=cfg.num_conditional_channels, spatial_shape=cfg.spatial_shape,
        seed=spec["seed"], min_max_time=(cfg.min_time, cfg.max_time), **cfg.model_kwargs())
    perturb(model, spec["seed"])
    g = torch.Generator().manual_seed(2000 + spec["seed"])
    B = spec["batch"]
    H, W = cfg.spatial_shape
    inputs = torch.randn(B, cfg.num_input_channels, H, W, generator=g)
    condition = torch.randn(B, cfg.num_conditional_channels, H, W, generator=g) if cfg.num_conditional_channels else None
    time = torch.tensor(spec["times"], dtype=torch.float32) if spec["times"] is not None else None

    taps = {}

    def hook(key):
        def fn(mod, args, out):
            taps[key] = out[0].clone() if isinstance(out, tuple) else out.clone()
        return fn

    handles = [model.blocks[0].filter.filter.forward_transform.register_forward_hook(hook("blocks.0.sht"))]
    for i, blk in enumerate(model.blocks):
        handles.append(blk.filter.register_forward_hook(hook(f"blocks.{i}.filter_out")))
        handles.append(blk.norm0.register_forward_hook(hook(f"blocks.{i}.norm0_out")))
        handles.append(blk.register_forward_hook(hook(f"blocks.{i}.out")))
    with torch.inference_mode():
        out, t_repr = model(inputs, time=time, condition=condition, return_time_emb=True)
    for h in handles:
        h.remove()
    if spec.get("taps") == "out":   # smaller fixture: what the parity tests read
        taps = {k: v for k, v in taps.items() if k.endswith(".out") or k == "blocks.0.sht"}

    fixture = dict(
        cfg=spec["cfg"], state_dict={k: v.clone() for k, v in model.state_dict().items()},
        inputs=inputs, condition=condition, time=time, output=out.clone(),
        t_repr=None if t_repr is None else t_repr.clone(), taps=taps,
        torch_version=torch.__version__,
    )
    path = os.path.join(OUT, f"{name}.pt")
    torch.save(fixture, path)
    print(f"{name}: out {tuple(out.shape)} std {out.std():.4f}  -> {path} ({os.path.getsize(path) / 1024:.0f} kB)")


if __name__ == "__main__":
    only = set(sys.argv[1:])
    for name, spec in CASES.items():
        if not only or name in only:
            make_case(name, spec)


# ---------------------------------------------------------------------------------------------------------------
# DYffusion sampling window: the reference's own sampler (src/diffusion/dyffusion.py:457-567) driving the reference's
# own SFNO forecaster + interpolator (interpolator dropout off so the window is deterministic).
# ---------------------------------------------------------------------------------------------------------------
SAMPLER_CASES = {
    "dyffusion_window_12x24_h4": dict(horizon=4, channels=3, forcing=2, spatial_shape=(12, 24), embed_dim=16, num_layers=2,
                                      batch=2, seed=10, forward_conditioning="none", condition_kind="static"),
    "dyffusion_window_12x24_h3_dyn": dict(horizon=3, channels=2, forcing=1, spatial_shape=(12, 24), embed_dim=16, num_layers=2,
                                          batch=2, seed=11, forward_conditioning="data", condition_kind="dynamical"),
    # naive sampling (x_{s+1} = interpolate(x0, x0_hat) without the cold-sampling correction) + refinement of the
    # intermediate predictions at the end of the window (dyffusion.py:541-565)
    "dyffusion_window_12x24_h4_naive_refine": dict(horizon=4, channels=2, forcing=2, spatial_shape=(12, 24), embed_dim=16, num_layers=2,
                                                   batch=2, seed=12, forward_conditioning="none", condition_kind="static",
                                                   sampling_type="naive", refine=True),
    # two artificial interpolation steps before t1 (schedule "before_t1_only", dyffusion.py:63-97,128-184): the
    # interpolator is called at fractional times in (0, 1), the forecaster at the matching "dynamics" time encodings
    "dyffusion_window_12x24_h3_addsteps": dict(horizon=3, channels=2, forcing=1, spatial_shape=(12, 24), embed_dim=16, num_layers=2,
                                               batch=2, seed=13, forward_conditioning="none", condition_kind="static",
                                               additional_interpolation_steps=2),
    # "linear" schedule with one artificial step between every pair of dynamical steps, discrete time encoding, and no
    # cold-sampling correction on the last step
    "dyffusion_window_12x24_h3_linear": dict(horizon=3, channels=2, forcing=1, spatial_shape=(12, 24), embed_dim=16, num_layers=2,
                                             batch=2, seed=14, forward_conditioning="none", condition_kind="static",
                                             schedule="linear", additional_interpolation_steps_factor=1, time_encoding="discrete",
                                             use_cold_sampling_for_last_step=False),
    # the released-checkpoint set-up: one input-only channel travels in front of the state
    # (hack_for_imprecise_interpolation, dyffusion.py:41-44,501-502,655-661)
    "dyffusion_window_12x24_h3_hack": dict(horizon=3, channels=2, forcing=1, spatial_shape=(12, 24), embed_dim=16, num_layers=2,
                                           batch=2, seed=15, forward_conditioning="none", condition_kind="static", hack=True),
    # no cold correction on the last step, but a cold-corrected initial state for the next autoregressive window
    # (preds_autoregressive_init, dyffusion.py:505-512); the reference cannot combine this with the hack above
    # (x_s and xhat_th differ by the extra channel at :508)
    "dyffusion_window_12x24_h3_arinit": dict(horizon=3, channels=2, forcing=1, spatial_shape=(12, 24), embed_dim=16, num_layers=2,
                                             batch=2, seed=16, forward_conditioning="none", condition_kind="static",
                                             use_cold_sampling_for_last_step=False, use_cold_sampling_for_init_of_ar_step=True),
}


class _InterpolatorHandle(torch.nn.Module):
    """Duck-typed stand-in for InterpolationExperiment: only what DYffusion touches."""

    def __init__(self, model, horizon):
        super().__init__()
        self.model = model
        self.window, self.true_horizon = 1, horizon
        self.ema_scope = None

    def predict_packed(self, *inputs, **kwargs):
        out = self.model.predict_forward(*inputs, **kwargs)
        return {"preds": out}

    def inference_dropout_scope(self, condition, context=None):
        return self.model.inference_dropout_scope(condition=condition, context=context)

    def get_dynamical_condition(self, dynamical_condition, target_time):
        if dynamical_condition is None:
            return None
        if isinstance(target_time, int):
            return dynamical_condition[:, target_time, ...]
        return dynamical_condition[torch.arange(dynamical_condition.shape[0]), target_time.long(), ...]


def make_sampler_case(name, spec):
    ref_shim.install()
    from src.diffusion.dyffusion import DYffusion

    h, C, F = spec["horizon"], spec["channels"], spec["forcing"]
    shape = spec["spatial_shape"]
    fc_cond = F + (C if spec["forward_conditioning"] == "data" else 0)
    common = dict(spatial_shape=shape, embed_dim=spec["embed_dim"], num_layers=spec["num_layers"], operator_type="dhconv",
                  data_grid="equiangular")
    Cx = C + (1 if spec.get("hack") else 0)   # state channels the sampler carries (input-only channel in front)
    fcfg = SFNOConfig(num_input_channels=Cx, num_output_channels=C, num_conditional_channels=fc_cond, min_time=0.0,
                      max_time=float(h - 1), **common)
    icfg = SFNOConfig(num_input_channels=2 * Cx, num_output_channels=C, num_conditional_channels=F, min_time=1.0,
                      max_time=float(h - 1), **common)
    n_diff = h + spec.get("additional_interpolation_steps", 0) + spec.get("additional_interpolation_steps_factor", 0) * (h - 1)
    fmax = (n_diff - 1) if spec.get("time_encoding", "dynamics") == "discrete" else (h - 1)
    fcfg.max_time = float(fmax)
    forecaster = ref_shim.build_reference_sfno(num_input_channels=Cx, num_output_channels=C, num_conditional_channels=fc_cond,
                                               spatial_shape=shape, seed=spec["seed"], min_max_time=(0, fmax), **fcfg.model_kwargs())
    add = spec.get("additional_interpolation_steps", 0)
    fac = spec.get("additional_interpolation_steps_factor", 0)
    tenc = spec.get("time_encoding", "dynamics")
    imin = 0 if (add or fac) else 1   # fractional interpolation times in (0, 1) when artificial steps exist (dyffusion.py:632-640)
    icfg.min_time = float(imin)
    interp = ref_shim.build_reference_sfno(num_input_channels=2 * Cx, num_output_channels=C, num_conditional_channels=F,
                                           spatial_shape=shape, seed=spec["seed"] + 100, min_max_time=(imin, h - 1), **icfg.model_kwargs())
    perturb(forecaster, spec["seed"])
    perturb(interp, spec["seed"] + 100)
    dy = DYffusion(model=forecaster, timesteps=h, interpolator=_InterpolatorHandle(interp, h), interpolator_local_checkpoint_path=None,
                   forward_conditioning=spec["forward_conditioning"], time_encoding=tenc, enable_interpolator_dropout=False,
                   sampling_type=spec.get("sampling_type", "cold"), refine_intermediate_predictions=spec.get("refine", False),
                   schedule=spec.get("schedule", "before_t1_only"), additional_interpolation_steps=add,
                   additional_interpolation_steps_factor=fac,
                   use_cold_sampling_for_last_step=spec.get("use_cold_sampling_for_last_step", True),
                   use_cold_sampling_for_init_of_ar_step=spec.get("use_cold_sampling_for_init_of_ar_step"),
                   hack_for_imprecise_interpolation=spec.get("hack", False))
    g = torch.Generator().manual_seed(3000 + spec["seed"])
    B = spec["batch"]
    x0 = torch.randn(B, Cx, *shape, generator=g)
    kwargs = {}
    if spec["condition_kind"] == "static":
        kwargs["static_condition"] = torch.randn(B, F, *shape, generator=g)
    else:
        kwargs["dynamical_condition"] = torch.randn(B, h + 1, F, *shape, generator=g)
    with torch.inference_mode():
        preds = dy.sample(x0, **kwargs)
    fixture = dict(spec=spec, forecaster_cfg={k: v for k, v in fcfg.__dict__.items()}, interpolator_cfg={k: v for k, v in icfg.__dict__.items()},
                   forecaster_sd={k: v.clone() for k, v in forecaster.state_dict().items()},
                   interpolator_sd={k: v.clone() for k, v in interp.state_dict().items()},
                   x0=x0, kwargs=kwargs, preds={k: v.clone() for k, v in preds.items() if k.endswith("_preds") or k == "preds_autoregressive_init"},
                   torch_version=torch.__version__)
    path = os.path.join(OUT, f"{name}.pt")
    torch.save(fixture, path)
    print(f"{name}: keys {sorted(fixture['preds'])} -> {path} ({os.path.getsize(path) / 1024:.0f} kB)")


if __name__ == "__main__":
    only = set(sys.argv[1:])
    for name, spec in SAMPLER_CASES.items():
        if not only or name in only:
            make_sampler_case(name, spec)
