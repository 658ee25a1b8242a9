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
    for name, spec in CASES.items():
        make_case(name, spec)
