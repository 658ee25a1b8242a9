"""Constructor keywords of the reference's published model sizes, for benchmarks and examples.

ACE-sized Spherical DYffusion (``src/configs/model/sfno.yaml``, ``experiment/fv3gfs_*.yaml``,
``datamodule/fv3gfs_prescriptive_only.yaml:22-60``): 34 prognostic channels + 2 forcings on the 180 x 360 equiangular
grid, embed 256, 8 blocks, dhconv, scale_factor 1, MLP ratio 2, time embedding.  The forecaster maps 34 (+2) -> 34, the
interpolator 2 x 34 (+2) -> 34 with MLP dropout / DropPath 0.1 (``experiment/fv3gfs_interpolation.yaml:21-23``).
``SCALED_*`` is BASELINE.json's configuration 5 (embed 512, 12 blocks, 0.25 degree grid).
"""
from __future__ import annotations

from typing import Any, Dict

_COMMON: Dict[str, Any] = dict(
    spectral_transform="sht", filter_type="linear", operator_type="dhconv", scale_factor=1, use_mlp=True, mlp_ratio=2.0,
    activation_function="gelu", encoder_layers=1, pos_embed=True, big_skip=True, normalization_layer="instance_norm",
    hard_thresholding_fraction=1.0, with_time_emb=True, time_dim_mult=2, time_rescale=False,
    time_scale_shift_before_filter=True, data_grid="equiangular")

ACE_GRID = (180, 360)

ACE_FORECASTER: Dict[str, Any] = dict(
    _COMMON, embed_dim=256, num_layers=8, num_input_channels=34, num_output_channels=34, num_output_channels_raw=34,
    num_conditional_channels=2, spatial_shape_in=ACE_GRID, spatial_shape_out=ACE_GRID, dropout_mlp=0.0, drop_path_rate=0.0)

ACE_INTERPOLATOR: Dict[str, Any] = dict(
    _COMMON, embed_dim=256, num_layers=8, num_input_channels=68, num_output_channels=34, num_output_channels_raw=34,
    num_conditional_channels=2, spatial_shape_in=ACE_GRID, spatial_shape_out=ACE_GRID, dropout_mlp=0.1, drop_path_rate=0.1)

SCALED_GRID = (720, 1440)

SCALED_FORECASTER: Dict[str, Any] = dict(
    _COMMON, embed_dim=512, num_layers=12, num_input_channels=34, num_output_channels=34, num_output_channels_raw=34,
    num_conditional_channels=2, spatial_shape_in=SCALED_GRID, spatial_shape_out=SCALED_GRID, dropout_mlp=0.0, drop_path_rate=0.0)


def build(config: Dict[str, Any], precision: str = "bf16", seed: int = 0, min_max_time=(0.0, 5.0), **overrides):
    """Random-init module of ``config`` (the reference's initialisers, ``sfnonet.py:732,746-754``) in eval mode on the CPU;
    move it with ``.to(device)``.  ``seed`` fixes the initialisation."""
    import torch

    from .sfnonet import SphericalFourierNeuralOperatorNet

    with torch.random.fork_rng(devices=[]):
        torch.manual_seed(seed)
        model = SphericalFourierNeuralOperatorNet(**{**config, **overrides}, precision=precision)
    if config.get("with_time_emb", False) and min_max_time is not None:
        model.set_min_max_time(*min_max_time)
    return model.eval()
