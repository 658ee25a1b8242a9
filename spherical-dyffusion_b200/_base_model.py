"""Host-side mirror of the reference ``BaseModel`` (``src/models/_base_model.py:23-300``) for the inference
path: channel bookkeeping, ``concat_condition_if_needed``, ``predict_forward``, inference-dropout scope.
Lightning/Hydra are deliberately absent (plumbing stays in the reference); ``hparams`` is an attribute dict.
Training entry points (``get_loss``) are out of scope for this path and raise.
"""
from __future__ import annotations

import logging
from contextlib import contextmanager
from typing import Any, Optional, Sequence, Union

import torch
import torch.nn as nn
from torch import Tensor


class AttrDict(dict):
    __getattr__ = dict.get

    def __setattr__(self, k, v):
        self[k] = v


class DropPath(nn.Module):
    """Marker module for stochastic depth (``src/models/modules/drop_path.py:25-36``); the per-sample factor
    is drawn inside the fused fc2 epilogue, the module only carries ``drop_prob`` and the train/eval switch."""

    def __init__(self, drop_rate=None):
        super().__init__()
        self.drop_prob = drop_rate

    def forward(self, x):  # never called on the fused path
        raise RuntimeError("DropPath is executed inside sfno_net_forward")


ALL_DROPOUT_LAYERS = (nn.Dropout, nn.Dropout2d, nn.Dropout3d, nn.AlphaDropout, nn.FeatureAlphaDropout, DropPath)


def enable_inference_dropout(model: nn.Module):
    """``src/utilities/utils.py:686-691``: put every dropout layer (incl. DropPath) in training mode."""
    for m in model.modules():
        if isinstance(m, ALL_DROPOUT_LAYERS):
            m.train()


def disable_inference_dropout(model: nn.Module):
    """``src/utilities/utils.py:695-700``."""
    for m in model.modules():
        if isinstance(m, ALL_DROPOUT_LAYERS):
            m.eval()


class BaseModel(nn.Module):
    def __init__(
        self,
        num_input_channels: int = None,
        num_output_channels: int = None,
        num_output_channels_raw: int = None,
        num_conditional_channels: int = 0,
        spatial_shape_in: Union[Sequence[int], int] = None,
        spatial_shape_out: Union[Sequence[int], int] = None,
        loss_function: Optional[str] = None,
        loss_function_weights: Optional[dict] = None,
        datamodule_config: Any = None,
        debug_mode: bool = False,
        name: str = "",
        verbose: bool = True,
    ):
        super().__init__()
        self.hparams = AttrDict(
            num_input_channels=num_input_channels, num_output_channels=num_output_channels,
            num_output_channels_raw=num_output_channels_raw, num_conditional_channels=num_conditional_channels,
            spatial_shape_in=spatial_shape_in, spatial_shape_out=spatial_shape_out, loss_function=loss_function,
            loss_function_weights=loss_function_weights, datamodule_config=datamodule_config, debug_mode=debug_mode,
            name=name)
        self.log_text = logging.getLogger(self.__class__.__name__ if name == "" else name)
        self.name = name
        self.verbose = verbose
        if not verbose:
            self.log_text.setLevel(logging.WARN)
        self.num_input_channels = num_input_channels
        self.num_output_channels = num_output_channels
        self.num_output_channels_raw = num_output_channels_raw
        self.num_conditional_channels = num_conditional_channels
        self.spatial_shape_in = spatial_shape_in
        self.spatial_shape_out = spatial_shape_out
        self.datamodule_config = datamodule_config
        self.criterion = None
        self._channel_dim = None
        self.ema_scope = None  # may be set by the experiment module (_base_experiment.py:386-401)

    @property
    def short_description(self) -> str:
        return self.name if self.name else self.__class__.__name__

    def get_parameters(self) -> list:
        return list(self.parameters())

    @property
    def num_params(self):
        return sum(p.numel() for p in self.get_parameters() if p.requires_grad)

    @property
    def channel_dim(self):
        if self._channel_dim is None:
            self._channel_dim = 1
        return self._channel_dim

    @property
    def device(self):
        try:
            return next(self.parameters()).device
        except StopIteration:
            return torch.device("cpu")

    def concat_condition_if_needed(self, inputs: Tensor, condition: Tensor = None, static_condition: Tensor = None):
        """``_base_model.py:166-192`` (same branches, same error types)."""
        if self.num_conditional_channels > 0:
            if condition is None and static_condition is None:
                raise ValueError(
                    f"condition and static_condition are both None but num_conditional_channels is {self.num_conditional_channels}")
            elif condition is not None and static_condition is not None:
                condition = torch.cat((condition, static_condition), dim=1)
            elif condition is None:
                condition = static_condition
            if hasattr(self, "upsample_condition"):
                condition = self.upsample_condition(condition)
            try:
                x = torch.cat((inputs, condition), dim=1)
            except RuntimeError as e:
                raise RuntimeError(f"inputs.shape: {inputs.shape}, condition.shape: {condition.shape}") from e
        else:
            x = inputs
            assert condition is None, "condition is not None but num_conditional_channels is 0"
            assert static_condition is None, "static_condition is not None but num_conditional_channels is 0"
        return x

    def get_loss(self, *args, **kwargs):
        raise NotImplementedError("training (get_loss / backward) is out of scope of the B200 inference path (SURVEY 8f-4)")

    def predict_forward(self, *inputs: Tensor, metadata: Any = None, **kwargs):
        """``_base_model.py:265-270``."""
        return self(*inputs, **kwargs)

    @contextmanager
    def inference_dropout_scope(self, condition: bool, context=None):
        """``_base_model.py:273-286``."""
        assert isinstance(condition, bool), f"Condition must be a boolean, got {condition}"
        if condition:
            enable_inference_dropout(self)
        try:
            yield None
        finally:
            if condition:
                disable_inference_dropout(self)

    def enable_inference_dropout(self):
        enable_inference_dropout(self)

    def disable_inference_dropout(self):
        disable_inference_dropout(self)
