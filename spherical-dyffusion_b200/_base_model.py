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


_CTOR_FIELDS = ("num_input_channels", "num_output_channels", "num_output_channels_raw", "num_conditional_channels",
                "spatial_shape_in", "spatial_shape_out", "loss_function", "loss_function_weights", "datamodule_config",
                "debug_mode", "name")


class BaseModel(nn.Module):
    """Keyword surface of the reference constructor (``_base_model.py:46-60``); everything is recorded in ``hparams`` and
    the channel / shape bookkeeping the diffusion wrapper reads (``_base_diffusion.py:31-35``) is exposed as attributes."""

    def __init__(self, num_input_channels: int = None, num_output_channels: int = None, num_output_channels_raw: int = None,
                 num_conditional_channels: int = 0, spatial_shape_in: Union[Sequence[int], int] = None,
                 spatial_shape_out: Union[Sequence[int], int] = None, loss_function: Optional[str] = None,
                 loss_function_weights: Optional[dict] = None, datamodule_config: Any = None, debug_mode: bool = False,
                 name: str = "", verbose: bool = True):
        super().__init__()
        given = locals()
        self.hparams = AttrDict({k: given[k] for k in _CTOR_FIELDS})
        for k in _CTOR_FIELDS[:6] + ("datamodule_config", "name"):   # channels, shapes, datamodule config, name
            setattr(self, k, given[k])
        self.verbose = verbose
        self.log_text = logging.getLogger(name or type(self).__name__)
        if not verbose:
            self.log_text.setLevel(logging.WARN)
        # src/losses/losses.py:66-79 (the two criteria the released configurations use; "preds" key as _base_model.py:80-103)
        self.criterion = None
        if loss_function is not None:
            name = (loss_function if isinstance(loss_function, str) else loss_function.get("_target_", "").split(".")[-1])
            name = name.lower().strip().replace("-", "_")
            if name in ("l1", "mae", "mean_absolute_error"):
                self.criterion = nn.ModuleDict({"preds": nn.L1Loss()})
            elif name in ("l2", "mse", "mean_squared_error"):
                self.criterion = nn.ModuleDict({"preds": nn.MSELoss()})
            else:
                raise ValueError(f"Unknown loss function {name}")
        self.loss_function_weights = loss_function_weights if loss_function_weights is not None else {}
        self.ema_scope = None      # may be set by the experiment module (_base_experiment.py:386-401)
        self._channel_dim = 1      # NCHW

    # ---- small read-only surface ------------------------------------------------------------------------------------
    short_description = property(lambda self: self.name or type(self).__name__)
    channel_dim = property(lambda self: self._channel_dim)
    num_params = property(lambda self: sum(p.numel() for p in self.get_parameters() if p.requires_grad))

    def get_parameters(self) -> list:
        return [*self.parameters()]

    @property
    def device(self):
        first = next(self.parameters(), None)
        return first.device if first is not None else torch.device("cpu")

    # ---- conditioning (``_base_model.py:166-192``: same outcomes, same exception classes) ---------------------------------
    def merged_condition(self, condition: Optional[Tensor], static_condition: Optional[Tensor]) -> Optional[Tensor]:
        """The tensor appended to the inputs on the channel axis, or None for an unconditional model."""
        if self.num_conditional_channels <= 0:
            assert condition is None, "condition is not None but num_conditional_channels is 0"
            assert static_condition is None, "static_condition is not None but num_conditional_channels is 0"
            return None
        present = [c for c in (condition, static_condition) if c is not None]
        if not present:
            raise ValueError(
                f"condition and static_condition are both None but num_conditional_channels is {self.num_conditional_channels}")
        merged = present[0] if len(present) == 1 else torch.cat(present, dim=1)
        upsample = getattr(self, "upsample_condition", None)
        return upsample(merged) if upsample is not None else merged

    def concat_condition_if_needed(self, inputs: Tensor, condition: Tensor = None, static_condition: Tensor = None):
        extra = self.merged_condition(condition, static_condition)
        if extra is None:
            return inputs
        try:
            return torch.cat((inputs, extra), dim=1)
        except RuntimeError as e:
            raise RuntimeError(f"inputs.shape: {inputs.shape}, condition.shape: {extra.shape}") from e

    # ---- entry points the experiment / diffusion wrappers call -------------------------------------------------------------
    def get_loss(self, inputs, targets, raw_targets=None, condition=None, metadata: Any = None, predictions_mask=None,
                 return_predictions: bool = False, predictions_post_process=None, targets_pre_process=None, **kwargs):
        """``_base_model.py:194-263`` (tensor predictions): forward, optional post-processing / masking, criterion.  The
        forward runs through the differentiable custom ops when gradients are enabled, so ``loss["loss"].backward()``
        backpropagates through the library's adjoint kernels."""
        if self.criterion is None:
            raise ValueError("the model was built without a loss_function")

        def mask_data(data):
            return data[..., predictions_mask] if predictions_mask is not None else data

        predictions = self(inputs, condition=condition, **kwargs) if torch.is_tensor(inputs) else self(**inputs, condition=condition, **kwargs)
        if predictions_post_process is not None:
            predictions = predictions_post_process(predictions)
        preds_m, targets_m = mask_data(predictions), mask_data(targets)
        assert preds_m.shape == targets_m.shape, \
            f"predictions {tuple(preds_m.shape)} and targets {tuple(targets_m.shape)} differ in shape (a missing singleton dimension would broadcast silently)"
        assert len(self.loss_function_weights) == 0, "Loss function weights are not supported for this case"
        loss_dict = dict(loss=self.criterion["preds"](preds_m, targets_m))
        if return_predictions:
            return loss_dict, preds_m      # the reference returns the (post-processed) predictions AFTER the mask (:235-236)
        return loss_dict

    def predict_forward(self, *inputs: Tensor, metadata: Any = None, **kwargs):
        """``_base_model.py:265-270``: plain call; ``metadata`` is accepted and unused, as in the reference."""
        return self(*inputs, **kwargs)

    @contextmanager
    def inference_dropout_scope(self, condition: bool, context=None):
        """``_base_model.py:273-286``: dropout layers (and DropPath) sample inside the scope when ``condition`` is True."""
        assert isinstance(condition, bool), f"Condition must be a boolean, got {condition}"
        if not condition:
            yield None
            return
        self.enable_inference_dropout()
        try:
            yield None
        finally:
            self.disable_inference_dropout()

    def enable_inference_dropout(self):
        enable_inference_dropout(self)

    def disable_inference_dropout(self):
        disable_inference_dropout(self)
