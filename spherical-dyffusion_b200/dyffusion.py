"""DYffusion cold-sampling loop around the B200 SFNO forward (caller of the hot path).

Mirror of ``BaseDYffusion.sample_loop`` / ``q_sample`` / ``predict_x_last`` and ``DYffusion._interpolate``
(``src/diffusion/dyffusion.py:190-240,286-355,457-567,642-662``) for inference with a frozen interpolator:
same schedule mapping (``diffusion_step_to_interpolation_step`` :128-184), same time encodings, same conditioning
rules, same dictionary of outputs (``t{k}_preds``).  The two networks are ``SphericalFourierNeuralOperatorNet``
modules of this package (or anything with ``predict_forward`` / ``inference_dropout_scope``); Lightning,
checkpoint lookup and the training loss stay in the reference.

The loop is host-synchronisation free: time tensors are built on the device and the range asserts of the reference
(`.all()` on device tensors, ``dyffusion.py:144-146,311,651-653``) are evaluated on the host from the Python schedule.
"""
from __future__ import annotations

import math
from contextlib import ExitStack
from typing import Dict, List, Optional, Sequence, Union

import numpy as np
import torch
from torch import Tensor


class DYffusion:
    def __init__(
        self,
        model,                      # forecaster  F(x_s, time(s), condition) -> x_{t+h}
        interpolator,               # interpolator I(cat(x_0, x_h), time i_n) -> x_{i_n}
        timesteps: int,             # horizon h (datamodule.horizon)
        forward_conditioning: str = "none",
        dynamic_cond_from_t: str = "h",
        schedule: str = "before_t1_only",
        additional_interpolation_steps: int = 0,
        additional_interpolation_steps_factor: int = 0,
        interpolate_before_t1: bool = True,
        sampling_type: str = "cold",
        sampling_schedule: Optional[Sequence[float]] = None,
        use_cold_sampling_for_intermediate_steps: bool = True,
        use_cold_sampling_for_last_step: bool = True,
        use_cold_sampling_for_init_of_ar_step: Optional[bool] = None,
        time_encoding: str = "dynamics",
        refine_intermediate_predictions: bool = False,
        enable_interpolator_dropout: Union[bool, str] = True,
        log_every_t: Optional[int] = None,
        hack_for_imprecise_interpolation: bool = False,
        interpolator_horizon: Optional[int] = None,
    ):
        if forward_conditioning not in ("data", "none"):
            raise ValueError(f"forward_conditioning={forward_conditioning!r} is not built (reference options with noise are training-time variants)")
        if enable_interpolator_dropout not in (True, False, "always", "except_dynamical_steps"):
            raise ValueError(f"Invalid enable_interpolator_dropout: {enable_interpolator_dropout}")
        self.model, self.interpolator = model, interpolator
        self.forward_conditioning = forward_conditioning
        self.dynamic_cond_from_t = dynamic_cond_from_t
        self.schedule = schedule
        self.sampling_type = sampling_type
        self.use_cold_sampling_for_intermediate_steps = use_cold_sampling_for_intermediate_steps
        self.use_cold_sampling_for_last_step = use_cold_sampling_for_last_step
        self.use_cold_sampling_for_init_of_ar_step = (use_cold_sampling_for_init_of_ar_step
                                                      if use_cold_sampling_for_init_of_ar_step is not None
                                                      else use_cold_sampling_for_last_step)
        self.time_encoding = time_encoding
        self.refine_intermediate_predictions = refine_intermediate_predictions
        self.enable_interpolator_dropout = enable_interpolator_dropout
        self.log_every_t = log_every_t
        self.hack_for_imprecise_interpolation = hack_for_imprecise_interpolation
        self.training = False

        horizon = timesteps
        assert horizon > 1, f"horizon must be > 1, but got {horizon}"
        self.num_timesteps = timesteps
        # dyffusion.py:63-97
        if schedule == "linear":
            assert additional_interpolation_steps == 0
            self.additional_interpolation_steps_fac = additional_interpolation_steps_factor
            if interpolate_before_t1:
                interpolated_steps, self.di_to_ti_add = horizon - 1, 0
            else:
                interpolated_steps, self.di_to_ti_add = horizon - 2, additional_interpolation_steps_factor
            self.additional_diffusion_steps = additional_interpolation_steps_factor * interpolated_steps
        elif schedule == "before_t1_only":
            assert additional_interpolation_steps_factor == 0 and interpolate_before_t1
            self.additional_diffusion_steps = additional_interpolation_steps
        elif schedule == "before_t1_then_linear":
            assert interpolate_before_t1
            self.additional_interpolation_steps_fac = additional_interpolation_steps_factor
            self.additional_diffusion_steps_pre_t1 = additional_interpolation_steps
            self.additional_diffusion_steps = additional_interpolation_steps + additional_interpolation_steps_factor * (horizon - 2)
        else:
            raise ValueError(f"Invalid schedule: {schedule}")
        self.num_timesteps += self.additional_diffusion_steps
        d_to_i = {d: self.diffusion_step_to_interpolation_step(d) for d in range(1, self.num_timesteps)}
        self.dynamical_steps = {d: i for d, i in d_to_i.items() if float(i).is_integer()}
        self.artificial_interpolation_steps = {d: i for d, i in d_to_i.items() if not float(i).is_integer()}
        self.sampling_schedule = list(sampling_schedule) if sampling_schedule else list(range(0, self.num_timesteps))
        for a, b in zip(self.sampling_schedule, self.sampling_schedule[1:]):
            assert b > a, f"Invalid sampling schedule not monotonically increasing: {self.sampling_schedule}"
        # dyffusion.py:632-640
        self.interpolator_horizon = interpolator_horizon if interpolator_horizon is not None else horizon
        last = self.diffusion_step_to_interpolation_step(self.num_timesteps - 1)
        if self.interpolator_horizon != last + 1:
            raise ValueError(f"interpolator horizon {self.interpolator_horizon} must be equal to the last interpolation step+1={last + 1}")
        # time ranges the networks will see (forecasting_multi_horizon.py:52-57, interpolation.py:24-25)
        if hasattr(model, "set_min_max_time") and getattr(model, "with_time_emb", False) and model.min_time is None:
            valid = self.valid_time_range_for_backbone_model
            model.set_min_max_time(min(valid), max(valid))
        if hasattr(interpolator, "set_min_max_time") and getattr(interpolator, "with_time_emb", False) and interpolator.min_time is None:
            interpolator.set_min_max_time(1, self.interpolator_horizon - 1) if self.additional_diffusion_steps == 0 else \
                interpolator.set_min_max_time(0, self.interpolator_horizon - 1)

    # ---- schedule (dyffusion.py:128-184) ---------------------------------------------------------------------------
    def diffusion_step_to_interpolation_step(self, d: Union[int, float]) -> float:
        assert 0 <= d <= self.num_timesteps - 1, f"diffusion_step must be in [0, {self.num_timesteps - 1}], but got {d}"
        if self.schedule == "linear":
            return (d + self.di_to_ti_add) / (self.additional_interpolation_steps_fac + 1)
        if self.schedule == "before_t1_only":
            if d >= self.additional_diffusion_steps + 1:
                return d - self.additional_diffusion_steps
            return d / (self.additional_diffusion_steps + 1)
        if d >= self.additional_diffusion_steps_pre_t1 + 1:
            return 1 + (d - self.additional_diffusion_steps_pre_t1 - 1) / (self.additional_interpolation_steps_fac + 1)
        return d / (self.additional_diffusion_steps_pre_t1 + 1)

    @property
    def valid_time_range_for_backbone_model(self) -> List[float]:
        steps = list(range(0, self.num_timesteps))
        if self.time_encoding == "discrete":
            return steps
        if self.time_encoding == "continuous":
            return list(np.array(steps) / self.num_timesteps)
        if self.time_encoding == "dynamics":
            return [self.diffusion_step_to_interpolation_step(d) for d in steps]
        raise ValueError(f"Invalid time_encoding: {self.time_encoding}")

    # ---- forecaster call (dyffusion.py:286-355) -----------------------------------------------------------------------
    def predict_x_last(self, initial_condition: Tensor, x_t: Tensor, t: int, dynamical_condition: Tensor = None, **kwargs):
        assert 0 <= t <= self.num_timesteps - 1, f"Invalid timestep: {t}. {self.num_timesteps=}"
        B, dev = initial_condition.shape[0], initial_condition.device
        forward_inputs = initial_condition if self.forward_conditioning == "data" else None
        dyn = None
        if dynamical_condition is not None:
            assert dynamical_condition.shape[1] == self.num_timesteps + 1, f"{dynamical_condition.shape}[1] != {self.num_timesteps + 1}"
            idx = {"0": 0, "h": -1, "t": int(t)}[self.dynamic_cond_from_t]
            dyn = dynamical_condition[:, idx]
        if forward_inputs is not None and dyn is not None:
            condition = torch.cat([forward_inputs, dyn], dim=1)
        else:
            condition = forward_inputs if forward_inputs is not None else dyn
        if self.time_encoding == "discrete":
            time = float(t)
        elif self.time_encoding == "continuous":
            time = t / self.num_timesteps
        else:
            time = float(self.diffusion_step_to_interpolation_step(t))
        time = torch.full((B,), time, dtype=torch.float32, device=dev)
        return self.model.predict_forward(x_t, time=time, condition=condition, **kwargs)

    # ---- interpolator call (dyffusion.py:190-240, 642-662) ---------------------------------------------------------------
    def q_sample(self, x0: Tensor, x_end: Tensor, t: Optional[int], interpolation_time: Optional[float] = None,
                 is_artificial_step: bool = True, dynamical_condition: Tensor = None, num_predictions: int = 1, **kwargs) -> Tensor:
        assert t is None or interpolation_time is None, "Either t or interpolation_time must be None."
        i_n = interpolation_time if t is None else self.diffusion_step_to_interpolation_step(t)
        assert 0 < i_n < self.interpolator_horizon, f"interpolate time must be in (0, {self.interpolator_horizon}), got {i_n}"
        if dynamical_condition is not None:
            assert float(i_n).is_integer(), "a dynamical condition needs an integer interpolation time"
            kwargs["condition"] = dynamical_condition[:, int(i_n)]  # interpolation.py:133-141
        B, dev = x0.shape[0], x0.device
        time = torch.full((B,), float(i_n), dtype=torch.float32, device=dev)
        do_enable = (self.training or self.enable_interpolator_dropout in (True, "always")
                     or (self.enable_interpolator_dropout == "except_dynamical_steps" and is_artificial_step))
        with ExitStack() as stack:
            if hasattr(self.interpolator, "inference_dropout_scope"):
                stack.enter_context(self.interpolator.inference_dropout_scope(condition=bool(do_enable)))
            x_last = x0
            if self.hack_for_imprecise_interpolation:
                x_last = torch.cat([x_end[:, :1], x_last], dim=1)
            out = self.interpolator.predict_forward(torch.cat([x_end, x_last], dim=1), time=time, **kwargs)
            if isinstance(out, dict):
                out = out["preds"]
            if self.hack_for_imprecise_interpolation:
                out = torch.cat([x_end[:, :1], out], dim=1)
        return out

    # ---- the loop (dyffusion.py:457-567) -------------------------------------------------------------------------------------
    @torch.inference_mode()
    def sample_loop(self, initial_condition: Tensor, log_every_t: Optional[int] = None, num_predictions: int = None, **kwargs):
        log_every_t = log_every_t or self.log_every_t
        sched = self.sampling_schedule
        assert initial_condition.dim() == 4, f"condition.shape: {initial_condition.shape} (should be 4D)"
        intermediates: Dict[str, Tensor] = {}
        xhat_th, dynamics_pred_step = None, 0
        last_plus_one = sched[-1] + 1
        triples = zip(sched, sched[1:] + [last_plus_one], sched[2:] + [last_plus_one, last_plus_one + 1])
        x_s = initial_condition
        for s, s_next, s_nnext in triples:
            is_first_step = s == 0
            is_last_step = s == self.num_timesteps - 1
            xhat_th = self.predict_x_last(initial_condition=initial_condition, x_t=x_s, t=s, **kwargs)
            time_i_n = self.diffusion_step_to_interpolation_step(s_next) if not is_last_step else np.inf
            is_dynamics_pred = float(time_i_n).is_integer() or is_last_step
            q_kwargs = dict(x0=xhat_th, x_end=initial_condition, is_artificial_step=not is_dynamics_pred,
                            num_predictions=num_predictions if is_first_step else 1)
            if s_next <= self.num_timesteps - 1:
                x_ip_next = self.q_sample(**q_kwargs, t=s_next, **kwargs)
            else:
                assert is_last_step, f"Invalid s_next: {s_next} (should be <= {self.num_timesteps - 1})"
                x_ip_next = xhat_th
                if self.hack_for_imprecise_interpolation:
                    x_ip_next = torch.cat([initial_condition[:, :1], x_ip_next], dim=1)
            x_ip_s = None
            if self.sampling_type == "cold":
                if not self.use_cold_sampling_for_last_step and is_last_step:
                    if self.use_cold_sampling_for_init_of_ar_step:
                        x_ip_s = self.q_sample(**q_kwargs, t=s, **kwargs)
                        ar_init = x_s + xhat_th - x_ip_s
                        if self.hack_for_imprecise_interpolation:
                            ar_init = ar_init[:, 1:]
                        intermediates["preds_autoregressive_init"] = ar_init
                    x_s = xhat_th
                else:
                    x_ip_s = self.q_sample(**q_kwargs, t=s, **kwargs) if s > 0 else x_s
                    x_s = x_s + (x_ip_next - x_ip_s)  # cold sampling update
            elif self.sampling_type == "naive":
                x_s = x_ip_next
            else:
                raise ValueError(f"unknown sampling type {self.sampling_type}")
            dynamics_pred_step = int(time_i_n) if s < self.num_timesteps - 1 else dynamics_pred_step + 1
            if is_dynamics_pred:
                preds_t = x_s if (self.use_cold_sampling_for_intermediate_steps or is_last_step) else x_ip_next
                if self.hack_for_imprecise_interpolation:
                    preds_t = preds_t[:, 1:]
                intermediates[f"t{dynamics_pred_step}_preds"] = preds_t
                if log_every_t is not None:
                    intermediates[f"t{dynamics_pred_step}_preds2"] = x_ip_next
            if log_every_t is not None:
                intermediates[f"x_{s}_dmodel"] = x_s
                intermediates[f"intermediate_{s}_x0hat"] = xhat_th
                intermediates[f"xipol_{s}_dmodel"] = x_ip_next
                if self.sampling_type == "cold" and x_ip_s is not None:
                    intermediates[f"xipol_{s}_dmodel2"] = x_ip_s
        if self.refine_intermediate_predictions:
            for i_n in [i for i in self.dynamical_steps.values() if i < self.num_timesteps]:
                key = f"t{int(i_n) if float(i_n).is_integer() else i_n}_preds"
                assert not float(i_n).is_integer() or key in intermediates, f"{key} not in intermediates"
                out = self.q_sample(x0=xhat_th, x_end=initial_condition, is_artificial_step=False, t=None, interpolation_time=i_n, **kwargs)
                intermediates[key] = out[:, 1:] if self.hack_for_imprecise_interpolation else out
        if last_plus_one < self.num_timesteps:
            return x_s, intermediates
        return xhat_th, intermediates

    @torch.inference_mode()
    def sample(self, initial_condition: Tensor, num_samples: int = 1, **kwargs) -> Dict[str, Tensor]:
        """``dyffusion.py:569-572``: returns the dictionary of predictions ``t1_preds .. t{h}_preds``."""
        _, intermediates = self.sample_loop(initial_condition, **kwargs)
        return intermediates

    def predict_forward(self, inputs: Tensor, condition: Tensor = None, **kwargs) -> Dict[str, Tensor]:
        """``dyffusion.py:574-582``."""
        assert inputs is not None or condition is not None
        initial_condition = inputs if inputs is not None else condition
        if inputs is not None and condition is not None:
            initial_condition = torch.cat([inputs, condition], dim=1)
        return self.sample(initial_condition, **kwargs)

    def forwards_per_window(self) -> Dict[str, int]:
        """Number of network calls of one sampling window (6 forecaster + 10 interpolator at h = 6, k = 0)."""
        nf = ni = 0
        sched = self.sampling_schedule
        last_plus_one = sched[-1] + 1
        for s, s_next in zip(sched, sched[1:] + [last_plus_one]):
            nf += 1
            if s_next <= self.num_timesteps - 1:
                ni += 1
            if self.sampling_type == "cold" and s > 0 and (self.use_cold_sampling_for_last_step or s != self.num_timesteps - 1):
                ni += 1
        return {"forecaster": nf, "interpolator": ni}
