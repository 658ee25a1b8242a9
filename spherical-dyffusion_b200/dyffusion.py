"""DYffusion sampling window around the B200 SFNO forward (caller of the hot path).

What it computes is ``BaseDYffusion.sample_loop`` with ``q_sample`` / ``predict_x_last`` / ``DYffusion._interpolate``
(``src/diffusion/dyffusion.py:190-240,286-355,457-567,642-662``) for inference with a frozen interpolator: the same
schedule mapping (``diffusion_step_to_interpolation_step`` :128-184), time encodings, conditioning rules and output
dictionary (``t{k}_preds``, ``preds_autoregressive_init``).  How it runs is different:

* **Window program.**  The reference walks the schedule in Python and decides per step what to call.  Every decision
  depends only on constructor options, so here the schedule is *compiled once* into a flat program of four instruction
  kinds -- ``Forecast``, ``Interpolate``, ``Update``, ``Emit`` -- over a handful of named tensor slots
  (``WindowProgram``).  Running a window is a loop over that program: no branching on tensors, no host synchronisation
  (time tensors are filled on the device; the reference's `.all()` range asserts, ``dyffusion.py:144-146,311,651-653``,
  are checked on the host from the schedule when the program is built).
* **Fused update.**  ``x_s + (x_interpolated_s_next - x_interpolated_s)`` (``dyffusion.py:519``) is one kernel
  (``sfno_cold_update``: three reads, one write) instead of two elementwise passes and a temporary.
* **One CUDA graph per window.**  With ``capture_graph=True`` the whole program -- 16 SFNO forwards (≈ 1 200 kernel
  launches) plus the glue at horizon 6 -- is captured once per input signature and replayed.  Interpolator dropout stays
  live under replay because the networks key their masks on a device-resident Philox state that each forward advances
  (``sfno_net_forward_parts_rng``), which is what makes a window with inference dropout capturable at all.

The two networks are ``SphericalFourierNeuralOperatorNet`` modules of this package (or anything with ``predict_forward``
/ ``inference_dropout_scope``); Lightning, checkpoint lookup and the training loss stay in the reference.
"""
from __future__ import annotations

from contextlib import ExitStack
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch
from torch import Tensor


# ---- the window program --------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class Forecast:
    """xhat <- forecaster(x, time(s), condition)   (``predict_x_last``, dyffusion.py:286-355)"""
    s: int
    time: float


@dataclass(frozen=True)
class Interpolate:
    """slot <- interpolator(cat(ic, xhat), time i_n)   (``q_sample`` :190-240 / ``_interpolate`` :642-662)"""
    i_n: float
    artificial: bool     # an artificial (non-integer) interpolation step: dropout policy "except_dynamical_steps"
    dst: str


@dataclass(frozen=True)
class Update:
    """How x advances at the end of a diffusion step (dyffusion.py:504-524).
    cold: x <- x + (next - cur);  cold0: the same with cur = x (first step);  next: x <- next;  xhat: x <- xhat;
    last: next <- xhat (the final step has nothing to interpolate to);  arinit: ar <- x + (xhat - cur)"""
    kind: str


@dataclass(frozen=True)
class Emit:
    """out[key] <- slot (minus the leading helper channel of ``hack_for_imprecise_interpolation``)"""
    key: str
    src: str
    strip: bool


Instruction = Union[Forecast, Interpolate, Update, Emit]


@dataclass(frozen=True)
class WindowProgram:
    code: Tuple[Instruction, ...]
    result: str                      # slot returned next to the dictionary ("x" or "xhat", dyffusion.py:563-567)

    def count(self, kind) -> int:
        return sum(isinstance(op, kind) for op in self.code)


class _GraphedWindow:
    """One captured window: static input buffers, the graph, static outputs."""

    def __init__(self, owner: "DYffusion", program: WindowProgram, ic: Tensor, kwargs: Dict[str, Tensor]):
        dev = ic.device
        self.ic = ic.clone()
        self.kwargs = {k: v.clone() for k, v in kwargs.items()}
        # eager warm-up on a side stream: creates the native nets, packs the parameters, sizes the workspaces and the
        # Philox states -- nothing of that may happen inside the capture
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            owner._execute(program, self.ic, self.kwargs)
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.result, self.outputs = owner._execute(program, self.ic, self.kwargs)

    def replay(self, ic: Tensor, kwargs: Dict[str, Tensor], clone: bool):
        self.ic.copy_(ic)
        for k, v in kwargs.items():
            self.kwargs[k].copy_(v)
        self.graph.replay()
        if not clone:
            return self.result, dict(self.outputs)
        return self.result.clone(), {k: v.clone() for k, v in self.outputs.items()}


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
        capture_graph: bool = False,
        graph_outputs: str = "clone",
    ):
        if forward_conditioning not in ("data", "none"):
            raise ValueError(f"forward_conditioning={forward_conditioning!r} is not built (reference options with noise are training-time variants)")
        if enable_interpolator_dropout not in (True, False, "always", "except_dynamical_steps"):
            raise ValueError(f"Invalid enable_interpolator_dropout: {enable_interpolator_dropout}")
        if sampling_type not in ("cold", "naive"):
            raise ValueError(f"unknown sampling type {sampling_type}")
        if graph_outputs not in ("clone", "view"):
            raise ValueError(f"graph_outputs must be 'clone' or 'view', got {graph_outputs!r}")
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
        self.capture_graph = capture_graph
        self.graph_outputs = graph_outputs
        self.training = False

        horizon = timesteps
        assert horizon > 1, f"horizon must be > 1, but got {horizon}"
        self.horizon = horizon             # dynamical steps of one window (keys t1_preds .. t{horizon}_preds)
        self.num_timesteps = timesteps     # diffusion steps: dynamical + artificial ones (dyffusion.py:63-97)
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
        # time ranges the networks will see (forecasting_multi_horizon.py:52-57; interpolation.py:23-25: ALWAYS
        # [1, horizon - 1] for the interpolator -- the rescale constants of a trained interpolator depend on it)
        if hasattr(model, "set_min_max_time") and getattr(model, "with_time_emb", False) and model.min_time is None:
            valid = self.valid_time_range_for_backbone_model
            model.set_min_max_time(min(valid), max(valid))
        if hasattr(interpolator, "set_min_max_time") and getattr(interpolator, "with_time_emb", False) and interpolator.min_time is None:
            interpolator.set_min_max_time(1, self.interpolator_horizon - 1)
            if self.artificial_interpolation_steps and hasattr(interpolator, "check_time_range"):
                # artificial steps query the interpolator at fractional times below 1 (dyffusion.py:139-146 allows (0, h));
                # only the range ASSERT is relaxed for them, never the rescale constants
                interpolator.check_time_range = False
        self._programs: Dict[bool, WindowProgram] = {}
        self._graphs: Dict[tuple, _GraphedWindow] = {}

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

    def _forecaster_time(self, t: int) -> float:
        if self.time_encoding == "discrete":
            return float(t)
        if self.time_encoding == "continuous":
            return t / self.num_timesteps
        return float(self.diffusion_step_to_interpolation_step(t))

    # ---- compile the schedule into a window program -------------------------------------------------------------------
    def program(self, log: bool = False) -> WindowProgram:
        """The instruction list of one window.  Built once per ``log`` flag; every range check the reference makes on
        device tensors at run time is made here on the host."""
        if log in self._programs:
            return self._programs[log]
        T, sched, strip = self.num_timesteps, self.sampling_schedule, self.hack_for_imprecise_interpolation
        code: List[Instruction] = []
        dyn_step = 0

        def interpolate(d: int, artificial: bool, dst: str):
            i_n = self.diffusion_step_to_interpolation_step(d)
            assert 0 < i_n < self.interpolator_horizon, f"interpolate time must be in (0, {self.interpolator_horizon}), got {i_n}"
            code.append(Interpolate(float(i_n), artificial, dst))

        for pos, s in enumerate(sched):
            s_next = sched[pos + 1] if pos + 1 < len(sched) else sched[-1] + 1
            assert 0 <= s <= T - 1, f"Invalid timestep: {s}. {self.num_timesteps=}"
            final = s == T - 1
            code.append(Forecast(int(s), self._forecaster_time(s)))
            i_next = self.diffusion_step_to_interpolation_step(s_next) if not final else np.inf
            lands_on_dynamics = final or float(i_next).is_integer()
            if s_next <= T - 1:
                interpolate(s_next, not lands_on_dynamics, "next")
            else:
                assert final, f"Invalid s_next: {s_next} (should be <= {T - 1})"
                code.append(Update("last"))
            has_cur = False
            if self.sampling_type == "naive":
                code.append(Update("next"))
            elif final and not self.use_cold_sampling_for_last_step:
                if self.use_cold_sampling_for_init_of_ar_step:
                    interpolate(s, not lands_on_dynamics, "cur")
                    has_cur = True
                    code.append(Update("arinit"))
                    code.append(Emit("preds_autoregressive_init", "ar", strip))
                code.append(Update("xhat"))
            elif s > 0:
                interpolate(s, not lands_on_dynamics, "cur")
                has_cur = True
                code.append(Update("cold"))
            else:
                code.append(Update("cold0"))
            dyn_step = int(i_next) if s < T - 1 else dyn_step + 1
            if lands_on_dynamics:
                src = "x" if (self.use_cold_sampling_for_intermediate_steps or final) else "next"
                code.append(Emit(f"t{dyn_step}_preds", src, strip))
                if log:
                    code.append(Emit(f"t{dyn_step}_preds2", "next", False))
            if log:
                code += [Emit(f"x_{s}_dmodel", "x", False), Emit(f"intermediate_{s}_x0hat", "xhat", False),
                         Emit(f"xipol_{s}_dmodel", "next", False)]
                if has_cur:
                    code.append(Emit(f"xipol_{s}_dmodel2", "cur", False))
        if self.refine_intermediate_predictions:
            # one more pass of the interpolator from the FINAL forecast for every intermediate dynamical step (:545-561)
            emitted = {op.key for op in code if isinstance(op, Emit)}
            for i_n in [i for i in self.dynamical_steps.values() if i < T]:
                key = f"t{int(i_n) if float(i_n).is_integer() else i_n}_preds"
                assert not float(i_n).is_integer() or key in emitted, f"{key} not in intermediates"
                assert 0 < i_n < self.interpolator_horizon
                code.append(Interpolate(float(i_n), False, "refined"))
                code.append(Emit(key, "refined", strip))
        result = "x" if sched[-1] + 1 < T else "xhat"
        self._programs[log] = WindowProgram(tuple(code), result)
        return self._programs[log]

    # ---- the two network calls -----------------------------------------------------------------------------------------
    def predict_x_last(self, initial_condition: Tensor, x_t: Tensor, t: int, dynamical_condition: Tensor = None, **kwargs):
        """Forecaster call (dyffusion.py:286-355): conditioning on the initial condition and / or one slice of the
        dynamical condition, time by ``time_encoding``."""
        assert 0 <= t <= self.num_timesteps - 1, f"Invalid timestep: {t}. {self.num_timesteps=}"
        cond = [initial_condition] if self.forward_conditioning == "data" else []
        if dynamical_condition is not None:
            assert dynamical_condition.shape[1] == self.num_timesteps + 1, f"{dynamical_condition.shape}[1] != {self.num_timesteps + 1}"
            cond.append(dynamical_condition[:, {"0": 0, "h": -1, "t": int(t)}[self.dynamic_cond_from_t]])
        condition = None if not cond else (cond[0] if len(cond) == 1 else torch.cat(cond, dim=1))
        time = torch.full((initial_condition.shape[0],), self._forecaster_time(t), dtype=torch.float32, device=initial_condition.device)
        return self.model.predict_forward(x_t, time=time, condition=condition, **kwargs)

    def q_sample(self, x0: Tensor, x_end: Tensor, t: Optional[int], interpolation_time: Optional[float] = None,
                 is_artificial_step: bool = True, dynamical_condition: Tensor = None, num_predictions: int = 1, **kwargs) -> Tensor:
        """Interpolator call (dyffusion.py:190-240, 642-662): ``x0`` is the forecast of the window's end, ``x_end`` the
        initial condition (the reference's argument names)."""
        assert t is None or interpolation_time is None, "Either t or interpolation_time must be None."
        i_n = interpolation_time if t is None else self.diffusion_step_to_interpolation_step(t)
        assert 0 < i_n < self.interpolator_horizon, f"interpolate time must be in (0, {self.interpolator_horizon}), got {i_n}"
        if dynamical_condition is not None:
            assert float(i_n).is_integer(), "a dynamical condition needs an integer interpolation time"
            kwargs["condition"] = dynamical_condition[:, int(i_n)]  # interpolation.py:133-141
        time = torch.full((x0.shape[0],), float(i_n), dtype=torch.float32, device=x0.device)
        live = (self.training or self.enable_interpolator_dropout in (True, "always")
                or (self.enable_interpolator_dropout == "except_dynamical_steps" and is_artificial_step))
        lead = x_end[:, :1] if self.hack_for_imprecise_interpolation else None
        pieces = [x_end] + ([lead] if lead is not None else []) + [x0]
        with ExitStack() as stack:
            if hasattr(self.interpolator, "inference_dropout_scope"):
                stack.enter_context(self.interpolator.inference_dropout_scope(condition=bool(live)))
            out = self.interpolator.predict_forward(torch.cat(pieces, dim=1), time=time, **kwargs)
        if isinstance(out, dict):
            out = out["preds"]
        return out if lead is None else torch.cat([lead, out], dim=1)

    # ---- run a program ------------------------------------------------------------------------------------------------
    @staticmethod
    def _cold(x: Tensor, nxt: Tensor, cur: Tensor) -> Tensor:
        """x + (nxt - cur) as one fused kernel of the library (``sfno_cold_update``).  There is no CPU path: the custom op
        raises for CPU tensors (the host-logic tests, where CPU stand-ins replace the networks, substitute this method
        in ``tests/conftest.py``)."""
        return torch.ops.sfno_b200.cold_update(x, nxt, cur)

    def _execute(self, prog: WindowProgram, ic: Tensor, kwargs: Dict) -> Tuple[Tensor, Dict[str, Tensor]]:
        slot: Dict[str, Tensor] = {"x": ic}
        out: Dict[str, Tensor] = {}
        for op in prog.code:
            if isinstance(op, Forecast):
                slot["xhat"] = self.predict_x_last(initial_condition=ic, x_t=slot["x"], t=op.s, **kwargs)
            elif isinstance(op, Interpolate):
                slot[op.dst] = self.q_sample(x0=slot["xhat"], x_end=ic, t=None, interpolation_time=op.i_n,
                                             is_artificial_step=op.artificial, **kwargs)
            elif isinstance(op, Update):
                if op.kind == "cold":
                    slot["x"] = self._cold(slot["x"], slot["next"], slot["cur"])
                elif op.kind == "cold0":
                    slot["x"] = self._cold(slot["x"], slot["next"], slot["x"])
                elif op.kind == "next":
                    slot["x"] = slot["next"]
                elif op.kind == "xhat":
                    slot["x"] = slot["xhat"]
                elif op.kind == "last":
                    slot["next"] = (torch.cat([ic[:, :1], slot["xhat"]], dim=1) if self.hack_for_imprecise_interpolation
                                    else slot["xhat"])
                else:  # arinit
                    slot["ar"] = self._cold(slot["x"], slot["xhat"], slot["cur"])
            else:
                out[op.key] = slot[op.src][:, 1:] if op.strip else slot[op.src]
        return slot[prog.result], out

    @torch.inference_mode()
    def sample_loop(self, initial_condition: Tensor, log_every_t: Optional[int] = None, num_predictions: int = None,
                    graph: Optional[bool] = None, **kwargs):
        """One sampling window -> (final state, dictionary of predictions).  ``graph=True`` (or ``capture_graph`` at
        construction) replays a CUDA graph of the window, captured on first use per input signature."""
        assert initial_condition.dim() == 4, f"condition.shape: {initial_condition.shape} (should be 4D)"
        prog = self.program(log=(log_every_t or self.log_every_t) is not None)
        use_graph = self.capture_graph if graph is None else graph
        if not use_graph or not initial_condition.is_cuda:
            return self._execute(prog, initial_condition, kwargs)
        tensors = {k: v for k, v in kwargs.items() if torch.is_tensor(v)}
        if len(tensors) != len(kwargs):
            raise ValueError("a captured window takes tensor keyword arguments only (dynamical_condition / static_condition)")
        handles = tuple(getattr(getattr(net, "_net", None), "value", None) for net in (self.model, self.interpolator))
        key = (prog, tuple(initial_condition.shape), initial_condition.device.index, handles,
               tuple(sorted((k, tuple(v.shape)) for k, v in tensors.items())))
        win = self._graphs.get(key)
        if win is None:
            if len(self._graphs) >= 4:   # signatures change rarely (batch size of the last window of a rollout)
                self._graphs.pop(next(iter(self._graphs)))
            win = self._graphs[key] = _GraphedWindow(self, prog, initial_condition, tensors)
            # (the key was formed before the warm-up created the native nets: store it under the final handles too)
            handles = tuple(getattr(getattr(net, "_net", None), "value", None) for net in (self.model, self.interpolator))
            self._graphs[(prog, key[1], key[2], handles, key[4])] = win
        # parameters may have been swapped (EMA) since the capture: the packed copies the graph reads are refreshed eagerly
        for net in (self.model, self.interpolator):
            if hasattr(net, "refresh_parameters"):
                net.refresh_parameters()
        return win.replay(initial_condition, tensors, clone=self.graph_outputs == "clone")

    @torch.inference_mode()
    def sample(self, initial_condition: Tensor, num_samples: int = 1, **kwargs) -> Dict[str, Tensor]:
        """``dyffusion.py:569-572``: returns the dictionary of predictions ``t1_preds .. t{h}_preds``."""
        return self.sample_loop(initial_condition, **kwargs)[1]

    def predict_forward(self, inputs: Tensor, condition: Tensor = None, **kwargs) -> Dict[str, Tensor]:
        """``dyffusion.py:574-582``."""
        assert inputs is not None or condition is not None
        parts = [t for t in (inputs, condition) if t is not None]
        return self.sample(parts[0] if len(parts) == 1 else torch.cat(parts, dim=1), **kwargs)

    def forwards_per_window(self) -> Dict[str, int]:
        """Number of network calls of one sampling window (6 forecaster + 10 interpolator at h = 6, k = 0)."""
        prog = self.program()
        return {"forecaster": prog.count(Forecast), "interpolator": prog.count(Interpolate)}

    def release_graphs(self) -> None:
        self._graphs.clear()
