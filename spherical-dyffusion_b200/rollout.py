"""Ensemble rollout driver: members sharded over the GPUs of one box, each rank advancing its
members as one batch through DYffusion sampling windows, with per-step ensemble statistics reduced over NCCL.

Replaces, for the synthetic benchmark, the roles of ``run_inference`` / ``run_on_batch_multistep``
(``src/ace_inference/inference/loop.py:158-264``, ``core/stepper_multistep.py:298-466``): the reference loops over
members sequentially at batch 1 and copies every step to the host; here the state stays resident on the GPU and
members are the batch dimension.  Data loading, normalisation, prescribers and writers stay in the reference.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence

import torch

from . import _lib
from ._util import stream_ptr
from .dyffusion import DYffusion
from .ensemble import EnsembleStatistics


class StepGlue:
    """Device-side step glue of an autoregressive rollout: what ``run_on_batch_multistep`` does around each
    call of the sampler with dicts of named fields (``stepper_multistep.py:365-427``) -- ``StandardNormalizer``
    (``normalizer.py:57-112``), ``Packer`` (``packer.py:20-77``) and ``Prescriber`` (``prescriber.py:50-95``) -- on packed
    [B, C, H, W] tensors that never leave the GPU, one kernel launch per side of the sampler:

    * ``normalize_pack(fields)``:  dict of raw [B, H, W] fields -> packed normalised input of the sampler;
    * ``finish(gen_norm, target_norm, mask)``: prescriber on the packed normalised prediction IN PLACE (it seeds the next
      window) fused with the denormalisation of the copy that is recorded.

    Same constructor vocabulary as the reference classes: variable names, per-variable means / stds, and the prescriber's
    ``prescribed_name`` / ``mask_name`` / ``mask_value`` / ``interpolate``."""

    def __init__(self, in_names: Sequence[str], out_names: Sequence[str], means: Dict[str, float], stds: Dict[str, float],
                 prescribed_name: Optional[str] = None, mask_name: Optional[str] = None, mask_value: int = 1,
                 interpolate: bool = False):
        self.in_names, self.out_names = list(in_names), list(out_names)
        self.means, self.stds = dict(means), dict(stds)
        if prescribed_name is not None and prescribed_name not in self.out_names:
            raise ValueError(f"Variables which are being prescribed in masked regions must be in out_names, but {prescribed_name} is not.")
        if (prescribed_name is None) != (mask_name is None):
            raise ValueError("prescribed_name and mask_name must be given together")
        self.prescribed_name, self.mask_name, self.mask_value, self.interpolate = prescribed_name, mask_name, int(mask_value), bool(interpolate)
        self._dev: Dict[str, tuple] = {}

    def _stats(self, names, device):
        key = (tuple(names), str(device))
        if key not in self._dev:   # variables without statistics pass through unchanged, as in normalizer._normalize
            mean = torch.tensor([float(self.means.get(n, 0.0)) for n in names], dtype=torch.float32).to(device)
            std = torch.tensor([float(self.stds.get(n, 1.0)) for n in names], dtype=torch.float32).to(device)
            self._dev[key] = (mean, std)
        return self._dev[key]

    def normalize_pack(self, fields: Dict[str, torch.Tensor]) -> torch.Tensor:
        """``Packer(in_names).pack(StandardNormalizer.normalize(fields), axis=-3)`` in one launch."""
        ts = []
        for n in self.in_names:
            t = fields[n]
            if not t.is_cuda:
                raise RuntimeError(f"field {n} must be a CUDA tensor: the B200 path has no CPU fallback")
            ts.append(t.float().contiguous())
        B, H, W = ts[0].shape
        for n, t in zip(self.in_names, ts):
            if tuple(t.shape) != (B, H, W):
                raise ValueError(f"field {n} has shape {tuple(t.shape)}, expected {(B, H, W)}")   # packer.py DataShapesNotUniform
        dev = ts[0].device
        mean, std = self._stats(self.in_names, dev)
        ptrs = torch.tensor([t.data_ptr() for t in ts], dtype=torch.int64).to(dev)
        out = torch.empty(B, len(ts), H, W, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().sfno_normalize_pack(ptrs.data_ptr(), len(ts), B, H * W, mean.data_ptr(), std.data_ptr(), out.data_ptr(),
                                                      stream_ptr(dev)), "sfno_normalize_pack")
        return out

    def finish(self, gen_norm: torch.Tensor, target_norm: Optional[torch.Tensor] = None, mask: Optional[torch.Tensor] = None,
               denormalize: bool = True):
        """Prescriber on ``gen_norm`` [B, C_out, H, W] in place + denormalised copy.  ``target_norm`` [B or 1, H, W] is the
        normalised target of the prescribed variable, ``mask`` [B or 1, H, W] the raw mask variable.  Returns
        ``(gen_norm, gen_denorm or None)``."""
        if not (gen_norm.is_cuda and gen_norm.dtype == torch.float32 and gen_norm.is_contiguous()):
            raise RuntimeError("gen_norm must be a contiguous fp32 CUDA tensor")
        B, C, H, W = gen_norm.shape
        dev = gen_norm.device
        mean, std = self._stats(self.out_names, dev)
        ch = self.out_names.index(self.prescribed_name) if self.prescribed_name is not None else -1
        tb = mb = 0
        if ch >= 0:
            if target_norm is None or mask is None:
                raise ValueError("the prescriber needs the normalised target of the prescribed variable and the mask field")
            target_norm, mask = target_norm.float().contiguous(), mask.float().contiguous()
            tb = H * W if target_norm.shape[0] == B else 0
            mb = H * W if mask.shape[0] == B else 0
            if target_norm.shape[0] not in (1, B) or mask.shape[0] not in (1, B):
                raise ValueError("target_norm / mask must have batch 1 or the batch of gen_norm")
        den = torch.empty_like(gen_norm) if denormalize else None
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().sfno_prescribe_denormalize(
                gen_norm.data_ptr(), target_norm.data_ptr() if ch >= 0 else None, tb, mask.data_ptr() if ch >= 0 else None, mb, ch,
                self.mask_value, int(self.interpolate), mean.data_ptr(), std.data_ptr(), den.data_ptr() if den is not None else None,
                C, B, H * W, stream_ptr(dev)), "sfno_prescribe_denormalize")
        return gen_norm, den


class EnsembleRollout:
    def __init__(self, sampler: DYffusion, stats: EnsembleStatistics, forcing_fn: Callable[[int, int, torch.device], torch.Tensor],
                 truth_fn: Optional[Callable[[int, torch.device], torch.Tensor]] = None, weights: Optional[torch.Tensor] = None):
        """forcing_fn(window_start_step, n_local, device) -> static_condition [n_local, F, H, W] for that window;
        truth_fn(step, device) -> [C, H, W] verification field (optional, enables rmse / ssr / crps)."""
        self.sampler, self.stats, self.forcing_fn, self.truth_fn, self.weights = sampler, stats, forcing_fn, truth_fn, weights
        self.time_stats = False                 # True: bracket every statistics step with CUDA events (bench.py)
        self.stat_events: List[tuple] = []

    @torch.inference_mode()
    def stats_ms(self) -> float:
        """Device time spent in the statistics steps (kernels + collectives) since the last call; synchronises."""
        torch.cuda.synchronize()
        ms = sum(a.elapsed_time(b) for a, b in self.stat_events)
        self.stat_events = []
        return ms

    def run(self, initial_condition: torch.Tensor, n_steps: int, record_every: int = 1) -> Dict[str, List[torch.Tensor]]:
        """initial_condition [C, H, W] (shared by all members; they diverge through the interpolator's dropout stream).
        Returns per-recorded-step lists of the scalar statistics (per channel)."""
        dev = initial_condition.device
        n_local = len(self.stats.local_ids)
        h = self.sampler.horizon   # dynamical steps per window (num_timesteps also counts artificial diffusion steps)
        state = initial_condition.unsqueeze(0).expand(n_local, *initial_condition.shape).contiguous()
        # give every member its own dropout stream
        for net in (self.sampler.model, self.sampler.interpolator):
            if hasattr(net, "seed_dropout"):
                net.seed_dropout(1 + self.stats.rank)
        history: Dict[str, List[torch.Tensor]] = {}
        step = 0
        while step < n_steps:
            forcing = self.forcing_fn(step, n_local, dev)
            preds = self.sampler.sample(state, static_condition=forcing)
            for k in range(1, h + 1):
                if step + k > n_steps:
                    break
                members = preds[f"t{k}_preds"]
                if (step + k) % record_every == 0:
                    truth = self.truth_fn(step + k, dev) if self.truth_fn is not None else None
                    if self.time_stats and members.is_cuda:
                        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                        ev[0].record()
                    out = self.stats.step(members.float(), truth=truth, weights=self.weights)
                    if self.time_stats and members.is_cuda:
                        ev[1].record()
                        self.stat_events.append(ev)
                    for name, v in out.items():
                        if name in ("mean", "var"):
                            continue
                        history.setdefault(name, []).append(v)
            # the next window starts from the sampler's autoregressive initial state when it produces one
            # (stepper_multistep.py:412-416, forecasting_multi_horizon.py:281), else from the last prediction
            state = preds.get("preds_autoregressive_init", preds[f"t{h}_preds"])
            step += h
        return history
