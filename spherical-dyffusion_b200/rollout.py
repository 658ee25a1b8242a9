"""Ensemble rollout driver: members sharded over the GPUs of one box, each rank advancing its
members as one batch through DYffusion sampling windows, with per-step ensemble statistics reduced over NCCL.

Replaces, for the synthetic benchmark, the roles of ``run_inference`` / ``run_on_batch_multistep``
(``src/ace_inference/inference/loop.py:158-264``, ``core/stepper_multistep.py:298-466``): the reference loops over
members sequentially at batch 1 and copies every step to the host; here the state stays resident on the GPU and
members are the batch dimension.  Data loading, normalisation, prescribers and writers stay in the reference.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional

import torch

from .dyffusion import DYffusion
from .ensemble import EnsembleStatistics


class EnsembleRollout:
    def __init__(self, sampler: DYffusion, stats: EnsembleStatistics, forcing_fn: Callable[[int, int, torch.device], torch.Tensor],
                 truth_fn: Optional[Callable[[int, torch.device], torch.Tensor]] = None, weights: Optional[torch.Tensor] = None):
        """forcing_fn(window_start_step, n_local, device) -> static_condition [n_local, F, H, W] for that window;
        truth_fn(step, device) -> [C, H, W] verification field (optional, enables rmse / ssr / crps)."""
        self.sampler, self.stats, self.forcing_fn, self.truth_fn, self.weights = sampler, stats, forcing_fn, truth_fn, weights

    @torch.inference_mode()
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
                    out = self.stats.step(members.float(), truth=truth, weights=self.weights)
                    for name, v in out.items():
                        if name in ("mean", "var"):
                            continue
                        history.setdefault(name, []).append(v)
            # the next window starts from the sampler's autoregressive initial state when it produces one
            # (stepper_multistep.py:412-416, forecasting_multi_horizon.py:281), else from the last prediction
            state = preds.get("preds_autoregressive_init", preds[f"t{h}_preds"])
            step += h
        return history
