"""TEST INFRASTRUCTURE ONLY -- stand-alone CPU restatement of the reference SFNO forward.

This is the travelling oracle: plain PyTorch on CPU, no dependency on ``/root/reference``, driven by
a config dict and a reference-layout ``state_dict``.  Every function cites the
reference lines it restates.  It is the checker for ``tests/`` and ``__graft_entry__.smoke()`` and
the timed arm of ``bench.py --impl reference`` / ``cpu_baseline``; the product never calls it.

Parity: unpinned by the reference (no tests there); pinned here against the reference's own
classes imported through ``oracle/ref_shim.py`` -- fixtures in ``tests/golden`` made by
``tests/golden/make_golden.py``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from . import harmonics


@dataclass
class SFNOConfig:
    """The subset of the reference ctor arguments (``sfnonet.py:426-466``) that the hot path uses."""

    num_input_channels: int
    num_output_channels: int
    num_conditional_channels: int = 0
    spatial_shape: tuple = (180, 360)
    embed_dim: int = 256
    num_layers: int = 8
    operator_type: str = "dhconv"
    scale_factor: int = 1
    use_mlp: bool = True
    mlp_ratio: float = 2.0
    activation_function: str = "gelu"
    encoder_layers: int = 1
    pos_embed: bool = True
    big_skip: bool = True
    normalization_layer: str = "instance_norm"
    hard_thresholding_fraction: float = 1.0
    with_time_emb: bool = True
    time_dim_mult: int = 2
    time_rescale: bool = False
    time_scale_shift_before_filter: bool = True
    data_grid: str = "equiangular"
    dropout_mlp: float = 0.0
    drop_path_rate: float = 0.0
    min_time: Optional[float] = 0.0
    max_time: Optional[float] = 5.0

    @property
    def in_chans(self):
        return self.num_input_channels + self.num_conditional_channels

    def model_kwargs(self) -> dict:
        """Keyword arguments for the reference ctor (used by the golden generator)."""
        return dict(
            embed_dim=self.embed_dim, num_layers=self.num_layers, operator_type=self.operator_type,
            scale_factor=self.scale_factor, use_mlp=self.use_mlp, mlp_ratio=self.mlp_ratio,
            activation_function=self.activation_function, encoder_layers=self.encoder_layers,
            pos_embed=self.pos_embed, big_skip=self.big_skip, normalization_layer=self.normalization_layer,
            hard_thresholding_fraction=self.hard_thresholding_fraction, with_time_emb=self.with_time_emb,
            time_dim_mult=self.time_dim_mult, time_rescale=self.time_rescale,
            time_scale_shift_before_filter=self.time_scale_shift_before_filter, data_grid=self.data_grid,
            dropout_mlp=self.dropout_mlp, drop_path_rate=self.drop_path_rate,
        )


ACE_FORECASTER = dict(num_input_channels=34, num_output_channels=34, num_conditional_channels=2)
ACE_INTERPOLATOR = dict(num_input_channels=68, num_output_channels=34, num_conditional_channels=2,
                        dropout_mlp=0.1, drop_path_rate=0.1, min_time=1.0, max_time=5.0)


# ----------------------------------------------------------------------------------------------
# initialisation -- sfnonet.py:725-732,746-754 ; initialization.py:21-73 ; s2convolutions.py:70-71,146,155-156
# ----------------------------------------------------------------------------------------------
def _trunc_normal(shape, std, gen):
    lo = (1.0 + math.erf(-2.0 / std / math.sqrt(2.0))) / 2.0
    hi = (1.0 + math.erf(2.0 / std / math.sqrt(2.0))) / 2.0
    t = torch.empty(shape).uniform_(2 * lo - 1, 2 * hi - 1, generator=gen)
    t.erfinv_().mul_(std * math.sqrt(2.0)).clamp_(-2.0, 2.0)
    return t


def random_state_dict(cfg: SFNOConfig, seed: int = 0, spectral_gain: float = 1.0) -> Dict[str, torch.Tensor]:
    """Random weights with the reference's distributions and ``state_dict`` layout.

    Not bit-identical to the reference's RNG consumption order (the goldens store the reference's
    actual weights); used for full-size parity and benchmarks where only the distribution matters.
    ``spectral_gain`` rescales the dhconv weights (x256 makes the spectral
    branch visible end to end).
    """
    g = torch.Generator().manual_seed(seed)
    C, Cin, Cout = cfg.embed_dim, cfg.in_chans, cfg.num_output_channels
    H, W = cfg.spatial_shape
    L = int((H // cfg.scale_factor) * cfg.hard_thresholding_fraction)
    M = int((W // cfg.scale_factor // 2 + 1) * cfg.hard_thresholding_fraction)
    hid = int(C * cfg.mlp_ratio)
    tdim = C * cfg.time_dim_mult
    sd: Dict[str, torch.Tensor] = {}
    if cfg.pos_embed:
        sd["pos_embed"] = _trunc_normal((1, C, H, W), 0.02, g)
    sd["encoder.0.weight"] = _trunc_normal((C, Cin, 1, 1), 0.02, g)
    sd["encoder.0.bias"] = torch.zeros(C)
    sd["encoder.2.weight"] = _trunc_normal((C, C, 1, 1), 0.02, g)
    if cfg.with_time_emb:
        sd["time_emb_mlp.1.weight"] = _trunc_normal((tdim, C), 0.02, g)
        sd["time_emb_mlp.1.bias"] = torch.zeros(tdim)
        sd["time_emb_mlp.3.weight"] = _trunc_normal((tdim, tdim), 0.02, g)
        sd["time_emb_mlp.3.bias"] = torch.zeros(tdim)
    fc2 = 3 if cfg.dropout_mlp > 0 else 2
    for i in range(cfg.num_layers):
        p = f"blocks.{i}."
        sd[p + "norm0.weight"] = torch.ones(C)
        sd[p + "norm0.bias"] = torch.zeros(C)
        if cfg.with_time_emb:
            sd[p + "time_mlp.1.weight"] = _trunc_normal((2 * C, tdim), 0.02, g)
            sd[p + "time_mlp.1.bias"] = torch.zeros(2 * C)
        wshape = (C, C, L, 2) if cfg.operator_type == "dhconv" else (C, C, L, M, 2)   # s2convolutions.py:139-147
        sd[p + "filter.filter.weight"] = spectral_gain / (C * C) * torch.randn(*wshape, generator=g)
        sd[p + "filter.filter.bias"] = torch.zeros(1, C, 1, 1)
        sd[p + "inner_skip.weight"] = _trunc_normal((C, C, 1, 1), 0.02, g)
        sd[p + "inner_skip.bias"] = torch.zeros(C)
        sd[p + "norm1.weight"] = torch.ones(C)
        sd[p + "norm1.bias"] = torch.zeros(C)
        sd[p + "mlp.fwd.0.weight"] = _trunc_normal((hid, C, 1, 1), 0.02, g)
        sd[p + "mlp.fwd.0.bias"] = torch.zeros(hid)
        sd[p + f"mlp.fwd.{fc2}.weight"] = _trunc_normal((C, hid, 1, 1), 0.02, g)
        sd[p + f"mlp.fwd.{fc2}.bias"] = torch.zeros(C)
    sd["decoder.0.weight"] = _trunc_normal((C, C + cfg.big_skip * Cin, 1, 1), 0.02, g)
    sd["decoder.0.bias"] = torch.zeros(C)
    sd["decoder.2.weight"] = _trunc_normal((Cout, C, 1, 1), 0.02, g)
    return sd


def perturb_affine_and_biases(sd: Dict[str, torch.Tensor], seed: int = 1, scale: float = 0.1):
    """Make biases / norm affine parameters non-trivial so parity tests exercise them."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k, v in sd.items():
        if k.endswith(".bias"):
            out[k] = v + scale * torch.randn(v.shape, generator=g)
        elif "norm" in k and k.endswith(".weight"):
            out[k] = v + scale * torch.randn(v.shape, generator=g)
        else:
            out[k] = v.clone()
    return out


# ----------------------------------------------------------------------------------------------
# the forward pass
# ----------------------------------------------------------------------------------------------
def _act(name):
    return {"gelu": F.gelu, "relu": F.relu, "silu": F.silu}[name]


_TF32_MATMUL = False   # set by SFNOOracle(tf32_matmul=True) around its forward: emulation of TF32 matmuls, see harmonics.round_tf32


def _conv1x1(x, w, b=None):
    if _TF32_MATMUL:
        x, w = harmonics.round_tf32(x), harmonics.round_tf32(w)
    return F.conv2d(x, w, b)


def sinusoidal_pos_emb(t: torch.Tensor, dim: int) -> torch.Tensor:
    """misc.py:21-33."""
    half = dim // 2
    f = math.log(10000) / (half - 1)
    f = torch.exp(torch.arange(half, dtype=t.dtype) * -f)
    e = t[:, None] * f[None, :]
    return torch.cat((e.sin(), e.cos()), dim=-1)


def time_embedding(sd, t: torch.Tensor, dim: int) -> torch.Tensor:
    """misc.py:145-147 (``sinusoidal_embedding == 'true'``): SinusoidalPosEmb -> Linear -> GELU -> Linear."""
    e = sinusoidal_pos_emb(t, dim)
    e = F.linear(e, sd["time_emb_mlp.1.weight"], sd["time_emb_mlp.1.bias"])
    e = F.gelu(e)
    return F.linear(e, sd["time_emb_mlp.3.weight"], sd["time_emb_mlp.3.bias"])


def instance_norm(x, weight, bias, eps=1e-6):
    """nn.InstanceNorm2d(C, eps=1e-6, affine=True, track_running_stats=False) -- sfnonet.py:641-647."""
    mu = x.mean(dim=(-2, -1), keepdim=True)
    var = x.var(dim=(-2, -1), unbiased=False, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * weight[None, :, None, None] + bias[None, :, None, None]


def time_scale_shift(x, t_repr, w, b):
    """sfnonet.py:280-287 with time_mlp = SiLU -> Linear (sfnonet.py:210-213)."""
    e = F.linear(F.silu(t_repr), w, b)
    scale, shift = e[:, :, None, None].chunk(2, dim=1)
    return x * (scale + 1) + shift


def dhconv_contract(x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """contractions.py:159-169: ``einsum('bixy,iox->boxy')`` on complex views; w is real [i,o,l,2]."""
    if _TF32_MATMUL:
        x = torch.view_as_complex(harmonics.round_tf32(torch.view_as_real(x)))
        w = harmonics.round_tf32(w)
    return torch.einsum("bixy,iox->boxy", x, torch.view_as_complex(w.contiguous()))


def diagonal_contract(x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """contractions.py:147-156: ``einsum('bixy,ioxy->boxy')``."""
    return torch.einsum("bixy,ioxy->boxy", x, torch.view_as_complex(w.contiguous()))


class SFNOOracle:
    """Functional restatement of ``SphericalFourierNeuralOperatorNet.forward`` (sfnonet.py:797-841)."""

    def __init__(self, cfg: SFNOConfig, state_dict: Dict[str, torch.Tensor], dtype=torch.float32, tf32_matmul: bool = False):
        """``tf32_matmul=True`` rounds the operands of every 1x1 convolution, Legendre contraction and spectral channel
        contraction to TF32 -- an emulation of the reference under ``torch_matmul_precision: high``
        (``config_utils.py:310-313``) used to put the tf32 mode's error next to the reference's own; parity is always
        judged against the plain fp32 oracle."""
        assert cfg.normalization_layer in ("instance_norm", "none")
        self.tf32_matmul = tf32_matmul
        assert cfg.encoder_layers == 1
        self.cfg = cfg
        self.dtype = dtype
        self.sd = {k: v.detach().to("cpu").to(dtype if v.is_floating_point() else v.dtype) for k, v in state_dict.items()}
        H, W = cfg.spatial_shape
        h, w = H // cfg.scale_factor, W // cfg.scale_factor
        L = int(h * cfg.hard_thresholding_fraction)
        M = int((w // 2 + 1) * cfg.hard_thresholding_fraction)
        # sfnonet.py:551-554 -- tables are fp64-built then cast by ``.float()``
        self.trans_down = harmonics.RealSHT(H, W, lmax=L, mmax=M, grid=cfg.data_grid).to(dtype)
        self.itrans_up = harmonics.InverseRealSHT(H, W, lmax=L, mmax=M, grid=cfg.data_grid).to(dtype)
        self.trans = harmonics.RealSHT(h, w, lmax=L, mmax=M, grid="legendre-gauss").to(dtype)
        self.itrans = harmonics.InverseRealSHT(h, w, lmax=L, mmax=M, grid="legendre-gauss").to(dtype)
        for t in (self.trans_down, self.itrans_up, self.trans, self.itrans):
            t.tf32_matmul = tf32_matmul
        # sfnonet.py:622 -- stochastic depth schedule
        self.dpr = [x.item() for x in torch.linspace(0, cfg.drop_path_rate, cfg.num_layers)]
        self.inference_dropout = False  # dyffusion.py:226-235 turns dropout layers on at sampling time
        self.generator: Optional[torch.Generator] = None
        self.taps: Optional[dict] = None  # set to {} to record intermediates

    # -- s2convolutions.py:158-193 ----------------------------------------------------------------
    def spectral_conv(self, i: int, x: torch.Tensor):
        cfg = self.cfg
        fwd = self.trans_down if i == 0 else self.trans
        inv = self.itrans_up if i == cfg.num_layers - 1 else self.itrans
        scale_residual = (fwd.nlat != inv.nlat) or (fwd.nlon != inv.nlon) or (fwd.grid != inv.grid)
        residual = x
        X = fwd(x)
        if scale_residual:
            residual = inv(X)
        wname = f"blocks.{i}.filter.filter.weight"
        if cfg.operator_type == "dhconv":
            Y = dhconv_contract(X, self.sd[wname])
        elif cfg.operator_type == "diagonal":
            Y = diagonal_contract(X, self.sd[wname])
        else:
            raise ValueError(cfg.operator_type)
        if self.taps is not None:
            self.taps[f"blocks.{i}.sht"] = X
            self.taps[f"blocks.{i}.contract"] = Y
        y = inv(Y) + self.sd[f"blocks.{i}.filter.filter.bias"]
        return y, residual

    def _dropout(self, x, p):
        if p > 0.0 and self.inference_dropout:
            mask = (torch.rand(x.shape, generator=self.generator, dtype=x.dtype) >= p).to(x.dtype)
            return x * mask / (1.0 - p)
        return x

    # -- layers.py:73-80 -----------------------------------------------------------------------------
    def mlp(self, i: int, x: torch.Tensor):
        cfg, sd = self.cfg, self.sd
        fc2 = 3 if cfg.dropout_mlp > 0 else 2
        p = f"blocks.{i}.mlp.fwd."
        x = _conv1x1(x, sd[p + "0.weight"], sd[p + "0.bias"])
        x = _act(cfg.activation_function)(x)
        x = self._dropout(x, cfg.dropout_mlp)
        x = _conv1x1(x, sd[p + f"{fc2}.weight"], sd[p + f"{fc2}.bias"])
        return self._dropout(x, cfg.dropout_mlp)

    # -- sfnonet.py:289-337 ------------------------------------------------------------------------
    def block(self, i: int, x: torch.Tensor, t_repr):
        cfg, sd = self.cfg, self.sd
        p = f"blocks.{i}."
        norm = instance_norm if cfg.normalization_layer == "instance_norm" else (lambda v, *_: v)
        has_time = cfg.with_time_emb
        x_norm = norm(x, sd.get(p + "norm0.weight"), sd.get(p + "norm0.bias"))
        if has_time and cfg.time_scale_shift_before_filter:
            x_norm = time_scale_shift(x_norm, t_repr, sd[p + "time_mlp.1.weight"], sd[p + "time_mlp.1.bias"])
        y, residual = self.spectral_conv(i, x_norm)
        y = y + _conv1x1(residual, sd[p + "inner_skip.weight"], sd[p + "inner_skip.bias"])
        y = _act(cfg.activation_function)(y)
        if self.taps is not None:
            self.taps[p + "x_norm"] = x_norm
            self.taps[p + "residual"] = residual
            self.taps[p + "act"] = y
        y = norm(y, sd.get(p + "norm1.weight"), sd.get(p + "norm1.bias"))
        if has_time and not cfg.time_scale_shift_before_filter:
            y = time_scale_shift(y, t_repr, sd[p + "time_mlp.1.weight"], sd[p + "time_mlp.1.bias"])
        if cfg.use_mlp:
            y = self.mlp(i, y)
        # drop_path.py:5-22 (per-sample Bernoulli keep); DropPath is in ``all_dropout_layers``
        # (utils.py:683), so inference dropout switches it on together with nn.Dropout.
        dp = self.dpr[i]
        if dp > 0.0 and self.inference_dropout:
            keep = 1.0 - dp
            mask = torch.floor(keep + torch.rand((y.shape[0], 1, 1, 1), generator=self.generator, dtype=y.dtype))
            y = y / keep * mask
        out = y + residual  # outer skip = identity on the *filter's* residual (sfnonet.py:335)
        if self.taps is not None:
            self.taps[p + "out"] = out
        return out

    @torch.inference_mode()
    def forward(self, inputs, time=None, condition=None, static_condition=None, return_time_emb=False):
        global _TF32_MATMUL
        prev, _TF32_MATMUL = _TF32_MATMUL, self.tf32_matmul
        try:
            return self._forward(inputs, time, condition, static_condition, return_time_emb)
        finally:
            _TF32_MATMUL = prev

    def _forward(self, inputs, time=None, condition=None, static_condition=None, return_time_emb=False):
        cfg, sd = self.cfg, self.sd
        dt = self.dtype
        # _base_model.py:166-192
        if cfg.num_conditional_channels > 0:
            if condition is None and static_condition is None:
                raise ValueError("condition and static_condition are both None")
            if condition is not None and static_condition is not None:
                condition = torch.cat((condition, static_condition), dim=1)
            elif condition is None:
                condition = static_condition
            x = torch.cat((inputs, condition), dim=1)
        else:
            assert condition is None and static_condition is None
            x = inputs
        x = x.to(dt)
        residual_big = x
        act = _act(cfg.activation_function)
        # sfnonet.py:610-618
        x = _conv1x1(x, sd["encoder.0.weight"], sd["encoder.0.bias"])
        x = act(x)
        x = _conv1x1(x, sd["encoder.2.weight"])
        if cfg.pos_embed:
            x = x + sd["pos_embed"]
        if self.taps is not None:
            self.taps["encoded"] = x
        # sfnonet.py:775-795
        t_repr = None
        if cfg.with_time_emb:
            assert cfg.min_time is not None and cfg.max_time is not None
            time = time.to(dt)
            assert (cfg.min_time <= time).all() and (time <= cfg.max_time).all(), f"time out of range: {time}"
            if cfg.time_rescale:
                time = time * (1000.0 / (cfg.max_time - cfg.min_time)) + (-cfg.min_time)
            t_repr = time_embedding(sd, time, cfg.embed_dim)
        for i in range(cfg.num_layers):
            x = self.block(i, x, t_repr)
        if cfg.big_skip:
            x = torch.cat((x, residual_big), dim=1)
        x = _conv1x1(x, sd["decoder.0.weight"], sd["decoder.0.bias"])
        x = act(x)
        x = _conv1x1(x, sd["decoder.2.weight"])
        if return_time_emb:
            return x, t_repr
        return x

    __call__ = forward


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    """Relative L2 error of ``a`` against reference ``b`` (the parity metric of north_star)."""
    wide = torch.complex128 if (a.is_complex() or b.is_complex()) else torch.float64
    a = a.detach().cpu().to(wide)
    b = b.detach().cpu().to(wide)
    return (torch.linalg.vector_norm(a - b) / torch.linalg.vector_norm(b).clamp_min(1e-300)).item()
