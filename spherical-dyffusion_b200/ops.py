"""PyTorch custom-op layer over the C ABI (``torch.ops.sfno_b200.*``).

Every op enqueues hand-written sm_100a kernels of ``libsfno_b200.so`` on PyTorch's current CUDA stream and allocates
nothing but its result; plan / network handles cross the dispatcher as plain integers (the address of the opaque C
handle).  Each op has a fake (meta) implementation for shape inference, so modules built on them trace under
``torch.compile`` / ``torch.export`` as opaque calls.  There is no CPU implementation: a CPU tensor raises.

=====================================  ==================================================================================
op                                     replaces (reference file:line)
=====================================  ==================================================================================
``sfno_b200::sht_forward``             ``torch_harmonics.RealSHT.forward`` called at ``s2convolutions.py:165``
``sfno_b200::sht_inverse``             ``torch_harmonics.InverseRealSHT.forward`` called at ``s2convolutions.py:168,186``
``sfno_b200::spectral_contract``       ``_contract_dhconv`` / ``_contract_diagonal`` ``contractions.py:147-169``
``sfno_b200::instance_norm``           ``nn.InstanceNorm2d`` ``sfnonet.py:641-647`` + ``time_scale_shift`` ``:280-287``
``sfno_b200::conv1x1``                 ``nn.Conv2d(.., 1)`` ``sfnonet.py:239,614-617,739-742``, ``layers.py:73-75``
``sfno_b200::conv1x1_ex``              the same + ``nn.Dropout`` (``layers.py:76-80``) on a selectable engine (fp32 / tf32 / bf16)
``sfno_b200::spectral_conv``           ``SpectralConvS2.forward`` ``s2convolutions.py:158-193`` as one fused call (bf16 / tf32 / fp32)
``sfno_b200::net_forward``             ``SphericalFourierNeuralOperatorNet.forward`` ``sfnonet.py:797-841``
``sfno_b200::cold_update``             ``x_s + (x_interpolated_s_next - x_interpolated_s)`` ``src/diffusion/dyffusion.py:519``
``sfno_b200::*_adjoint / *_backward``  autograd formulas of the ops above (registered with ``torch.library.register_autograd``)
=====================================  ==================================================================================
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._util import require_cuda_f32, stream_ptr, workspace

_ACT = {"none": 0, "gelu": 1}


def _handle(h: int) -> ctypes.c_void_p:
    return ctypes.c_void_p(int(h))


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


# ---- spherical harmonic transforms ------------------------------------------------------------------------------------
@torch.library.custom_op("sfno_b200::sht_forward", mutates_args=())
def sht_forward(plan: int, x: torch.Tensor, lmax: int, mmax: int) -> torch.Tensor:
    """x fp32 [..., nlat, nlon] -> coefficients fp32 [..., lmax, mmax, 2] (view_as_complex gives the reference layout)."""
    xf = require_cuda_f32(x, "x")
    lead = xf.shape[:-2]
    fields = int(torch.Size(lead).numel()) if len(lead) else 1
    out = torch.empty(*lead, lmax, mmax, 2, dtype=torch.float32, device=xf.device)
    if fields == 0:
        return out
    L = _lib.lib()
    with torch.cuda.device(xf.device):
        ws = workspace(xf.device, L.sfno_sht_workspace_bytes(_handle(plan), fields), "sht")
        _lib.check(L.sfno_sht_forward(_handle(plan), xf.data_ptr(), out.data_ptr(), fields, ws.data_ptr(), ws.numel(),
                                      stream_ptr(xf.device)), "sfno_sht_forward")
    return out


@sht_forward.register_fake
def _(plan, x, lmax, mmax):
    return x.new_empty(*x.shape[:-2], lmax, mmax, 2, dtype=torch.float32)


@torch.library.custom_op("sfno_b200::sht_inverse", mutates_args=())
def sht_inverse(plan: int, coeffs: torch.Tensor, nlat: int, nlon: int) -> torch.Tensor:
    """coefficients fp32 [..., lmax, mmax, 2] -> x fp32 [..., nlat, nlon]."""
    if not coeffs.is_cuda:
        raise RuntimeError("coefficients must be a CUDA tensor: the B200 path has no CPU fallback")
    xr = coeffs.to(torch.float32).contiguous()
    lead = xr.shape[:-3]
    fields = int(torch.Size(lead).numel()) if len(lead) else 1
    out = torch.empty(*lead, nlat, nlon, dtype=torch.float32, device=xr.device)
    if fields == 0:
        return out
    L = _lib.lib()
    with torch.cuda.device(xr.device):
        ws = workspace(xr.device, L.sfno_sht_workspace_bytes(_handle(plan), fields), "sht")
        _lib.check(L.sfno_sht_inverse(_handle(plan), xr.data_ptr(), out.data_ptr(), fields, ws.data_ptr(), ws.numel(),
                                      stream_ptr(xr.device)), "sfno_sht_inverse")
    return out


@sht_inverse.register_fake
def _(plan, coeffs, nlat, nlon):
    return coeffs.new_empty(*coeffs.shape[:-3], nlat, nlon, dtype=torch.float32)


# ---- spectral channel contraction -------------------------------------------------------------------------------------
@torch.library.custom_op("sfno_b200::spectral_contract", mutates_args=())
def spectral_contract(operator_type: int, x: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    """x fp32 [B, Cin, L, M, 2]; weight fp32 [Cin, Cout, L, 2] (dhconv) or [Cin, Cout, L, M, 2] (diagonal) -> [B, Cout, L, M, 2]."""
    xf = require_cuda_f32(x, "x")
    w = require_cuda_f32(weight, "weight")
    B, Ci, Lm, Mm = (int(v) for v in xf.shape[:4])
    Co = int(w.shape[1])
    out = torch.empty(B, Co, Lm, Mm, 2, dtype=torch.float32, device=xf.device)
    if out.numel() == 0:
        return out
    with torch.cuda.device(xf.device):
        _lib.check(_lib.lib().sfno_spectral_contract(int(operator_type), xf.data_ptr(), w.data_ptr(), out.data_ptr(), B, Ci, Co, Lm, Mm,
                                                     stream_ptr(xf.device)), "sfno_spectral_contract")
    return out


@spectral_contract.register_fake
def _(operator_type, x, weight):
    return x.new_empty(x.shape[0], weight.shape[1], x.shape[2], x.shape[3], 2, dtype=torch.float32)


# ---- InstanceNorm (+ time scale / shift) ------------------------------------------------------------------------------
@torch.library.custom_op("sfno_b200::instance_norm", mutates_args=())
def instance_norm(x: torch.Tensor, gamma: Optional[torch.Tensor], beta: Optional[torch.Tensor], scale: Optional[torch.Tensor],
                  shift: Optional[torch.Tensor], eps: float) -> torch.Tensor:
    """x fp32 [B, C, H, W]; gamma/beta [C]; scale/shift [B, C] (``x * (scale + 1) + shift`` after the norm) -> same shape."""
    xf = require_cuda_f32(x, "x")
    B, C = int(xf.shape[0]), int(xf.shape[1])
    hw = int(xf.numel() // max(B * C, 1))
    y = torch.empty_like(xf)
    if xf.numel() == 0:
        return y
    opt = [None if t is None else require_cuda_f32(t, "parameter") for t in (gamma, beta, scale, shift)]
    L = _lib.lib()
    with torch.cuda.device(xf.device):
        ws = workspace(xf.device, L.sfno_instance_norm_workspace_bytes(B, C), "norm")
        _lib.check(L.sfno_instance_norm(xf.data_ptr(), y.data_ptr(), _ptr(opt[0]), _ptr(opt[1]), _ptr(opt[2]), _ptr(opt[3]), B, C, hw,
                                        float(eps), ws.data_ptr(), ws.numel(), stream_ptr(xf.device)), "sfno_instance_norm")
    return y


@instance_norm.register_fake
def _(x, gamma, beta, scale, shift, eps):
    return torch.empty_like(x, dtype=torch.float32)


# ---- 1x1 convolution ---------------------------------------------------------------------------------------------------
@torch.library.custom_op("sfno_b200::conv1x1", mutates_args=())
def conv1x1(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], residual: Optional[torch.Tensor], act: int) -> torch.Tensor:
    """x fp32 [B, Cin, H, W]; weight [Cout, Cin(,1,1)]; bias [Cout]; residual [B, Cout, H, W]; act 0 none / 1 GELU."""
    xf = require_cuda_f32(x, "x")
    w = require_cuda_f32(weight, "weight").reshape(weight.shape[0], -1)
    B, Ci = int(xf.shape[0]), int(xf.shape[1])
    Co = int(w.shape[0])
    hw = int(xf.numel() // max(B * Ci, 1))
    y = torch.empty(B, Co, *xf.shape[2:], dtype=torch.float32, device=xf.device)
    if y.numel() == 0:
        return y
    b = None if bias is None else require_cuda_f32(bias, "bias")
    r = None if residual is None else require_cuda_f32(residual, "residual")
    with torch.cuda.device(xf.device):
        _lib.check(_lib.lib().sfno_conv1x1(xf.data_ptr(), w.data_ptr(), _ptr(b), _ptr(r), y.data_ptr(), B, Ci, Co, hw, int(act),
                                           stream_ptr(xf.device)), "sfno_conv1x1")
    return y


@conv1x1.register_fake
def _(x, weight, bias, residual, act):
    return x.new_empty(x.shape[0], weight.shape[0], *x.shape[2:], dtype=torch.float32)


# ---- fused spectral convolution (SpectralConvS2.forward) ----------------------------------------------------------------
@torch.library.custom_op("sfno_b200::spectral_conv", mutates_args=())
def spectral_conv(plan_fwd: int, plan_inv: int, weight: int, x: torch.Tensor, cout: int, nlat_out: int, nlon_out: int,
                  want_residual: bool) -> List[torch.Tensor]:
    """x fp32 [B, Cin, H, W] -> [y [B, Cout, H', W'], residual [B, Cin, H', W'] (empty unless want_residual)] in ONE library
    call (SHT -> contraction -> inverse SHT + bias) on the engine of the plans' precision; ``weight`` is a
    ``sfno_spectral_weight`` handle holding the packed filter weight and bias."""
    xf = require_cuda_f32(x, "x")
    B, cin = int(xf.shape[0]), int(xf.shape[1])
    y = torch.empty(B, cout, nlat_out, nlon_out, dtype=torch.float32, device=xf.device)
    res = torch.empty((B, cin, nlat_out, nlon_out) if want_residual else (0,), dtype=torch.float32, device=xf.device)
    if B == 0:
        return [y, res]
    L = _lib.lib()
    with torch.cuda.device(xf.device):
        ws = workspace(xf.device, L.sfno_spectral_conv_workspace_bytes(_handle(plan_fwd), _handle(plan_inv), _handle(weight), B), "spectral_conv")
        _lib.check(L.sfno_spectral_conv(_handle(plan_fwd), _handle(plan_inv), _handle(weight), xf.data_ptr(), y.data_ptr(),
                                        res.data_ptr() if want_residual else None, B, ws.data_ptr(), ws.numel(), stream_ptr(xf.device)),
                   "sfno_spectral_conv")
    return [y, res]


@spectral_conv.register_fake
def _(plan_fwd, plan_inv, weight, x, cout, nlat_out, nlon_out, want_residual):
    y = x.new_empty(x.shape[0], cout, nlat_out, nlon_out, dtype=torch.float32)
    res = x.new_empty((x.shape[0], x.shape[1], nlat_out, nlon_out) if want_residual else (0,), dtype=torch.float32)
    return [y, res]


# ---- 1x1 convolution with the full fused epilogue and a selectable engine -----------------------------------------------------
@torch.library.custom_op("sfno_b200::conv1x1_ex", mutates_args=())
def conv1x1_ex(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], residual: Optional[torch.Tensor], act: int,
               dropout_p: float, seed: int, offset: int, precision: int) -> torch.Tensor:
    """``conv1x1`` + dropout (Philox stream (seed, offset); p = 0: off) on the engine selected by ``precision``
    (``_lib.SFNO_PREC``: fp32 CUDA cores | bf16 / tf32 tensor cores); fp32 tensors in and out."""
    xf = require_cuda_f32(x, "x")
    w = require_cuda_f32(weight, "weight").reshape(weight.shape[0], -1)
    B, Ci, Co = int(xf.shape[0]), int(xf.shape[1]), int(w.shape[0])
    hw = int(xf.numel() // max(B * Ci, 1))
    y = torch.empty(B, Co, *xf.shape[2:], dtype=torch.float32, device=xf.device)
    if y.numel() == 0:
        return y
    b = None if bias is None else require_cuda_f32(bias, "bias")
    r = None if residual is None else require_cuda_f32(residual, "residual")
    L = _lib.lib()
    with torch.cuda.device(xf.device):
        ws = workspace(xf.device, L.sfno_conv1x1_ex_workspace_bytes(B, Ci, Co, hw, int(precision)), "conv1x1")
        _lib.check(L.sfno_conv1x1_ex(xf.data_ptr(), w.data_ptr(), _ptr(b), _ptr(r), y.data_ptr(), B, Ci, Co, hw, int(act), float(dropout_p),
                                     int(seed), int(offset), int(precision), ws.data_ptr(), ws.numel(), stream_ptr(xf.device)),
                   "sfno_conv1x1_ex")
    return y


@conv1x1_ex.register_fake
def _(x, weight, bias, residual, act, dropout_p, seed, offset, precision):
    return x.new_empty(x.shape[0], weight.shape[0], *x.shape[2:], dtype=torch.float32)


# ---- whole network -----------------------------------------------------------------------------------------------------
@torch.library.custom_op("sfno_b200::net_forward", mutates_args=("rng_state",))
def net_forward(net: int, parts: Sequence[torch.Tensor], time: Optional[torch.Tensor], out_channels: int, dropout: bool,
                rng_state: Optional[torch.Tensor]) -> torch.Tensor:
    """parts: the tensors the reference concatenates on dim 1 (inputs, condition, static condition), fp32 [B, c_i, H, W];
    time fp32 [B] or None -> [B, out_channels, H, W] fp32.  `net` must have its parameters set (``sfno_net_set_param``).
    rng_state: int64 [2] = {seed, offset} on the device (the Philox key of the dropout masks; advanced by the call, so a
    captured CUDA graph draws fresh masks on every replay); required iff ``dropout``."""
    x = parts[0]
    B = int(x.shape[0])
    out = torch.empty(B, out_channels, *x.shape[2:], dtype=torch.float32, device=x.device)
    if B == 0:
        return out
    if dropout and (rng_state is None or rng_state.dtype != torch.int64 or rng_state.numel() != 2 or rng_state.device != x.device):
        raise RuntimeError("net_forward with dropout needs rng_state: an int64 [2] tensor on the input's device")
    L = _lib.lib()
    with torch.cuda.device(x.device):
        ws = workspace(x.device, L.sfno_net_workspace_bytes(_handle(net), B), f"net{int(net)}")
        ptrs = (ctypes.c_void_p * len(parts))(*[t.data_ptr() for t in parts])
        chans = (ctypes.c_int * len(parts))(*[int(t.shape[1]) for t in parts])
        _lib.check(L.sfno_net_forward_parts_rng(_handle(net), ptrs, chans, len(parts), _ptr(time), out.data_ptr(), B, int(dropout),
                                                _ptr(rng_state), ws.data_ptr(), ws.numel(), stream_ptr(x.device)),
                   "sfno_net_forward_parts_rng")
    return out


@net_forward.register_fake
def _(net, parts, time, out_channels, dropout, rng_state):
    x = parts[0]
    return x.new_empty(x.shape[0], out_channels, *x.shape[2:], dtype=torch.float32)


# ---- sampler glue ---------------------------------------------------------------------------------------------------------
@torch.library.custom_op("sfno_b200::cold_update", mutates_args=())
def cold_update(x_s: torch.Tensor, x_next: torch.Tensor, x_cur: torch.Tensor) -> torch.Tensor:
    """``x_s + (x_next - x_cur)`` in one pass: the cold-sampling update of ``dyffusion.py:519``."""
    a, b, c = (require_cuda_f32(t, "x") for t in (x_s, x_next, x_cur))
    if not (a.shape == b.shape == c.shape):
        raise RuntimeError(f"cold_update: shapes differ: {tuple(a.shape)}, {tuple(b.shape)}, {tuple(c.shape)}")
    out = torch.empty_like(a)
    if a.numel():
        with torch.cuda.device(a.device):
            _lib.check(_lib.lib().sfno_cold_update(a.data_ptr(), b.data_ptr(), c.data_ptr(), out.data_ptr(), a.numel(),
                                                   stream_ptr(a.device)), "sfno_cold_update")
    return out


@cold_update.register_fake
def _(x_s, x_next, x_cur):
    return torch.empty_like(x_s, dtype=torch.float32)


# =========================================================================================================================
# Backward pass: autograd formulas of the custom ops.  Every formula is itself a custom op over the C ABI --
# the adjoint transforms run the opposite transform's GEMM ops on transposed tables (on the plan's engine), the weight
# gradient of the 1x1 convolution is a split-K GEMM, the contraction and norm gradients are dedicated kernels.
# =========================================================================================================================
@torch.library.custom_op("sfno_b200::sht_forward_adjoint", mutates_args=())
def sht_forward_adjoint(plan: int, grad_coeffs: torch.Tensor, nlat: int, nlon: int) -> torch.Tensor:
    """Transposed RealSHT: gradient w.r.t. x [..., nlat, nlon] from the gradient w.r.t. the (re, im) pairs [..., lmax, mmax, 2]."""
    g = require_cuda_f32(grad_coeffs, "grad_coeffs")
    lead = g.shape[:-3]
    fields = int(torch.Size(lead).numel()) if len(lead) else 1
    out = torch.empty(*lead, nlat, nlon, dtype=torch.float32, device=g.device)
    if fields == 0:
        return out
    L = _lib.lib()
    with torch.cuda.device(g.device):
        ws = workspace(g.device, L.sfno_sht_workspace_bytes(_handle(plan), fields), "sht")
        _lib.check(L.sfno_sht_forward_adjoint(_handle(plan), g.data_ptr(), out.data_ptr(), fields, ws.data_ptr(), ws.numel(),
                                              stream_ptr(g.device)), "sfno_sht_forward_adjoint")
    return out


@sht_forward_adjoint.register_fake
def _(plan, grad_coeffs, nlat, nlon):
    return grad_coeffs.new_empty(*grad_coeffs.shape[:-3], nlat, nlon, dtype=torch.float32)


@torch.library.custom_op("sfno_b200::sht_inverse_adjoint", mutates_args=())
def sht_inverse_adjoint(plan: int, grad_x: torch.Tensor, lmax: int, mmax: int) -> torch.Tensor:
    """Transposed InverseRealSHT: gradient w.r.t. the coefficient pairs [..., lmax, mmax, 2] from the gradient w.r.t. x."""
    g = require_cuda_f32(grad_x, "grad_x")
    lead = g.shape[:-2]
    fields = int(torch.Size(lead).numel()) if len(lead) else 1
    out = torch.empty(*lead, lmax, mmax, 2, dtype=torch.float32, device=g.device)
    if fields == 0:
        return out
    L = _lib.lib()
    with torch.cuda.device(g.device):
        ws = workspace(g.device, L.sfno_sht_workspace_bytes(_handle(plan), fields), "sht")
        _lib.check(L.sfno_sht_inverse_adjoint(_handle(plan), g.data_ptr(), out.data_ptr(), fields, ws.data_ptr(), ws.numel(),
                                              stream_ptr(g.device)), "sfno_sht_inverse_adjoint")
    return out


@sht_inverse_adjoint.register_fake
def _(plan, grad_x, lmax, mmax):
    return grad_x.new_empty(*grad_x.shape[:-2], lmax, mmax, 2, dtype=torch.float32)


@torch.library.custom_op("sfno_b200::spectral_contract_backward", mutates_args=())
def spectral_contract_backward(operator_type: int, x: torch.Tensor, weight: torch.Tensor, grad_out: torch.Tensor) -> List[torch.Tensor]:
    """[grad_x, grad_weight] of ``spectral_contract`` (complex as (re, im) pairs): sum_o g conj(w), sum_{b(,m)} conj(x) g."""
    xf, w, g = require_cuda_f32(x, "x"), require_cuda_f32(weight, "weight"), require_cuda_f32(grad_out, "grad_out")
    B, Ci, Lm, Mm = (int(v) for v in xf.shape[:4])
    Co = int(w.shape[1])
    gx, gw = torch.empty_like(xf), torch.empty_like(w)
    if xf.numel() and g.numel():
        with torch.cuda.device(xf.device):
            _lib.check(_lib.lib().sfno_spectral_contract_backward(int(operator_type), xf.data_ptr(), w.data_ptr(), g.data_ptr(), gx.data_ptr(),
                                                                  gw.data_ptr(), B, Ci, Co, Lm, Mm, stream_ptr(xf.device)),
                       "sfno_spectral_contract_backward")
    return [gx, gw]


@spectral_contract_backward.register_fake
def _(operator_type, x, weight, grad_out):
    return [torch.empty_like(x, dtype=torch.float32), torch.empty_like(weight, dtype=torch.float32)]


@torch.library.custom_op("sfno_b200::conv1x1_weight_grad", mutates_args=())
def conv1x1_weight_grad(x: torch.Tensor, grad_y: torch.Tensor) -> List[torch.Tensor]:
    """[grad_weight [Cout, Cin], grad_bias [Cout]] of a 1x1 convolution from x [B, Cin, H, W] and grad_y [B, Cout, H, W]."""
    xf, g = require_cuda_f32(x, "x"), require_cuda_f32(grad_y, "grad_y")
    B, Ci, Co = int(xf.shape[0]), int(xf.shape[1]), int(g.shape[1])
    hw = int(xf.numel() // max(B * Ci, 1))
    gw = torch.zeros(Co, Ci, dtype=torch.float32, device=xf.device)
    gb = torch.zeros(Co, dtype=torch.float32, device=xf.device)
    if xf.numel() and g.numel():
        L = _lib.lib()
        with torch.cuda.device(xf.device):
            ws = workspace(xf.device, L.sfno_conv1x1_weight_grad_workspace_bytes(B, Ci, Co, hw), "wgrad")
            _lib.check(L.sfno_conv1x1_weight_grad(xf.data_ptr(), g.data_ptr(), gw.data_ptr(), gb.data_ptr(), B, Ci, Co, hw, ws.data_ptr(),
                                                  ws.numel(), stream_ptr(xf.device)), "sfno_conv1x1_weight_grad")
    return [gw, gb]


@conv1x1_weight_grad.register_fake
def _(x, grad_y):
    return [x.new_empty(grad_y.shape[1], x.shape[1], dtype=torch.float32), x.new_empty(grad_y.shape[1], dtype=torch.float32)]


@torch.library.custom_op("sfno_b200::instance_norm_backward", mutates_args=())
def instance_norm_backward(x: torch.Tensor, grad_out: torch.Tensor, affine_a: Optional[torch.Tensor], eps: float) -> List[torch.Tensor]:
    """[grad_x, grad_a [B, C] = sum grad_out xhat, grad_d [B, C] = sum grad_out] for y = xhat * affine_a + d per (b, c) plane."""
    xf, g = require_cuda_f32(x, "x"), require_cuda_f32(grad_out, "grad_out")
    B, C = int(xf.shape[0]), int(xf.shape[1])
    hw = int(xf.numel() // max(B * C, 1))
    gx = torch.empty_like(xf)
    da = torch.zeros(B, C, dtype=torch.float32, device=xf.device)
    dd = torch.zeros(B, C, dtype=torch.float32, device=xf.device)
    a = None if affine_a is None else require_cuda_f32(affine_a, "affine_a")
    if xf.numel():
        with torch.cuda.device(xf.device):
            _lib.check(_lib.lib().sfno_instance_norm_backward(xf.data_ptr(), g.data_ptr(), _ptr(a), gx.data_ptr(), da.data_ptr(), dd.data_ptr(),
                                                              B, C, hw, float(eps), stream_ptr(xf.device)), "sfno_instance_norm_backward")
    return [gx, da, dd]


@instance_norm_backward.register_fake
def _(x, grad_out, affine_a, eps):
    return [torch.empty_like(x, dtype=torch.float32), x.new_empty(x.shape[0], x.shape[1], dtype=torch.float32),
            x.new_empty(x.shape[0], x.shape[1], dtype=torch.float32)]


@torch.library.custom_op("sfno_b200::conv1x1_backward", mutates_args=())
def conv1x1_backward(x: torch.Tensor, grad_y: torch.Tensor, weight: torch.Tensor, need_x: bool, need_weight: bool, need_bias: bool,
                     precision: int) -> List[torch.Tensor]:
    """[grad_x, grad_weight [Cout, Cin], grad_bias [Cout]] (each empty when not needed) of a 1x1 convolution on the engine of
    ``precision``: one staging of grad_y serves the data gradient (forward op, transposed weight) and the weight gradient
    (split-K GEMM over the pixels, tensor cores with fp32 partial sums in bf16 / tf32)."""
    xf, g = require_cuda_f32(x, "x"), require_cuda_f32(grad_y, "grad_y")
    w = require_cuda_f32(weight, "weight").reshape(weight.shape[0], -1)
    B, Ci, Co = int(xf.shape[0]), int(xf.shape[1]), int(g.shape[1])
    hw = int(xf.numel() // max(B * Ci, 1))
    dev = xf.device
    e = lambda: torch.empty(0, dtype=torch.float32, device=dev)
    gx = torch.empty_like(xf) if need_x else e()
    gw = torch.empty(Co, Ci, dtype=torch.float32, device=dev) if need_weight else e()
    gb = torch.empty(Co, dtype=torch.float32, device=dev) if need_bias else e()
    if xf.numel() == 0 or g.numel() == 0:
        return [gx.zero_(), gw.zero_(), gb.zero_()]
    L = _lib.lib()
    with torch.cuda.device(dev):
        ws = workspace(dev, L.sfno_conv1x1_backward_workspace_bytes(B, Ci, Co, hw, int(precision)), "conv1x1_bwd")
        _lib.check(L.sfno_conv1x1_backward(xf.data_ptr(), g.data_ptr(), w.data_ptr(), gx.data_ptr() if need_x else None,
                                           gw.data_ptr() if need_weight else None, gb.data_ptr() if need_bias else None, B, Ci, Co, hw,
                                           int(precision), ws.data_ptr(), ws.numel(), stream_ptr(dev)), "sfno_conv1x1_backward")
    return [gx, gw, gb]


@conv1x1_backward.register_fake
def _(x, grad_y, weight, need_x, need_weight, need_bias, precision):
    e = lambda: x.new_empty(0, dtype=torch.float32)
    return [torch.empty_like(x, dtype=torch.float32) if need_x else e(),
            x.new_empty(grad_y.shape[1], x.shape[1], dtype=torch.float32) if need_weight else e(),
            x.new_empty(grad_y.shape[1], dtype=torch.float32) if need_bias else e()]


@torch.library.custom_op("sfno_b200::spectral_conv_diff", mutates_args=())
def spectral_conv_diff(plan_fwd: int, plan_inv: int, handle: int, weight: torch.Tensor, bias: Optional[torch.Tensor], x: torch.Tensor,
                       nlat_out: int, nlon_out: int, want_residual: bool) -> Tuple[torch.Tensor, torch.Tensor]:
    """``spectral_conv`` with the filter weight / bias as tensor inputs so that autograd reaches them (``handle`` must hold
    their packed form: ``SpectralConvS2._weight_handle``); the backward is ``spectral_conv_backward``."""
    y, res = spectral_conv(plan_fwd, plan_inv, handle, x, int(weight.shape[1]), nlat_out, nlon_out, want_residual)
    return y, res


@spectral_conv_diff.register_fake
def _(plan_fwd, plan_inv, handle, weight, bias, x, nlat_out, nlon_out, want_residual):
    y = x.new_empty(x.shape[0], weight.shape[1], nlat_out, nlon_out, dtype=torch.float32)
    res = x.new_empty((x.shape[0], x.shape[1], nlat_out, nlon_out) if want_residual else (0,), dtype=torch.float32)
    return y, res


@torch.library.custom_op("sfno_b200::spectral_conv_backward", mutates_args=())
def spectral_conv_backward(plan_fwd: int, plan_inv: int, handle: int, weight: torch.Tensor, x: torch.Tensor, grad_y: torch.Tensor,
                           grad_residual: Optional[torch.Tensor], need_x: bool, need_weight: bool, need_bias: bool) -> List[torch.Tensor]:
    """[grad_x, grad_weight, grad_bias] of the fused spectral convolution (each empty when not needed): transposed transforms and
    the conjugate-transposed contraction on the plans' engine, the weight gradient as an fp32 per-degree GEMM."""
    xf, w, gy = require_cuda_f32(x, "x"), require_cuda_f32(weight, "weight"), require_cuda_f32(grad_y, "grad_y")
    gr = None if grad_residual is None else require_cuda_f32(grad_residual, "grad_residual")
    B = int(xf.shape[0])
    dev = xf.device
    gx = torch.empty_like(xf) if need_x else torch.empty(0, dtype=torch.float32, device=dev)
    gw = torch.empty_like(w) if need_weight else torch.empty(0, dtype=torch.float32, device=dev)
    gb = torch.empty(int(w.shape[1]), dtype=torch.float32, device=dev) if need_bias else torch.empty(0, dtype=torch.float32, device=dev)
    if B == 0:
        return [gx, gw.zero_(), gb.zero_()]
    L = _lib.lib()
    with torch.cuda.device(dev):
        ws = workspace(dev, L.sfno_spectral_conv_backward_workspace_bytes(_handle(plan_fwd), _handle(plan_inv), _handle(handle), B),
                       "spectral_conv_bwd")
        _lib.check(L.sfno_spectral_conv_backward(_handle(plan_fwd), _handle(plan_inv), _handle(handle), w.data_ptr(), xf.data_ptr(),
                                                 gy.data_ptr(), _ptr(gr), gx.data_ptr() if need_x else None,
                                                 gw.data_ptr() if need_weight else None, gb.data_ptr() if need_bias else None, B,
                                                 ws.data_ptr(), ws.numel(), stream_ptr(dev)), "sfno_spectral_conv_backward")
    return [gx, gw, gb]


@spectral_conv_backward.register_fake
def _(plan_fwd, plan_inv, handle, weight, x, grad_y, grad_residual, need_x, need_weight, need_bias):
    e = lambda: x.new_empty(0, dtype=torch.float32)
    return [torch.empty_like(x, dtype=torch.float32) if need_x else e(), torch.empty_like(weight, dtype=torch.float32) if need_weight else e(),
            x.new_empty(weight.shape[1], dtype=torch.float32) if need_bias else e()]


# ---- autograd registrations -------------------------------------------------------------------------------------------------
def _spec_setup(ctx, inputs, output):
    plan_fwd, plan_inv, handle, weight, bias, x, _, _, want_residual = inputs
    ctx.args = (int(plan_fwd), int(plan_inv), int(handle), bool(want_residual), None if bias is None else tuple(bias.shape))
    ctx.save_for_backward(weight, x)


def _spec_backward(ctx, grad_y, grad_res):
    weight, x = ctx.saved_tensors
    plan_fwd, plan_inv, handle, want_residual, bias_shape = ctx.args
    need_x, need_w = ctx.needs_input_grad[5], ctx.needs_input_grad[3]
    need_b = bias_shape is not None and ctx.needs_input_grad[4]
    gres = grad_res.contiguous() if (want_residual and grad_res is not None and grad_res.numel()) else None
    gx, gw, gb = torch.ops.sfno_b200.spectral_conv_backward(plan_fwd, plan_inv, handle, weight, x, grad_y.contiguous(), gres, need_x, need_w,
                                                            need_b)
    return (None, None, None, gw if need_w else None, gb.reshape(bias_shape) if need_b else None, gx if need_x else None, None, None, None)


def _conv_ex_setup(ctx, inputs, output):
    x, weight, bias, residual, act, dropout_p, _, _, precision = inputs
    ctx.cfg = (int(act), float(dropout_p), int(precision), bias is not None, residual is not None, tuple(weight.shape))
    ctx.save_for_backward(x, weight)


def _conv_ex_backward(ctx, grad):
    act, p, precision, has_bias, has_res, wshape = ctx.cfg
    if act != 0 or p != 0.0:
        raise NotImplementedError("conv1x1_ex with a fused activation / dropout has no backward: use act=0, dropout_p=0 and apply them "
                                  "outside (the trainable forward of SphericalFourierNeuralOperatorNet does)")
    x, weight = ctx.saved_tensors
    need_x, need_w, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1], has_bias and ctx.needs_input_grad[2]
    gx, gw, gb = torch.ops.sfno_b200.conv1x1_backward(x, grad.contiguous(), weight, need_x, need_w, need_b, precision)
    gx = gx if need_x else None
    gw = gw.reshape(wshape) if need_w else None
    gb = gb if need_b else None
    return gx, gw, (gb if has_bias else None), (grad if has_res else None), None, None, None, None, None



def _sht_forward_setup(ctx, inputs, output):
    plan, x, _, _ = inputs
    ctx.plan, ctx.grid = int(plan), (int(x.shape[-2]), int(x.shape[-1]))


def _sht_forward_backward(ctx, grad):
    return None, torch.ops.sfno_b200.sht_forward_adjoint(ctx.plan, grad.contiguous(), ctx.grid[0], ctx.grid[1]), None, None


def _sht_inverse_setup(ctx, inputs, output):
    plan, coeffs, _, _ = inputs
    ctx.plan, ctx.modes = int(plan), (int(coeffs.shape[-3]), int(coeffs.shape[-2]))


def _sht_inverse_backward(ctx, grad):
    return None, torch.ops.sfno_b200.sht_inverse_adjoint(ctx.plan, grad.contiguous(), ctx.modes[0], ctx.modes[1]), None, None


def _contract_setup(ctx, inputs, output):
    op, x, w = inputs
    ctx.op = int(op)
    ctx.save_for_backward(x, w)


def _contract_backward(ctx, grad):
    x, w = ctx.saved_tensors
    gx, gw = torch.ops.sfno_b200.spectral_contract_backward(ctx.op, x, w, grad.contiguous())
    return None, gx.reshape(x.shape), gw.reshape(w.shape)


def _conv_setup(ctx, inputs, output):
    x, weight, bias, residual, act = inputs
    ctx.act, ctx.has_bias, ctx.has_res, ctx.wshape = int(act), bias is not None, residual is not None, tuple(weight.shape)
    ctx.save_for_backward(x, weight)


def _conv_backward(ctx, grad):
    if ctx.act != 0:
        raise NotImplementedError("conv1x1 with a fused activation has no backward: use act=0 and apply the activation outside "
                                  "(the trainable forward of SphericalFourierNeuralOperatorNet does)")
    x, weight = ctx.saved_tensors
    g = grad.contiguous()
    w2 = weight.reshape(weight.shape[0], -1)
    gx = torch.ops.sfno_b200.conv1x1(g, w2.t().contiguous(), None, None, 0).reshape(x.shape) if ctx.needs_input_grad[0] else None
    gw = gb = None
    if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
        gw, gb = torch.ops.sfno_b200.conv1x1_weight_grad(x, g)
        gw = gw.reshape(ctx.wshape)
    return gx, gw, (gb if ctx.has_bias else None), (grad if ctx.has_res else None), None


def _norm_setup(ctx, inputs, output):
    x, gamma, beta, scale, shift, eps = inputs
    ctx.eps = float(eps)
    ctx.flags = (gamma is not None, beta is not None, scale is not None, shift is not None)
    ctx.save_for_backward(x, gamma, beta, scale, shift)


def _norm_backward(ctx, grad):
    x, gamma, beta, scale, shift = ctx.saved_tensors
    B, C = x.shape[0], x.shape[1]
    one_plus = (1.0 + scale) if scale is not None else None                       # [B, C]
    a = None
    if gamma is not None or one_plus is not None:
        a = torch.ones(B, C, dtype=torch.float32, device=x.device)
        if gamma is not None:
            a = a * gamma.reshape(1, C)
        if one_plus is not None:
            a = a * one_plus
    gx, da, dd = torch.ops.sfno_b200.instance_norm_backward(x, grad.contiguous(), a, ctx.eps)
    # y = xhat * gamma (1 + scale) + beta (1 + scale) + shift: chain rule on the [B, C] sums
    s1 = one_plus if one_plus is not None else 1.0
    g_gamma = (da * s1).sum(0).reshape(gamma.shape) if gamma is not None else None
    g_beta = (dd * s1).sum(0).reshape(beta.shape) if beta is not None else None
    g_scale = None
    if scale is not None:
        g_scale = da * (gamma.reshape(1, C) if gamma is not None else 1.0)
        if beta is not None:
            g_scale = g_scale + dd * beta.reshape(1, C)
        g_scale = g_scale.reshape(scale.shape)
    g_shift = dd.reshape(shift.shape) if shift is not None else None
    return gx.reshape(x.shape), g_gamma, g_beta, g_scale, g_shift, None


torch.library.register_autograd("sfno_b200::sht_forward", _sht_forward_backward, setup_context=_sht_forward_setup)
torch.library.register_autograd("sfno_b200::sht_inverse", _sht_inverse_backward, setup_context=_sht_inverse_setup)
torch.library.register_autograd("sfno_b200::spectral_contract", _contract_backward, setup_context=_contract_setup)
torch.library.register_autograd("sfno_b200::conv1x1", _conv_backward, setup_context=_conv_setup)
torch.library.register_autograd("sfno_b200::instance_norm", _norm_backward, setup_context=_norm_setup)
torch.library.register_autograd("sfno_b200::spectral_conv_diff", _spec_backward, setup_context=_spec_setup)
torch.library.register_autograd("sfno_b200::conv1x1_ex", _conv_ex_backward, setup_context=_conv_ex_setup)
