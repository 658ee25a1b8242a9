"""TEST INFRASTRUCTURE ONLY -- CPU restatement of ``torch_harmonics.RealSHT`` / ``InverseRealSHT``.

``torch-harmonics`` is a third-party dependency of the reference that is *not* vendored under
``/root/reference`` and is unpinned there (``setup.py:98,175``; ``environment/install_dependencies.sh:13``).
This file restates the published 0.6.x algorithm (``quadrature.py``, ``legendre.py``, ``sht.py``) as
summarised in SURVEY.md Appendix A; the reference's call sites are ``src/models/sfno/sfnonet.py:539-554``
and ``src/models/sfno/s2convolutions.py:165,168,186``.

Parity unpinned by the reference (it has no tests).  Pinned here by analytical identities in
``tests/test_oracle.py``: Gram orthonormality of the tables, agreement with
``scipy.special.sph_harm_y`` and the Legendre-Gauss round trip on band-limited fields.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn


# ----------------------------------------------------------------------------------------------
# quadrature nodes / weights on [-1, 1]
# ----------------------------------------------------------------------------------------------
def legendre_gauss_weights(n: int):
    """Gauss-Legendre nodes (ascending in cos(theta)) and weights."""
    nodes, weights = np.polynomial.legendre.leggauss(n)
    return nodes, weights


def clenshaw_curtiss_weights(n: int):
    """Clenshaw-Curtis nodes ``cos(linspace(pi, 0, n))`` and weights (classic FFT construction)."""
    assert n > 1
    nodes = np.cos(np.linspace(np.pi, 0.0, n))
    if n == 2:
        return nodes, np.array([1.0, 1.0])
    n1 = n - 1
    odd = np.arange(1, n1, 2)
    n_odd = len(odd)
    rest = n1 - n_odd
    v = np.concatenate([2.0 / odd / (odd - 2), 1.0 / odd[-1:], np.zeros(rest)])
    v = 0.0 - v[:-1] - v[-1:0:-1]
    g = -np.ones(n1)
    g[n_odd] += n1
    g[rest] += n1
    g = g / (n1**2 - 1 + (n1 % 2))
    w = np.fft.ifft(v + g).real
    w = np.concatenate((w, w[:1]))
    return nodes, w


def quadrature(grid: str, nlat: int):
    if grid == "legendre-gauss":
        return legendre_gauss_weights(nlat)
    if grid == "equiangular":
        return clenshaw_curtiss_weights(nlat)
    raise ValueError(f"Unknown quadrature mode {grid!r}")


# ----------------------------------------------------------------------------------------------
# orthonormal associated Legendre functions  P[m, l, k]  (fp64)
# ----------------------------------------------------------------------------------------------
def legpoly(mmax: int, lmax: int, x: np.ndarray, csphase: bool = True) -> np.ndarray:
    """``norm='ortho'`` table of shape [mmax, lmax, len(x)], zero for l < m."""
    n = max(mmax, lmax)
    p = np.zeros((n, n, len(x)), dtype=np.float64)
    p[0, 0, :] = 1.0 / np.sqrt(4.0 * np.pi)
    for l in range(1, n):
        p[l - 1, l, :] = np.sqrt(2 * l + 1) * x * p[l - 1, l - 1, :]
        p[l, l, :] = np.sqrt((2 * l + 1) * (1 + x) * (1 - x) / 2 / l) * p[l - 1, l - 1, :]
    for l in range(2, n):
        for m in range(0, l - 1):
            a = np.sqrt((2 * l - 1) / (l - m) * (2 * l + 1) / (l + m))
            b = np.sqrt((l + m - 1) / (l - m) * (2 * l + 1) / (2 * l - 3) * (l - m - 1) / (l + m))
            p[m, l, :] = x * a * p[m, l - 1, :] - b * p[m, l - 2, :]
    p = p[:mmax, :lmax]
    if csphase:
        p[1::2] *= -1.0
    return p


def sht_tables(nlat: int, nlon: int, lmax: int | None, mmax: int | None, grid: str):
    """Returns (weights[m,l,k], pct[m,l,k], lmax, mmax) in fp64.

    ``weights`` is the analysis table (Legendre x quadrature weight), ``pct`` the synthesis table.
    Colatitudes run north -> south (``flip(arccos(nodes))``) while the weights keep node order; both
    supported quadratures are symmetric so this is immaterial, but it is restated as published.
    """
    nodes, w = quadrature(grid, nlat)
    lmax = lmax or nlat
    mmax = mmax or nlon // 2 + 1
    colat = np.flip(np.arccos(nodes))
    pct = legpoly(mmax, lmax, np.cos(colat))
    weights = pct * w[None, None, :]
    return weights, pct, lmax, mmax


def round_tf32(t: torch.Tensor) -> torch.Tensor:
    """fp32 -> nearest TF32 value (10-bit mantissa), kept in fp32: the operand rounding of a TF32 matmul.  Used only to
    EMULATE what the reference computes under ``torch.set_float32_matmul_precision("high")``
    (``src/utilities/config_utils.py:310-313``) next to the B200 library's tf32 mode; the parity oracle is fp32."""
    if t.dtype != torch.float32:
        return t
    bits = t.contiguous().view(torch.int32)
    return ((bits + 0x1000) & ~0x1FFF).view(torch.float32)


# ----------------------------------------------------------------------------------------------
# modules with the attributes the reference reads (.nlat .nlon .lmax .mmax .grid, .float())
# ----------------------------------------------------------------------------------------------
class RealSHT(nn.Module):
    def __init__(self, nlat, nlon, lmax=None, mmax=None, grid="lobatto", norm="ortho", csphase=True):
        super().__init__()
        assert norm == "ortho" and csphase, "only the configuration used by the reference is restated"
        self.nlat, self.nlon, self.grid, self.norm, self.csphase = nlat, nlon, grid, norm, csphase
        weights, _, self.lmax, self.mmax = sht_tables(nlat, nlon, lmax, mmax, grid)
        self.register_buffer("weights", torch.from_numpy(weights), persistent=False)

    def extra_repr(self):
        return f"nlat={self.nlat}, nlon={self.nlon}, lmax={self.lmax}, mmax={self.mmax}, grid={self.grid}"

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        assert x.shape[-2] == self.nlat
        assert x.shape[-1] == self.nlon
        xf = 2.0 * torch.pi * torch.fft.rfft(x, dim=-1, norm="forward")
        xf = torch.view_as_real(xf)
        w = self.weights.to(xf.dtype)
        if getattr(self, "tf32_matmul", False):
            xf, w = round_tf32(xf), round_tf32(w)
        re = torch.einsum("...km,mlk->...lm", xf[..., : self.mmax, 0], w)
        im = torch.einsum("...km,mlk->...lm", xf[..., : self.mmax, 1], w)
        return torch.view_as_complex(torch.stack((re, im), dim=-1).contiguous())


class InverseRealSHT(nn.Module):
    def __init__(self, nlat, nlon, lmax=None, mmax=None, grid="lobatto", norm="ortho", csphase=True):
        super().__init__()
        assert norm == "ortho" and csphase, "only the configuration used by the reference is restated"
        self.nlat, self.nlon, self.grid, self.norm, self.csphase = nlat, nlon, grid, norm, csphase
        _, pct, self.lmax, self.mmax = sht_tables(nlat, nlon, lmax, mmax, grid)
        self.register_buffer("pct", torch.from_numpy(pct), persistent=False)

    def extra_repr(self):
        return f"nlat={self.nlat}, nlon={self.nlon}, lmax={self.lmax}, mmax={self.mmax}, grid={self.grid}"

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        assert x.shape[-2] == self.lmax
        assert x.shape[-1] == self.mmax
        xr = torch.view_as_real(x)
        p = self.pct.to(xr.dtype)
        if getattr(self, "tf32_matmul", False):
            xr, p = round_tf32(xr), round_tf32(p)
        re = torch.einsum("...lm,mlk->...km", xr[..., 0], p)
        im = torch.einsum("...lm,mlk->...km", xr[..., 1], p)
        xs = torch.view_as_complex(torch.stack((re, im), dim=-1).contiguous())
        return torch.fft.irfft(xs, n=self.nlon, dim=-1, norm="forward")
