"""``RealSHT`` / ``InverseRealSHT`` with the interface the reference consumes from ``torch_harmonics``
(constructed at ``sfnonet.py:551-554``, attributes read at ``s2convolutions.py:76-83``), executed by the custom ops
``torch.ops.sfno_b200.sht_forward / sht_inverse`` (``ops.py``) over ``sfno_sht_forward`` / ``sfno_sht_inverse`` of the
C ABI: longitude DFT and Legendre contraction as GEMMs on the GPU.  Tables are built in fp64 by the library (``sfno_sht_tables_host``).
"""
from __future__ import annotations

import ctypes

import torch
import torch.nn as nn

from . import _lib
from . import ops as _ops  # noqa: F401  (registers torch.ops.sfno_b200.*)
from ._util import require_cuda_f32


class _ShtBase(nn.Module):
    def __init__(self, nlat, nlon, lmax=None, mmax=None, grid="lobatto", norm="ortho", csphase=True, precision="fp32"):
        super().__init__()
        if grid not in _lib.SFNO_GRID:
            raise ValueError("Unknown quadrature mode")  # same error class as torch_harmonics
        if norm != "ortho" or not csphase:
            raise NotImplementedError("only norm='ortho', csphase=True (the reference's configuration) is built")
        self.nlat, self.nlon, self.grid, self.norm, self.csphase = nlat, nlon, grid, norm, csphase
        self.lmax = lmax or nlat
        self.mmax = mmax or nlon // 2 + 1
        self.precision = precision
        self._plans: dict = {}

    def extra_repr(self):
        return f"nlat={self.nlat}, nlon={self.nlon}, lmax={self.lmax}, mmax={self.mmax}, grid={self.grid}, precision={self.precision}"

    def _plan(self, device):
        key = str(device)
        if key not in self._plans:
            handle = ctypes.c_void_p()
            with torch.cuda.device(device):
                _lib.check(_lib.lib().sfno_sht_plan_create(ctypes.byref(handle), self.nlat, self.nlon, self.lmax, self.mmax,
                                                           _lib.SFNO_GRID[self.grid], _lib.SFNO_PREC[self.precision]),
                           "sfno_sht_plan_create")
            self._plans[key] = handle
        return self._plans[key]

    def __del__(self):
        try:
            for h in self._plans.values():
                _lib.lib().sfno_sht_plan_destroy(h)
        except Exception:
            pass

    def tables(self):
        """(nodes, quad_w, weights[m,l,k], pct[m,l,k]) as fp64 CPU tensors, computed by the library on the host."""
        n = self.mmax * self.lmax * self.nlat
        nodes = (ctypes.c_double * self.nlat)()
        qw = (ctypes.c_double * self.nlat)()
        wts = (ctypes.c_double * n)()
        pct = (ctypes.c_double * n)()
        _lib.check(_lib.lib().sfno_sht_tables_host(self.nlat, self.nlon, self.lmax, self.mmax, _lib.SFNO_GRID[self.grid],
                                                   nodes, qw, wts, pct), "sfno_sht_tables_host")
        shape = (self.mmax, self.lmax, self.nlat)
        return (torch.tensor(list(nodes), dtype=torch.float64), torch.tensor(list(qw), dtype=torch.float64),
                torch.frombuffer(wts, dtype=torch.float64).reshape(shape).clone(),
                torch.frombuffer(pct, dtype=torch.float64).reshape(shape).clone())


class RealSHT(_ShtBase):
    """x[..., nlat, nlon] (real) -> complex64 [..., lmax, mmax]."""

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        assert x.shape[-2] == self.nlat
        assert x.shape[-1] == self.nlon
        xf = require_cuda_f32(x, "x")
        out = torch.ops.sfno_b200.sht_forward(self._plan(xf.device).value, xf, self.lmax, self.mmax)
        return torch.view_as_complex(out)


class InverseRealSHT(_ShtBase):
    """complex [..., lmax, mmax] -> real fp32 [..., nlat, nlon]."""

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        assert x.shape[-2] == self.lmax
        assert x.shape[-1] == self.mmax
        if not x.is_cuda:
            raise RuntimeError("coefficients must be a CUDA tensor: the B200 path has no CPU fallback")
        xr = torch.view_as_real(x.to(torch.complex64).contiguous())
        return torch.ops.sfno_b200.sht_inverse(self._plan(xr.device).value, xr, self.nlat, self.nlon)
