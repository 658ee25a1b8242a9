"""TEST-ONLY host emulation of the op sequence executed by ``sfno_net_forward`` (csrc/net.cu).

It re-derives, in PyTorch on the CPU and on the library's *internal layouts*, the algebra the CUDA path
relies on -- DFT-as-GEMM tables, InstanceNorm/time affine folded into the DFT epilogue and into per-sample
conv weights, the packed real form of the complex dhconv weights, flipped inverse transforms -- so that the
decomposition can be checked against the oracle without a GPU.  It is not part of the product and shares
no code with it; a mismatch here means the design (not a kernel) is wrong.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


def round_up(a, b):
    return (a + b - 1) // b * b


def dft_forward_table(nlon, mmax, dtype):
    j = np.arange(nlon)
    m = np.arange(mmax)
    ang = 2 * np.pi * ((m[:, None] * j[None, :]) % nlon) / nlon
    e = np.zeros((2 * mmax, nlon))
    e[0::2] = 2 * np.pi / nlon * np.cos(ang)
    e[1::2] = -2 * np.pi / nlon * np.sin(ang)
    return torch.from_numpy(e).to(dtype)


def dft_inverse_table(nlon, mmax, dtype):
    j = np.arange(nlon)
    m = np.arange(mmax)
    ang = 2 * np.pi * ((j[:, None] * m[None, :]) % nlon) / nlon
    cm = np.where((m == 0) | (2 * m == nlon), 1.0, 2.0)
    e = np.zeros((nlon, 2 * mmax))
    e[:, 0::2] = cm * np.cos(ang)
    e[:, 1::2] = np.where((m == 0) | (2 * m == nlon), 0.0, -cm * np.sin(ang))
    return torch.from_numpy(e).to(dtype)


class NetEmulator:
    def __init__(self, cfg, sd, tables_fn, dtype=torch.float64):
        """cfg: oracle SFNOConfig; sd: reference state_dict; tables_fn(grid) -> (weights[m,l,k], pct[m,l,k]) fp64."""
        self.cfg, self.dt = cfg, dtype
        self.sd = {k: v.to(dtype) for k, v in sd.items()}
        H, W = cfg.spatial_shape
        self.H, self.W, self.C = H, W, cfg.embed_dim
        self.L = int(H * cfg.hard_thresholding_fraction)
        self.M = int((W // 2 + 1) * cfg.hard_thresholding_fraction)
        self.Kp = round_up(H, 8)
        self.tab = {}
        for grid in {cfg.data_grid, "legendre-gauss"}:
            wq, pct = tables_fn(grid)
            self.tab[grid] = (wq.to(dtype), pct.to(dtype))
        self.efwd = dft_forward_table(W, self.M, dtype)
        self.einv = dft_inverse_table(W, self.M, dtype)

    # ---- ops on internal layouts (see csrc/ops.cuh) -------------------------------------------------------
    def dft(self, x, a, d):
        """x [B,C,H,W]; a,d [B,C] -> F [M][B][2][C][Kp]"""
        B, C = x.shape[:2]
        acc = torch.einsum("bckj,nj->bckn", x, self.efwd)  # rows (b,c,k), cols n = 2m+ri
        acc = a[:, :, None, None] * acc
        acc[..., 0] += 2 * math.pi * d[:, :, None]
        Fm = torch.zeros(self.M, B, 2, C, self.Kp, dtype=self.dt)
        Fm[..., : self.H] = acc.reshape(B, C, self.H, self.M, 2).permute(3, 0, 4, 1, 2)
        return Fm

    def leg(self, Fm, grid):
        """F [M][B][2][C][Kp] -> X [L][M][B][2][C]"""
        wq = self.tab[grid][0]  # [m][l][k]
        return torch.einsum("mbrck,mlk->lmbrc", Fm[..., : self.H], wq)

    def dhconv(self, X, w):
        """X [L][M][B][2][C], w [Cin][Cout][L][2] -> Y [M][L][B][2][C] via the packed real weight"""
        C = self.C
        wr, wi = w[..., 0].permute(2, 1, 0), w[..., 1].permute(2, 1, 0)  # [L][o][c]
        Wp = torch.cat((torch.cat((wr, -wi), dim=2), torch.cat((wi, wr), dim=2)), dim=1)  # [L][(ri',o)][(ri,c)]
        Lx, M, B = X.shape[:3]
        xin = X.reshape(Lx, M * B, 2 * C)
        D = torch.einsum("lpk,lnk->lpn", Wp, xin)  # [L][(ri',o)][(m,b)]
        return D.reshape(Lx, 2, C, M, B).permute(3, 0, 4, 1, 2).contiguous()

    def diagonal(self, X, w):
        """w [Cin][Cout][L][M][2]"""
        Xc = torch.complex(X[..., 0, :], X[..., 1, :])  # [L][M][B][C]
        wc = torch.complex(w[..., 0], w[..., 1])         # [i][o][L][M]
        Yc = torch.einsum("lmbi,iolm->mlbo", Xc, wc)
        return torch.stack((Yc.real, Yc.imag), dim=3)

    def ileg(self, S, grid, x_layout):
        """S = Y [M][L][B][2][C] (or X [L][M][B][2][C]) -> G [M][2][B][C][Kp]"""
        pct = self.tab[grid][1]  # [m][l][k]
        if x_layout:
            S = S.permute(1, 0, 2, 3, 4)
        D = torch.einsum("mlk,mlbrc->mrbck", pct, S)
        G = torch.zeros(self.M, 2, S.shape[2], self.C, self.Kp, dtype=self.dt)
        G[..., : self.H] = D
        return G

    def idft(self, G, bias=None, add=None, act=None):
        """G [M][2][B][C][Kp] -> [B][C][H][W]"""
        Gk = G[..., : self.H].permute(2, 3, 4, 0, 1).reshape(G.shape[2], self.C, self.H, 2 * self.M)  # kk = 2m+ri
        y = torch.einsum("bckn,jn->bckj", Gk, self.einv)
        if bias is not None:
            y = y + bias.reshape(1, -1, 1, 1)
        if add is not None:
            y = y + add
        return act(y) if act is not None else y

    @staticmethod
    def conv(x, w, b=None):
        """per-sample weights allowed: w [o,c] or [B,o,c]; b [o] or [B,o]"""
        if w.dim() == 2:
            y = torch.einsum("oc,bchw->bohw", w, x)
        else:
            y = torch.einsum("boc,bchw->bohw", w, x)
        if b is not None:
            y = y + (b[None, :, None, None] if b.dim() == 1 else b[:, :, None, None])
        return y

    def stats_affine(self, x, gamma, beta, ts):
        """per-(b,c) (a, d) of InstanceNorm -> time scale/shift (pointwise.cuh: norm_affine_kernel)"""
        B, C = x.shape[:2]
        a = torch.ones(B, C, dtype=self.dt)
        d = torch.zeros(B, C, dtype=self.dt)
        if self.cfg.normalization_layer == "instance_norm":
            mu = x.mean(dim=(2, 3))
            var = x.var(dim=(2, 3), unbiased=False)
            rstd = 1.0 / torch.sqrt(var + 1e-6)
            a = gamma[None] * rstd
            d = beta[None] - gamma[None] * mu * rstd
        if ts is not None:
            sc, sf = ts[:, :C] + 1.0, ts[:, C:]
            a = a * sc
            d = d * sc + sf
        return a, d

    def forward(self, inputs, time=None, condition=None):
        cfg, sd, C = self.cfg, self.sd, self.C
        act = {"gelu": F.gelu, "relu": F.relu, "silu": F.silu}[cfg.activation_function]
        x_in = inputs if condition is None else torch.cat((inputs, condition), dim=1)
        x_in = x_in.to(self.dt)
        x = act(self.conv(x_in, sd["encoder.0.weight"][:, :, 0, 0], sd["encoder.0.bias"]))
        x = self.conv(x, sd["encoder.2.weight"][:, :, 0, 0])
        if cfg.pos_embed:
            x = x + sd["pos_embed"]
        ts_all = None
        if cfg.with_time_emb:
            t = time.to(self.dt)
            if cfg.time_rescale:
                t = t * (1000.0 / (cfg.max_time - cfg.min_time)) + (-cfg.min_time)
            half = C // 2
            f = torch.exp(torch.arange(half, dtype=self.dt) * -(math.log(10000) / (half - 1)))
            e = t[:, None] * f[None]
            e = torch.cat((e.sin(), e.cos()), dim=-1)
            e = F.gelu(F.linear(e, sd["time_emb_mlp.1.weight"], sd["time_emb_mlp.1.bias"]))
            t_repr = F.linear(e, sd["time_emb_mlp.3.weight"], sd["time_emb_mlp.3.bias"])
            ts_all = [F.linear(F.silu(t_repr), sd[f"blocks.{i}.time_mlp.1.weight"], sd[f"blocks.{i}.time_mlp.1.bias"])
                      for i in range(cfg.num_layers)]
        nl = cfg.num_layers
        fc2 = 3 if cfg.dropout_mlp > 0 else 2
        for i in range(nl):
            p = f"blocks.{i}."
            fwd_grid = cfg.data_grid if i == 0 else "legendre-gauss"
            inv_grid = cfg.data_grid if i == nl - 1 else "legendre-gauss"
            scale_residual = fwd_grid != inv_grid
            ts_i = ts_all[i] if ts_all is not None else None
            before = cfg.with_time_emb and cfg.time_scale_shift_before_filter
            a0, d0 = self.stats_affine(x, sd.get(p + "norm0.weight"), sd.get(p + "norm0.bias"), ts_i if before else None)
            Fm = self.dft(x, a0, d0)
            X = self.leg(Fm, fwd_grid)
            res = None
            if scale_residual:
                res = self.idft(self.ileg(X, inv_grid, True))
            w = sd[p + "filter.filter.weight"]
            Y = self.dhconv(X, w) if cfg.operator_type == "dhconv" else self.diagonal(X, w)
            G = self.ileg(Y, inv_grid, False)
            wsk, bsk = sd[p + "inner_skip.weight"][:, :, 0, 0], sd[p + "inner_skip.bias"]
            if scale_residual:
                t1 = self.conv(res, wsk, bsk)
            else:  # fold (a0,d0) into per-sample weights
                t1 = self.conv(x, wsk[None] * a0[:, None, :], bsk[None] + d0 @ wsk.T)
            t1 = self.idft(G, sd[p + "filter.filter.bias"].reshape(-1), t1, act)
            after = cfg.with_time_emb and not cfg.time_scale_shift_before_filter
            a1, d1 = self.stats_affine(t1, sd.get(p + "norm1.weight"), sd.get(p + "norm1.bias"), ts_i if after else None)
            w1, b1 = sd[p + "mlp.fwd.0.weight"][:, :, 0, 0], sd[p + "mlp.fwd.0.bias"]
            hd = act(self.conv(t1, w1[None] * a1[:, None, :], b1[None] + d1 @ w1.T))
            out = self.conv(hd, sd[p + f"mlp.fwd.{fc2}.weight"][:, :, 0, 0], sd[p + f"mlp.fwd.{fc2}.bias"])
            if scale_residual:
                x = out + res
            else:
                x = out + (a0[:, :, None, None] * x + d0[:, :, None, None])
        if cfg.big_skip:
            x = torch.cat((x, x_in), dim=1)
        x = act(self.conv(x, sd["decoder.0.weight"][:, :, 0, 0], sd["decoder.0.bias"]))
        return self.conv(x, sd["decoder.2.weight"][:, :, 0, 0])
