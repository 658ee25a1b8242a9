"""CPU tests that pin the oracle: against the reference-generated goldens and analytical identities."""
import numpy as np
import pytest
import torch

from conftest import golden_cases
from oracle import harmonics
from oracle.sfno_oracle import SFNOConfig, SFNOOracle, random_state_dict, rel_l2


@pytest.mark.parametrize("case", golden_cases())
def test_oracle_matches_reference_golden(case, load_golden):
    fx = load_golden(case)
    cfg = SFNOConfig(**fx["cfg"])
    orc = SFNOOracle(cfg, fx["state_dict"])
    orc.taps = {}
    out, t_repr = orc(fx["inputs"], time=fx["time"], condition=fx["condition"], return_time_emb=True)
    assert out.shape == fx["output"].shape
    assert rel_l2(out, fx["output"]) < 2e-6
    if fx["t_repr"] is not None:
        assert rel_l2(t_repr, fx["t_repr"]) < 2e-6
    taps = fx["taps"]
    assert rel_l2(orc.taps["blocks.0.sht"], taps["blocks.0.sht"]) < 2e-6
    for i in range(cfg.num_layers):
        assert rel_l2(orc.taps[f"blocks.{i}.out"], taps[f"blocks.{i}.out"]) < 2e-6


def test_state_dict_layout_matches_reference(load_golden):
    fx = load_golden("sfno_dhconv_12x24")
    cfg = SFNOConfig(**fx["cfg"])
    mine = random_state_dict(cfg)
    assert list(mine.keys()) == list(fx["state_dict"].keys())
    for k, v in fx["state_dict"].items():
        assert tuple(mine[k].shape) == tuple(v.shape), k


@pytest.mark.parametrize("grid", ["legendre-gauss", "equiangular"])
def test_legendre_table_gram_orthonormal(grid):
    # 2*pi * sum_k w_k P_l^m P_l'^m = delta_ll'  (exact for LG; for CC only up to degree nlat-1 total)
    nlat, nlon = 32, 64
    lmax = nlat if grid == "legendre-gauss" else nlat // 2
    weights, pct, _, mmax = harmonics.sht_tables(nlat, nlon, lmax, None, grid)
    for m in (0, 1, 5, lmax - 1):
        gram = 2 * np.pi * weights[m, m:] @ pct[m, m:].T
        assert np.abs(gram - np.eye(lmax - m)).max() < 1e-12


def test_legendre_table_matches_scipy_sph_harm():
    from scipy.special import sph_harm_y

    nlat, nlon = 24, 48
    nodes, _ = harmonics.quadrature("legendre-gauss", nlat)
    colat = np.flip(np.arccos(nodes))
    _, pct, lmax, mmax = harmonics.sht_tables(nlat, nlon, None, None, "legendre-gauss")
    for m in range(0, mmax, 3):
        for l in range(m, lmax, 2):
            ref = sph_harm_y(l, m, colat, 0.0).real
            assert np.abs(pct[m, l] - ref).max() < 1e-13, (l, m)


def test_table_structural_zeros():
    _, pct, lmax, mmax = harmonics.sht_tables(12, 24, None, None, "equiangular")
    for m in range(mmax):
        assert np.all(pct[m, : min(m, lmax)] == 0.0)


def test_lg_round_trip_identity_on_band_limited_field():
    nlat, nlon = 32, 64
    sht = harmonics.RealSHT(nlat, nlon, grid="legendre-gauss").float()
    isht = harmonics.InverseRealSHT(nlat, nlon, grid="legendre-gauss").float()
    x = torch.randn(2, 3, nlat, nlon, generator=torch.Generator().manual_seed(0))
    xb = isht(sht(x))          # band-limit
    xr = isht(sht(xb))
    assert rel_l2(xr, xb) < 2e-6


def test_clenshaw_curtis_weights_sum_and_symmetry():
    for n in (2, 5, 12, 181):
        x, w = harmonics.clenshaw_curtiss_weights(n)
        assert abs(w.sum() - 2.0) < 1e-13
        assert np.allclose(w, w[::-1], atol=1e-15)
        assert np.allclose(x, -x[::-1], atol=1e-15)


def test_oracle_fp64_close_to_fp32(load_golden):
    fx = load_golden("sfno_dhconv_18x36_lg")
    cfg = SFNOConfig(**fx["cfg"])
    o32 = SFNOOracle(cfg, fx["state_dict"])(fx["inputs"], time=fx["time"], condition=fx["condition"])
    o64 = SFNOOracle(cfg, fx["state_dict"], dtype=torch.float64)(fx["inputs"].double(), time=fx["time"].double(),
                                                                   condition=fx["condition"].double())
    assert rel_l2(o32, o64) < 1e-5
