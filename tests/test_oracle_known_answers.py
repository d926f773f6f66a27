"""Pin the restated torch_harmonics (oracle/th_shim.py) by mathematical known answers -- the reference
holds no tests or golden vectors for this boundary."""
import numpy as np
import pytest
import torch
from scipy.special import sph_harm_y

from oracle import th_shim


def test_legpoly_matches_scipy_spherical_harmonics():
    theta = np.linspace(0.05, np.pi - 0.05, 17)
    tab = th_shim.precompute_legpoly(12, 15, theta)  # inclusive bounds -> [13, 16, 17]
    for m in range(13):
        for l in range(16):
            want = sph_harm_y(l, m, theta, 0.0).real if l >= m else np.zeros_like(theta)
            assert np.abs(tab[m, l] - want).max() < 5e-13, (m, l)


@pytest.mark.parametrize("n", [8, 33, 120, 721])
def test_quadrature_rules(n):
    x, w = th_shim.legendre_gauss_weights(n)
    xr, wr = np.polynomial.legendre.leggauss(n)
    assert np.allclose(x, xr) and np.allclose(w, wr)
    for rule in (th_shim.clenshaw_curtiss_weights, th_shim.legendre_gauss_weights, th_shim.lobatto_weights):
        if rule is th_shim.lobatto_weights and n > 200:
            continue
        x, w = rule(n)
        assert abs(w.sum() - 2.0) < 1e-12
        assert abs((w * x ** 2).sum() - 2.0 / 3.0) < 1e-12
        assert abs((w * x ** 3).sum()) < 1e-12
        assert np.all(np.diff(x) > 0)


@pytest.mark.parametrize("grid,nlat,nlon,L,M", [("legendre-gauss", 24, 48, 24, 25), ("equiangular", 49, 96, 24, 25),
                                                ("legendre-gauss", 120, 240, 120, 121)])
def test_analysis_of_synthesis_is_identity(grid, nlat, nlon, L, M):
    sht = th_shim.RealSHT(nlat, nlon, lmax=L, mmax=M, grid=grid)
    isht = th_shim.InverseRealSHT(nlat, nlon, lmax=L, mmax=M, grid=grid)
    g = torch.Generator().manual_seed(0)
    c = torch.randn(2, L, M, 2, generator=g, dtype=torch.float64)
    ii, jj = torch.triu_indices(L, M, offset=1)
    c[:, ii, jj] = 0  # l < m
    c[:, :, 0, 1] = 0  # Im(m = 0)
    if M - 1 == nlon // 2:
        c[:, :, M - 1] = 0  # Nyquist column
    c = torch.view_as_complex(c)
    back = sht(isht(c))
    assert (back - c).abs().max() < 1e-10


def test_single_harmonic_delta_response():
    nlat, nlon, L, M = 32, 64, 16, 17
    sht = th_shim.RealSHT(nlat, nlon, lmax=L, mmax=M, grid="legendre-gauss")
    x, _ = th_shim.legendre_gauss_weights(nlat)
    theta = np.flip(np.arccos(x))
    phi = 2 * np.pi * np.arange(nlon) / nlon
    for (l, m) in [(0, 0), (3, 0), (5, 2), (15, 15), (9, 4)]:
        Y = sph_harm_y(l, m, theta[:, None], phi[None, :])
        field = torch.from_numpy(np.ascontiguousarray((Y.real if m == 0 else 2 * Y.real)))
        c = sht(field)
        want = torch.zeros(L, M, dtype=torch.complex128)
        want[l, m] = 1.0
        assert (c - want).abs().max() < 1e-11, (l, m)


def test_irfft_drops_imag_of_dc_and_nyquist():
    nlat, nlon, L, M = 12, 24, 12, 13
    isht = th_shim.InverseRealSHT(nlat, nlon, lmax=L, mmax=M, grid="legendre-gauss")
    g = torch.Generator().manual_seed(1)
    c = torch.view_as_complex(torch.randn(L, M, 2, generator=g, dtype=torch.float64))
    c2 = c.clone()
    c2[:, 0] = c2[:, 0].real + 0j
    c2[:, M - 1] = c2[:, M - 1].real + 0j
    assert (isht(c) - isht(c2)).abs().max() < 1e-12
