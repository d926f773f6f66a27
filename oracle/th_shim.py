"""ORACLE (test infrastructure, not product code) -- restated `torch_harmonics`.

The reference imports `torch_harmonics` (PyPI `torch-harmonics`, UN-PINNED in
/root/reference/conda_environment.yml:62, 0.6.x era) for RealSHT / InverseRealSHT /
quadrature.  The package is not vendored under /root/reference and is not installed
in this image, so this file restates its published algorithm on CPU (numpy fp64
tables, torch fp32 transforms, same op order as the library).

PARITY STATUS: the reference holds no tests, golden vectors or fixtures for this
boundary, so this shim is pinned by mathematical known-answer
tests instead (tests/test_oracle_known_answers.py): scipy.special.sph_harm_y,
numpy leggauss, quadrature exactness, analysis(synthesis(c)) == c, single-harmonic
delta responses, irfft DC/Nyquist convention.  Against the *library itself* parity
is UNPINNED (no copy of the library exists here to run).

Reference call sites this shim serves:
  MSFNO/Models/sfno/sfnonet.py:537-548  (RealSHT / InverseRealSHT construction)
  MSFNO/Models/sfno/sfnonet.py:551-555  (weights *= 1e5, pct /= 1e5 after construction)
  MSFNO/Models/sfno/layers.py:405,421,629,638 (forward calls)
  MSFNO/Models/losses.py:90,129 (quadrature.legendre_gauss_weights(n, -1, 1)[1])

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.
"""
import types

import numpy as np
import torch
import torch.nn as nn


# --------------------------------------------------------------------------- quadrature
def legendre_gauss_weights(n, a=-1.0, b=1.0):
    """Gauss-Legendre nodes/weights on [a,b], nodes ascending."""
    xlg, wlg = np.polynomial.legendre.leggauss(n)
    xlg = (b - a) * 0.5 * xlg + (b + a) * 0.5
    wlg = wlg * (b - a) * 0.5
    return xlg, wlg


def lobatto_weights(n, a=-1.0, b=1.0, tol=1e-16, maxiter=100):
    """Gauss-Lobatto nodes/weights (Newton iteration on the Legendre Vandermonde)."""
    wlg = np.zeros((n,))
    tlg = np.zeros((n,))
    tmp = np.zeros((n,))
    vdm = np.zeros((n, n))
    # Chebyshev nodes as first guess
    for i in range(n):
        tlg[i] = -np.cos(np.pi * i / (n - 1))
    tmp = 2.0
    for _ in range(maxiter):
        tmp = tlg
        vdm[:, 0] = 1.0
        vdm[:, 1] = tlg
        for k in range(2, n):
            vdm[:, k] = ((2 * k - 1) * tlg * vdm[:, k - 1] - (k - 1) * vdm[:, k - 2]) / k
        tlg = tmp - (tlg * vdm[:, n - 1] - vdm[:, n - 2]) / (n * vdm[:, n - 1])
        if max(abs(tlg - tmp).flatten()) < tol:
            break
    wlg = 2.0 / ((n * (n - 1)) * (vdm[:, n - 1] ** 2))
    tlg = (b - a) * 0.5 * tlg + (b + a) * 0.5
    wlg = wlg * (b - a) * 0.5
    return tlg, wlg


def clenshaw_curtiss_weights(n, a=-1.0, b=1.0):
    """Clenshaw-Curtis nodes/weights (Waldvogel's FFT construction), both poles included."""
    assert n > 1
    tcc = np.cos(np.linspace(np.pi, 0, n))
    if n == 2:
        wcc = np.array([1.0, 1.0])
    else:
        n1 = n - 1
        N = np.arange(1, n1, 2)
        l = len(N)
        m = n1 - l
        v = np.concatenate([2 / N / (N - 2), 1 / N[-1:], np.zeros(m)])
        v = 0 - v[:-1] - v[-1:0:-1]
        g0 = -np.ones(n1)
        g0[l] = g0[l] + n1
        g0[m] = g0[m] + n1
        g = g0 / (n1 ** 2 - 1 + (n1 % 2))
        wcc = np.fft.ifft(v + g).real
        wcc = np.concatenate((wcc, wcc[:1]))
    tcc = (b - a) * 0.5 * tcc + (b + a) * 0.5
    wcc = wcc * (b - a) * 0.5
    return tcc, wcc


# --------------------------------------------------------------------------- Legendre tables
def legpoly(mmax, lmax, x, norm="ortho", inverse=False, csphase=True):
    """Normalised associated Legendre functions P_l^m(x), table [mmax+1?]: returns
    vdm[m, l, k] for m <= mmax, l <= lmax (INCLUSIVE bounds, as in the library's
    `precompute_legpoly`, which is called with mmax-1 / lmax-1 by the transforms)."""
    nmax = max(mmax, lmax)
    vdm = np.zeros((nmax + 1, nmax + 1, len(x)), dtype=np.float64)

    norm_factor = 1.0 if norm == "ortho" else np.sqrt(4 * np.pi)
    norm_factor = 1.0 / norm_factor if inverse else norm_factor

    vdm[0, 0, :] = norm_factor / np.sqrt(4 * np.pi)

    # diagonal and first off-diagonal
    for l in range(1, nmax + 1):
        vdm[l - 1, l, :] = np.sqrt(2 * l + 1) * x * vdm[l - 1, l - 1, :]
        vdm[l, l, :] = np.sqrt((2 * l + 1) * (1 + x) * (1 - x) / 2 / l) * vdm[l - 1, l - 1, :]

    # three-term recurrence in l for fixed m
    for l in range(2, nmax + 1):
        for m in range(0, l - 1):
            vdm[m, l, :] = (
                x * np.sqrt((2 * l - 1) / (l - m) * (2 * l + 1) / (l + m)) * vdm[m, l - 1, :]
                - np.sqrt((l + m - 1) / (l - m) * (2 * l + 1) / (2 * l - 3) * (l - m - 1) / (l + m))
                * vdm[m, l - 2, :]
            )

    if norm == "schmidt":
        for l in range(0, nmax + 1):
            if inverse:
                vdm[:, l, :] = vdm[:, l, :] * np.sqrt(2 * l + 1)
            else:
                vdm[:, l, :] = vdm[:, l, :] / np.sqrt(2 * l + 1)

    vdm = vdm[: mmax + 1, : lmax + 1]

    if csphase:
        for m in range(1, mmax + 1, 2):
            vdm[m] *= -1

    return vdm


def precompute_legpoly(mmax, lmax, t, norm="ortho", inverse=False, csphase=True):
    """Table evaluated at colatitudes t (x = cos t)."""
    return legpoly(mmax, lmax, np.cos(t), norm=norm, inverse=inverse, csphase=csphase)


def _grid_nodes(grid, nlat):
    if grid == "legendre-gauss":
        cost, w = legendre_gauss_weights(nlat, -1, 1)
        lmax_default = nlat
    elif grid == "lobatto":
        cost, w = lobatto_weights(nlat, -1, 1)
        lmax_default = nlat - 1
    elif grid == "equiangular":
        cost, w = clenshaw_curtiss_weights(nlat, -1, 1)
        lmax_default = nlat
    else:
        raise ValueError("Unknown quadrature mode")
    return cost, w, lmax_default


# --------------------------------------------------------------------------- transforms
class RealSHT(nn.Module):
    """Forward real spherical harmonic transform (analysis).

    x[..., nlat, nlon] real -> coeffs[..., lmax, mmax] complex:
      X = 2*pi * rfft(x, dim=-1, norm="forward")
      c[l, m] = sum_k X[k, m] * weights[m, l, k]      (real and imaginary parts separately)
    """

    def __init__(self, nlat, nlon, lmax=None, mmax=None, grid="lobatto", norm="ortho", csphase=True):
        super().__init__()
        self.nlat, self.nlon, self.grid, self.norm, self.csphase = nlat, nlon, grid, norm, csphase
        cost, w, lmax_default = _grid_nodes(grid, nlat)
        self.lmax = lmax or lmax_default
        tq = np.flip(np.arccos(cost))  # colatitudes ascending 0..pi (row 0 = north pole)
        self.mmax = mmax or self.nlon // 2 + 1
        weights = torch.from_numpy(w)
        pct = precompute_legpoly(self.mmax - 1, self.lmax - 1, tq, norm=self.norm, csphase=self.csphase)
        pct = torch.from_numpy(pct)
        weights = torch.einsum("mlk,k->mlk", pct, weights)
        self.register_buffer("weights", weights, persistent=False)

    def extra_repr(self):
        return f"nlat={self.nlat}, nlon={self.nlon},\n lmax={self.lmax}, mmax={self.mmax},\n grid={self.grid}, csphase={self.csphase}"

    def forward(self, x):
        assert x.shape[-2] == self.nlat
        assert x.shape[-1] == self.nlon
        x = 2.0 * torch.pi * torch.fft.rfft(x, dim=-1, norm="forward")
        x = torch.view_as_real(x)
        out_shape = list(x.size())
        out_shape[-3] = self.lmax
        out_shape[-2] = self.mmax
        xout = torch.zeros(out_shape, dtype=x.dtype, device=x.device)
        xout[..., 0] = torch.einsum("...km,mlk->...lm", x[..., : self.mmax, 0], self.weights.to(x.dtype))
        xout[..., 1] = torch.einsum("...km,mlk->...lm", x[..., : self.mmax, 1], self.weights.to(x.dtype))
        return torch.view_as_complex(xout)


class InverseRealSHT(nn.Module):
    """Inverse real SHT (synthesis).

    c[..., lmax, mmax] complex -> x[..., nlat, nlon] real:
      X[k, m] = sum_l c[l, m] * pct[m, l, k]
      x = irfft(X, n=nlon, dim=-1, norm="forward")
    """

    def __init__(self, nlat, nlon, lmax=None, mmax=None, grid="lobatto", norm="ortho", csphase=True):
        super().__init__()
        self.nlat, self.nlon, self.grid, self.norm, self.csphase = nlat, nlon, grid, norm, csphase
        cost, _, lmax_default = _grid_nodes(grid, nlat)
        self.lmax = lmax or lmax_default
        t = np.flip(np.arccos(cost))
        self.mmax = mmax or self.nlon // 2 + 1
        pct = precompute_legpoly(self.mmax - 1, self.lmax - 1, t, norm=self.norm, inverse=True, csphase=self.csphase)
        pct = torch.from_numpy(pct)
        self.register_buffer("pct", pct, persistent=False)

    def extra_repr(self):
        return f"nlat={self.nlat}, nlon={self.nlon},\n lmax={self.lmax}, mmax={self.mmax},\n grid={self.grid}, csphase={self.csphase}"

    def forward(self, x):
        assert x.shape[-2] == self.lmax
        assert x.shape[-1] == self.mmax
        x = torch.view_as_real(x)
        rl = torch.einsum("...lm, mlk->...km", x[..., 0], self.pct.to(x.dtype))
        im = torch.einsum("...lm, mlk->...km", x[..., 1], self.pct.to(x.dtype))
        xs = torch.stack((rl, im), -1)
        x = torch.view_as_complex(xs)
        x = torch.fft.irfft(x, n=self.nlon, dim=-1, norm="forward")
        return x


quadrature = types.SimpleNamespace(
    legendre_gauss_weights=legendre_gauss_weights,
    lobatto_weights=lobatto_weights,
    clenshaw_curtiss_weights=clenshaw_curtiss_weights,
)
legendre = types.SimpleNamespace(legpoly=legpoly, precompute_legpoly=precompute_legpoly)
