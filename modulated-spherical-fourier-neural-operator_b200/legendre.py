"""Host-side (fp64 numpy) orthonormal associated Legendre tables P_l^m(cos theta) with the
Condon-Shortley phase, as `torch_harmonics.legendre` provides them to RealSHT / InverseRealSHT
(SURVEY.md Appendix A.2).  Built once per transform; O(mmax * lmax) vector operations over latitude.
"""
import numpy as np


def legpoly(mmax, lmax, x, norm="ortho", inverse=False, csphase=True):
    """table[m, l, k] for m < mmax, l < lmax at nodes x[k] = cos(theta_k); zero for l < m."""
    x = np.asarray(x, dtype=np.float64)
    tab = np.zeros((mmax, lmax, x.shape[0]), dtype=np.float64)
    nf = 1.0 if norm == "ortho" else np.sqrt(4.0 * np.pi)
    nf = 1.0 / nf if inverse else nf
    s2 = (1.0 + x) * (1.0 - x)  # sin^2(theta)
    pmm = np.full_like(x, nf / np.sqrt(4.0 * np.pi))  # P_0^0
    for m in range(mmax):
        if m > 0:
            pmm = np.sqrt((2 * m + 1) * s2 / (2.0 * m)) * pmm  # P_m^m from P_{m-1}^{m-1}
        if m < lmax:
            tab[m, m] = pmm
        if m + 1 < lmax:
            tab[m, m + 1] = np.sqrt(2 * m + 3.0) * x * pmm
        for l in range(m + 2, lmax):
            a = np.sqrt((2 * l - 1.0) / (l - m) * (2 * l + 1.0) / (l + m))
            b = np.sqrt((l + m - 1.0) / (l - m) * (2 * l + 1.0) / (2 * l - 3.0) * (l - m - 1.0) / (l + m))
            tab[m, l] = a * x * tab[m, l - 1] - b * tab[m, l - 2]
    if norm == "schmidt":
        ls = np.sqrt(2.0 * np.arange(lmax) + 1.0)[None, :, None]
        tab = tab * ls if inverse else tab / ls
    if csphase:
        tab[1::2] *= -1.0
    return tab


def precompute_legpoly(mmax, lmax, theta, norm="ortho", inverse=False, csphase=True):
    return legpoly(mmax, lmax, np.cos(theta), norm=norm, inverse=inverse, csphase=csphase)
