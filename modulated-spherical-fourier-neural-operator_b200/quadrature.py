"""Host-side (fp64 numpy) quadrature rules for the latitude grids of the SHT.

Mirrors the public helpers of `torch_harmonics.quadrature` that the reference calls
(/root/reference MSFNO/Models/losses.py:90,129: `legendre_gauss_weights(n, -1, 1)[1]`;
the transforms use them through the `grid=` argument, sfnonet.py:537-548).
Nodes are returned ascending in x = cos(theta) on [a, b].
"""
import numpy as np


def legendre_gauss_weights(n, a=-1.0, b=1.0):
    x, w = np.polynomial.legendre.leggauss(n)
    return 0.5 * (b - a) * x + 0.5 * (b + a), 0.5 * (b - a) * w


def clenshaw_curtiss_weights(n, a=-1.0, b=1.0):
    """Clenshaw-Curtis rule on the n equi-angular nodes including both poles, by the closed-form
    cosine series  w_k = c_k/N * (1 - sum_j b_j/(4 j^2 - 1) cos(2 j k pi / N)),  N = n - 1."""
    assert n > 1
    N = n - 1
    k = np.arange(n)
    x = np.cos(np.pi * (N - k) / N)  # ascending
    if n == 2:
        w = np.array([1.0, 1.0])
    else:
        j = np.arange(1, N // 2 + 1)
        bj = np.where(2 * j == N, 1.0, 2.0)
        series = (bj / (4.0 * j * j - 1.0))[None, :] * np.cos(2.0 * np.pi * np.outer(k, j) / N)
        ck = np.where((k == 0) | (k == N), 1.0, 2.0)
        w = ck / N * (1.0 - series.sum(axis=1))
    return 0.5 * (b - a) * x + 0.5 * (b + a), 0.5 * (b - a) * w


def lobatto_weights(n, a=-1.0, b=1.0, tol=1e-16, maxiter=100):
    """Gauss-Lobatto rule by Newton iteration on P_{n-1}."""
    x = -np.cos(np.pi * np.arange(n) / (n - 1))
    P = np.zeros((n, n))
    for _ in range(maxiter):
        xo = x
        P[:, 0] = 1.0
        P[:, 1] = x
        for k in range(2, n):
            P[:, k] = ((2 * k - 1) * x * P[:, k - 1] - (k - 1) * P[:, k - 2]) / k
        x = xo - (x * P[:, n - 1] - P[:, n - 2]) / (n * P[:, n - 1])
        if np.max(np.abs(x - xo)) < tol:
            break
    w = 2.0 / (n * (n - 1) * P[:, n - 1] ** 2)
    return 0.5 * (b - a) * x + 0.5 * (b + a), 0.5 * (b - a) * w


def grid_nodes(grid, nlat):
    """(cos(theta) ascending, weights, default lmax) for a named latitude grid."""
    if grid == "legendre-gauss":
        x, w = legendre_gauss_weights(nlat)
        return x, w, nlat
    if grid == "equiangular":
        x, w = clenshaw_curtiss_weights(nlat)
        return x, w, nlat
    if grid == "lobatto":
        x, w = lobatto_weights(nlat)
        return x, w, nlat - 1
    raise ValueError("Unknown quadrature mode")
