"""GPU parity: SpectralConvS2 / SpectralAttentionS2 / FiLM (forward + backward) against the golden vectors
produced by the unmodified reference, and against the oracle at larger sizes."""
import os

import pytest
import torch

from conftest import GOLD, TOL_FP32, rel_l2
from oracle import sfno_oracle, th_shim

pytestmark = pytest.mark.gpu

import msfno_b200


def _transforms(nlat=24, nlon=48, L=12, M=13, grid="equiangular"):
    sht = msfno_b200.RealSHT(nlat, nlon, lmax=L, mmax=M, grid=grid).float().cuda()
    isht = msfno_b200.InverseRealSHT(nlat, nlon, lmax=L, mmax=M, grid=grid).float().cuda()
    sht.weights = sht.weights * 1e5
    isht.pct = isht.pct / 1e5
    return sht, isht


def _oracle_transforms(nlat, nlon, L, M, grid):
    sht = th_shim.RealSHT(nlat, nlon, lmax=L, mmax=M, grid=grid).float()
    isht = th_shim.InverseRealSHT(nlat, nlon, lmax=L, mmax=M, grid=grid).float()
    sht.weights = sht.weights * 1e5
    isht.pct = isht.pct / 1e5
    return sht, isht


def test_spectral_conv_golden_fwd_bwd():
    d = torch.load(os.path.join(GOLD, "filter_linear_24x48.pt"))
    sht, isht = _transforms()
    mod = msfno_b200.SpectralConvS2(sht, isht, 8, use_complex_kernels=True).cuda()
    with torch.no_grad():
        mod.w.copy_(d["w"])
    x = d["x"].cuda().requires_grad_(True)
    y = mod(x)
    assert rel_l2(y, d["y"]) < TOL_FP32
    y.backward(d["gy"].cuda())
    assert rel_l2(x.grad, d["gx"]) < TOL_FP32
    assert rel_l2(mod.w.grad, d["gw"]) < TOL_FP32


def test_spectral_attention_golden_fwd_bwd():
    d = torch.load(os.path.join(GOLD, "filter_nonlinear_24x48.pt"))
    sht, isht = _transforms()
    mod = msfno_b200.SpectralAttentionS2(sht, isht, 8, use_complex_kernels=True, hidden_size_factor=2,
                                         complex_activation="real", spectral_layers=3, bias=False).cuda()
    with torch.no_grad():
        for p, w in zip(mod.w, d["ws"]):
            p.copy_(w)
        mod.wout.copy_(d["wout"])
    x = d["x"].cuda().requires_grad_(True)
    y = mod(x)
    assert rel_l2(y, d["y"]) < TOL_FP32
    y.backward(d["gy"].cuda())
    assert rel_l2(x.grad, d["gx"]) < TOL_FP32
    assert rel_l2(mod.wout.grad, d["gwout"]) < TOL_FP32
    for p, gw in zip(mod.w, d["gws"]):
        assert rel_l2(p.grad, gw) < TOL_FP32


@pytest.mark.parametrize("B,C", [(1, 32), (2, 16), (3, 8), (5, 8)])
def test_spectral_conv_inner_grid_vs_oracle(B, C):
    """120x240 Legendre-Gauss grid, lmax=120, mmax=121 (n = 7260 modes): every batch-tile variant."""
    o_s, o_i = _oracle_transforms(120, 240, 120, 121, "legendre-gauss")
    sht, isht = _transforms(120, 240, 120, 121, "legendre-gauss")
    g = torch.Generator().manual_seed(B * 100 + C)
    x = torch.randn(B, C, 120, 240, generator=g)
    w = 0.02 * torch.randn(C, C, 7260, 2, generator=g)
    want = sfno_oracle.spectral_conv_s2(x, w, o_s, o_i)
    mod = msfno_b200.SpectralConvS2(sht, isht, C, use_complex_kernels=True).cuda()
    with torch.no_grad():
        mod.w.copy_(w)
        got = mod(x.cuda())
    assert rel_l2(got, want) < TOL_FP32


def test_spectral_attention_inner_grid_vs_oracle():
    o_s, o_i = _oracle_transforms(120, 240, 120, 121, "legendre-gauss")
    sht, isht = _transforms(120, 240, 120, 121, "legendre-gauss")
    g = torch.Generator().manual_seed(7)
    C = 32
    x = torch.randn(2, C, 120, 240, generator=g)
    ws = [0.2 * torch.randn(C, 2 * C, 2, generator=g), 0.2 * torch.randn(2 * C, 2 * C, 2, generator=g)]
    wout = 0.2 * torch.randn(2 * C, C, 2, generator=g)
    want = sfno_oracle.spectral_attention_s2(x, ws, wout, o_s, o_i)
    mod = msfno_b200.SpectralAttentionS2(sht, isht, C, use_complex_kernels=True, hidden_size_factor=2,
                                         spectral_layers=2).cuda()
    with torch.no_grad():
        for p, w in zip(mod.w, ws):
            p.copy_(w)
        mod.wout.copy_(wout)
        got = mod(x.cuda())
    assert rel_l2(got, want) < TOL_FP32
    # reference-compatible forward_mlp entry point on the standard complex layout
    c = o_s(x)
    want_c = sfno_oracle.spectral_attention_mlp(c, ws, wout)
    with torch.no_grad():
        got_c = mod.forward_mlp(torch.view_as_real(c).cuda())
    ii, jj = torch.tril_indices(120, 121)
    assert rel_l2(got_c[:, :, ii, jj], torch.view_as_real(want_c)[:, :, ii, jj]) < TOL_FP32


def test_film_golden_fwd_bwd():
    d = torch.load(os.path.join(GOLD, "film_small.pt"))
    x = d["x"].cuda().requires_grad_(True)
    gam = d["gamma"].cuda().requires_grad_(True)
    bet = d["beta"].cuda().requires_grad_(True)
    y = msfno_b200.FiLM()(x, gam, bet, d["scale"])
    assert rel_l2(y, d["y"]) < 1e-6
    gy = torch.randn(d["y"].shape, generator=torch.Generator().manual_seed(0))
    y.backward(gy.cuda())
    xo, go, bo = (d[k].clone().requires_grad_(True) for k in ("x", "gamma", "beta"))
    sfno_oracle.film(xo, go, bo, d["scale"]).backward(gy)
    assert rel_l2(x.grad, xo.grad) < 1e-6 and rel_l2(gam.grad, go.grad) < 1e-5 and rel_l2(bet.grad, bo.grad) < 1e-5


def test_norm_film_folding():
    from msfno_b200.sfnonet import norm_film_coeffs, plane_affine, plane_stats
    g = torch.Generator().manual_seed(9)
    B, C = 2, 5
    x = torch.randn(B, C, 37, 72, generator=g) * 3 + 1
    norm = torch.nn.InstanceNorm2d(C, eps=1e-6, affine=True)
    with torch.no_grad():
        norm.weight.copy_(torch.randn(C, generator=g))
        norm.bias.copy_(torch.randn(C, generator=g))
    gam, bet = torch.randn(B, C, generator=g), torch.randn(B, C, generator=g)
    want = sfno_oracle.film(sfno_oracle.instance_norm(x, norm.weight, norm.bias), gam, bet, 0.4)
    xc = x.cuda()
    A, S = norm_film_coeffs(plane_stats(xc), norm.cuda(), B, C, 37 * 72, gam.cuda(), bet.cuda(), 0.4)
    assert rel_l2(plane_affine(xc, A, S), want) < TOL_FP32


@pytest.mark.parametrize("M,N,K", [(1, 1, 1), (7, 5, 3), (128, 128, 16), (130, 70, 721), (300, 1030, 517)])
def test_gemm_nt(M, N, K):
    from msfno_b200._lib import lib, check, ptr
    g = torch.Generator().manual_seed(M + N + K)
    A, Bm = torch.randn(M, K, generator=g).cuda(), torch.randn(N, K, generator=g).cuda()
    D = torch.empty(M, N, device="cuda")
    check(lib.msfno_gemm_nt(ptr(A), K, ptr(Bm), K, ptr(D), N, M, N, K, 0, 0, torch.cuda.current_stream().cuda_stream))
    want = A.double() @ Bm.double().T
    assert rel_l2(D, want) < 1e-6


def test_specconv_c_abi_paths_agree():
    """msfno_specconv_fwd / bwd_x / bwd_w through the C ABI: the TMA-ring kernels (workspace given) and the
    register-load fallback (ws = NULL) against the complex einsum of contractions.py:37-41 and its adjoints."""
    from msfno_b200._lib import check, lib, ptr
    from msfno_b200.sht import relayout
    from msfno_b200 import _lib
    B, C = 3, 8
    sht = msfno_b200.RealSHT(120, 240, lmax=120, mmax=121, grid="legendre-gauss").float().cuda()
    plan = sht._get_plan(torch.device("cuda:0"))
    g = torch.Generator().manual_seed(11)
    a = torch.randn(B, C, 120, 121, dtype=torch.complex64, generator=g)
    gy = torch.randn(B, C, 120, 121, dtype=torch.complex64, generator=g)
    ii, jj = torch.tril_indices(120, 121)
    a, gy = a * 0, gy * 0 + 0 * a  # keep shapes; fill only the tril modes (the packed layouts store l >= m only)
    a[..., ii, jj] = torch.randn(B, C, len(ii), dtype=torch.complex64, generator=g)
    gy[..., ii, jj] = torch.randn(B, C, len(ii), dtype=torch.complex64, generator=g)
    w = torch.randn(C, C, len(ii), 2, generator=g)
    wc = torch.view_as_complex(w)
    want = torch.einsum("bin,kin->bkn", a[..., ii, jj].to(torch.complex128), wc.to(torch.complex128))
    want_ga = torch.einsum("bkn,kin->bin", gy[..., ii, jj].to(torch.complex128), wc.conj().to(torch.complex128))
    want_gw = torch.einsum("bkn,bin->kin", gy[..., ii, jj].to(torch.complex128), a[..., ii, jj].conj().to(torch.complex128))
    st = torch.cuda.current_stream().cuda_stream
    a_pm = relayout(torch.view_as_real(a).contiguous().cuda(), sht, _lib.LAYOUT_STD, _lib.LAYOUT_PM, B, C)
    g_pm = relayout(torch.view_as_real(gy).contiguous().cuda(), sht, _lib.LAYOUT_STD, _lib.LAYOUT_PM, B, C)
    wd = w.cuda()
    for use_ws in (True, False):
        ws = torch.empty(lib.msfno_specconv_ws_floats(plan.h, B, C, C), device="cuda") if use_ws else None
        out = torch.full((B, plan.P, 2 * C), float("nan"), device="cuda")
        ga = torch.full((B, plan.P, 2 * C), float("nan"), device="cuda")
        gw = torch.full_like(wd, float("nan"))
        check(lib.msfno_specconv_fwd(plan.h, ptr(a_pm), ptr(wd), ptr(out), ptr(ws), B, C, C, st))
        check(lib.msfno_specconv_bwd_x(plan.h, ptr(g_pm), ptr(wd), ptr(ga), ptr(ws), B, C, C, st))
        check(lib.msfno_specconv_bwd_w(plan.h, ptr(a_pm), ptr(g_pm), ptr(gw), ptr(ws), B, C, C, st))
        got = torch.view_as_complex(relayout(out, sht, _lib.LAYOUT_PM, _lib.LAYOUT_STD, B, C))[..., ii, jj]
        got_ga = torch.view_as_complex(relayout(ga, sht, _lib.LAYOUT_PM, _lib.LAYOUT_STD, B, C))[..., ii, jj]
        assert torch.isfinite(out).all() and torch.isfinite(ga).all() and torch.isfinite(gw).all()
        assert rel_l2(torch.view_as_real(got), torch.view_as_real(want)) < TOL_FP32
        assert rel_l2(torch.view_as_real(got_ga), torch.view_as_real(want_ga)) < TOL_FP32
        assert rel_l2(gw, torch.view_as_real(want_gw)) < TOL_FP32
