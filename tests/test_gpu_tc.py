"""GPU parity of the tensor-core (TF32, tcgen05/TMEM/TMA) tier: <= 2e-3 rel-L2 (north_star)."""
import pytest
import torch

from conftest import TOL_TF32, rel_l2
from oracle import sfno_oracle, th_shim

pytestmark = pytest.mark.gpu

import msfno_b200
from msfno_b200._lib import PREC_TF32, check, lib, ptr


@pytest.mark.parametrize("M,N,K,relu", [(128, 128, 32, 0), (128, 128, 64, 0), (128, 128, 256, 1), (256, 384, 1024, 0),
                                         (7440, 1024, 1024, 1), (130, 72, 736, 0), (5, 8, 12, 0), (1000, 136, 100, 1)])
def test_gemm_nt_tf32(M, N, K, relu):
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    A, Bm = torch.randn(M, K, generator=g).cuda(), torch.randn(N, K, generator=g).cuda()
    D = torch.full((M, N), float("nan"), device="cuda")
    check(lib.msfno_gemm_nt(ptr(A), K, ptr(Bm), K, ptr(D), N, M, N, K, relu, PREC_TF32, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    want = A.double() @ Bm.double().T
    if relu:
        want[:, 0::2] = want[:, 0::2].clamp_min(0)
    assert torch.isfinite(D).all()
    err = rel_l2(D, want)
    assert err < TOL_TF32, err


def test_gemm_nt_tf32_strided_rows():
    """leading dimensions larger than the logical extents (sub-matrices of bigger buffers)."""
    g = torch.Generator().manual_seed(5)
    M, N, K, lda, ldb, ldd = 200, 160, 96, 128, 104, 192
    A, Bm = torch.randn(M, lda, generator=g).cuda(), torch.randn(N, ldb, generator=g).cuda()
    D = torch.zeros(M, ldd, device="cuda")
    check(lib.msfno_gemm_nt(ptr(A), lda, ptr(Bm), ldb, ptr(D), ldd, M, N, K, 0, PREC_TF32, torch.cuda.current_stream().cuda_stream))
    want = A[:, :K].double() @ Bm[:, :K].double().T
    assert rel_l2(D[:, :N], want) < TOL_TF32
    assert float(D[:, N:].abs().max()) == 0.0


@pytest.mark.parametrize("B,C", [(1, 64), (2, 32)])
def test_spectral_attention_tf32_tier(B, C):
    o_s = th_shim.RealSHT(120, 240, lmax=120, mmax=121, grid="legendre-gauss").float()
    o_i = th_shim.InverseRealSHT(120, 240, lmax=120, mmax=121, grid="legendre-gauss").float()
    sht = msfno_b200.RealSHT(120, 240, lmax=120, mmax=121, grid="legendre-gauss").float().cuda()
    isht = msfno_b200.InverseRealSHT(120, 240, lmax=120, mmax=121, grid="legendre-gauss").float().cuda()
    for t in (o_s, sht):
        t.weights = t.weights * 1e5
    for t in (o_i, isht):
        t.pct = t.pct / 1e5
    g = torch.Generator().manual_seed(B + C)
    x = torch.randn(B, C, 120, 240, generator=g)
    ws = [0.1 * torch.randn(C, 2 * C, 2, generator=g), 0.1 * torch.randn(2 * C, 2 * C, 2, generator=g),
          0.1 * torch.randn(2 * C, 2 * C, 2, generator=g)]
    wout = 0.1 * torch.randn(2 * C, C, 2, generator=g)
    want = sfno_oracle.spectral_attention_s2(x, ws, wout, o_s, o_i)
    mod = msfno_b200.SpectralAttentionS2(sht, isht, C, hidden_size_factor=2, spectral_layers=3, precision="tf32").cuda()
    with torch.no_grad():
        for p, w in zip(mod.w, ws):
            p.copy_(w)
        mod.wout.copy_(wout)
        got = mod(x.cuda())
    err = rel_l2(got, want)
    assert err < TOL_TF32, err
