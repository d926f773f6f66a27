"""GPU parity of the fp32 tier's tensor-core engine (3xTF32: three tcgen05 MMAs per k-step on hi / lo operand splits,
csrc/gemm_tc.cu) against fp64, and against the CUDA-core FFMA engine it replaces.  Tolerance: the fp32 tier's 1e-5
rel-L2 (north_star); a single contraction is expected around 1e-6 or better, asserted at 3e-6."""
import pytest
import torch

from conftest import TOL_FP32, rel_l2

pytestmark = pytest.mark.gpu

import msfno_b200
from msfno_b200._lib import check, lib, ptr

FFMA, TF32, X3 = 0, 1, 3
TOL_ONE_GEMM = 3e-6


def _st():
    return torch.cuda.current_stream().cuda_stream


def _gemm_ex(A, a_k, Bm, b_k, M, N, K, engine, relu=0, mask=None, acc=None):
    D = torch.full((M, N), float("nan"), device="cuda") if acc is None else acc.clone()
    check(lib.msfno_gemm_ex(ptr(A), A.shape[1], a_k, ptr(Bm), Bm.shape[1], b_k, ptr(D), N, M, N, K, relu, ptr(mask),
                            N if mask is not None else 0, 0 if acc is None else 1, engine, _st()), "gemm_ex")
    torch.cuda.synchronize()
    return D


@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (128, 128, 256), (256, 384, 1024), (130, 72, 736), (5, 8, 12),
                                   (1000, 136, 100), (300, 200, 1440),
                                   # CTA-pair kernel (fp32 tier: M >= 256, N >= 256, M N >= 2^20)
                                   (2048, 512, 512), (7440, 1024, 1024), (7440, 512, 1024), (512, 7440, 1024), (1100, 1032, 72)])
@pytest.mark.parametrize("a_k,b_k", [(1, 1), (1, 0), (0, 1), (0, 0)])
def test_gemm_x3_all_majornesses(M, N, K, a_k, b_k):
    if (a_k, b_k) != (1, 1) and M * N * K > 2048 * 512 * 512:
        pytest.skip("large shapes only in the K-major form")
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    pad4 = lambda n: (n + 3) // 4 * 4
    A = torch.randn((M, pad4(K)) if a_k else (K, pad4(M)), generator=g).cuda()
    Bm = torch.randn((N, pad4(K)) if b_k else (K, pad4(N)), generator=g).cuda()
    Ad = (A[:, :K] if a_k else A[:, :M].T).double()
    Bd = (Bm[:, :K] if b_k else Bm[:, :N].T).double()
    want = Ad @ Bd.T
    if N % 4:
        pytest.skip("D leading dimension must be a multiple of 4 here")
    got = _gemm_ex(A, a_k, Bm, b_k, M, N, K, X3)
    assert torch.isfinite(got).all()
    err = rel_l2(got, want)
    ffma = rel_l2(_gemm_ex(A, a_k, Bm, b_k, M, N, K, FFMA), want)
    assert err < TOL_ONE_GEMM, (err, ffma)
    # and the plain TF32 MMA through the same kernel (truncated operands): the tensor-core tier's tolerance
    assert rel_l2(_gemm_ex(A, a_k, Bm, b_k, M, N, K, TF32), want) < 2e-3


def test_gemm_x3_relu_mask_accumulate():
    g = torch.Generator().manual_seed(3)
    M, N, K = 700, 264, 520
    A, Bm = torch.randn(M, K, generator=g).cuda(), torch.randn(N, K, generator=g).cuda()
    mask = torch.randn(M, N, generator=g).cuda()
    old = torch.randn(M, N, generator=g).cuda()
    want = A.double() @ Bm.double().T
    w_relu = want.clone()
    w_relu[:, 0::2] = w_relu[:, 0::2].clamp_min(0)
    assert rel_l2(_gemm_ex(A, 1, Bm, 1, M, N, K, X3, relu=1), w_relu) < TOL_ONE_GEMM
    w_mask = want.clone()
    w_mask[:, 0::2] = torch.where(mask[:, 0::2].double() > 0, w_mask[:, 0::2], torch.zeros_like(w_mask[:, 0::2]))
    assert rel_l2(_gemm_ex(A, 1, Bm, 1, M, N, K, X3, mask=mask), w_mask) < TOL_ONE_GEMM
    assert rel_l2(_gemm_ex(A, 1, Bm, 1, M, N, K, X3, acc=old), want + old.double()) < TOL_ONE_GEMM


def test_gemm_x3_wide_dynamic_range():
    """operands spanning many binades (Legendre tables span 1e-6 .. 4): the split must stay exact per element."""
    g = torch.Generator().manual_seed(9)
    M, N, K = 256, 256, 736
    A = (torch.randn(M, K, generator=g) * torch.exp(6 * torch.randn(M, K, generator=g))).cuda()
    Bm = (torch.randn(N, K, generator=g) * torch.exp(6 * torch.randn(N, K, generator=g))).cuda()
    want = A.double() @ Bm.double().T
    assert rel_l2(_gemm_ex(A, 1, Bm, 1, M, N, K, X3), want) < TOL_ONE_GEMM


def test_fp32_engines_agree_on_spectral_filter():
    """SpectralAttentionS2 forward + backward, fp32 tier: tensor-core engine vs FFMA engine vs each other."""
    from oracle import sfno_oracle, th_shim
    nlat, nlon, L, Mm, B, C = 120, 240, 120, 121, 2, 32
    o_s = th_shim.RealSHT(nlat, nlon, lmax=L, mmax=Mm, grid="legendre-gauss").float()
    o_i = th_shim.InverseRealSHT(nlat, nlon, lmax=L, mmax=Mm, grid="legendre-gauss").float()
    sht = msfno_b200.RealSHT(nlat, nlon, lmax=L, mmax=Mm, grid="legendre-gauss").float().cuda()
    isht = msfno_b200.InverseRealSHT(nlat, nlon, lmax=L, mmax=Mm, grid="legendre-gauss").float().cuda()
    for t in (o_s, sht):
        t.weights = t.weights * 1e5
    for t in (o_i, isht):
        t.pct = t.pct / 1e5
    g = torch.Generator().manual_seed(4)
    x = torch.randn(B, C, nlat, nlon, generator=g)
    gy = torch.randn(B, C, nlat, nlon, generator=g)
    ws = [0.1 * torch.randn(C, 2 * C, 2, generator=g), 0.1 * torch.randn(2 * C, 2 * C, 2, generator=g)]
    wout = 0.1 * torch.randn(2 * C, C, 2, generator=g)
    xo = x.clone().requires_grad_(True)
    wso = [w.clone().requires_grad_(True) for w in ws]
    wouto = wout.clone().requires_grad_(True)
    yo = sfno_oracle.spectral_attention_s2(xo, wso, wouto, o_s, o_i)
    yo.backward(gy)
    res = {}
    try:
        for engine in ("tc3x", "ffma"):
            msfno_b200.set_fp32_engine(engine)
            mod = msfno_b200.SpectralAttentionS2(sht, isht, C, hidden_size_factor=2, spectral_layers=2).cuda()
            with torch.no_grad():
                for p, w in zip(mod.w, ws):
                    p.copy_(w)
                mod.wout.copy_(wout)
            xg = x.cuda().requires_grad_(True)
            y = mod(xg)
            y.backward(gy.cuda())
            res[engine] = dict(y=rel_l2(y, yo), gx=rel_l2(xg.grad, xo.grad), gw0=rel_l2(mod.w[0].grad, wso[0].grad),
                               gw1=rel_l2(mod.w[1].grad, wso[1].grad), gwout=rel_l2(mod.wout.grad, wouto.grad))
    finally:
        msfno_b200.set_fp32_engine("tc3x")
    print(res)
    for engine, e in res.items():
        for k, v in e.items():
            assert v < TOL_FP32, (engine, k, v, res)
