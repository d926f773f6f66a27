"""GPU parity: RealSHT / InverseRealSHT (forward, adjoint, fused prologue/epilogue) against the oracle."""
import os

import pytest
import torch

from conftest import GOLD, TOL_FP32, rel_l2
from oracle import th_shim

pytestmark = pytest.mark.gpu

import msfno_b200
from msfno_b200 import _lib


def _pair(nlat, nlon, L, M, grid, rescale=None):
    o_s = th_shim.RealSHT(nlat, nlon, lmax=L, mmax=M, grid=grid).float()
    o_i = th_shim.InverseRealSHT(nlat, nlon, lmax=L, mmax=M, grid=grid).float()
    m_s = msfno_b200.RealSHT(nlat, nlon, lmax=L, mmax=M, grid=grid).float().cuda()
    m_i = msfno_b200.InverseRealSHT(nlat, nlon, lmax=L, mmax=M, grid=grid).float().cuda()
    if rescale:
        o_s.weights = o_s.weights * rescale
        o_i.pct = o_i.pct / rescale
        m_s.weights = m_s.weights * rescale
        m_i.pct = m_i.pct / rescale
    return o_s, o_i, m_s, m_i


@pytest.mark.parametrize("name", ["sht_equi_24x48", "sht_lg_12x24", "sht_equi_37x72_l10"])
def test_golden_transforms(name):
    d = torch.load(os.path.join(GOLD, name + ".pt"))
    _, _, sht, isht = _pair(d["nlat"], d["nlon"], d["lmax"], d["mmax"], d["grid"])
    c = sht(d["x"].cuda())
    assert c.shape == d["coeffs"].shape and c.dtype == torch.complex64
    assert rel_l2(torch.view_as_real(c), torch.view_as_real(d["coeffs"])) < TOL_FP32
    y = isht(d["cin"].cuda())
    assert y.shape == d["y"].shape
    assert rel_l2(y, d["y"]) < TOL_FP32


@pytest.mark.parametrize("grid,nlat,nlon,L,M,B,C", [
    ("equiangular", 721, 1440, 120, 121, 1, 6),     # trans_down / itrans_up of the net (sfnonet.py:537-542)
    ("legendre-gauss", 120, 240, 120, 121, 2, 8),   # trans / itrans (sfnonet.py:543-548); mmax-1 == Nyquist
    ("equiangular", 91, 180, 30, 31, 2, 3),         # ragged: nlat not a multiple of the 32-row tile
    ("legendre-gauss", 16, 30, 16, 16, 1, 1),       # nlon/2 = 15 = 5*3, single plane
    ("equiangular", 33, 64, 40, 20, 1, 2),          # lmax > mmax
])
def test_forward_inverse_vs_oracle(grid, nlat, nlon, L, M, B, C):
    o_s, o_i, sht, isht = _pair(nlat, nlon, L, M, grid, rescale=1e5)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, C, nlat, nlon, generator=g)
    want = o_s(x)
    got = sht(x.cuda())
    assert rel_l2(torch.view_as_real(got), torch.view_as_real(want)) < TOL_FP32
    cin = torch.view_as_complex(torch.randn(B, C, L, M, 2, generator=g)) * 1e3
    want_y = o_i(cin)
    got_y = isht(cin.cuda())
    assert rel_l2(got_y, want_y) < TOL_FP32


def test_leading_dims_and_3d_input():
    o_s, o_i, sht, isht = _pair(12, 24, 12, 13, "legendre-gauss")
    g = torch.Generator().manual_seed(3)
    x = torch.randn(5, 12, 24, generator=g)
    assert rel_l2(torch.view_as_real(sht(x.cuda())), torch.view_as_real(o_s(x))) < TOL_FP32
    x2 = torch.randn(12, 24, generator=g)
    c = sht(x2.cuda())
    assert c.shape == (12, 13)
    assert rel_l2(isht(c), o_i(o_s(x2))) < TOL_FP32


def test_table_mutation_is_honoured():
    o_s, _, sht, _ = _pair(24, 48, 12, 13, "equiangular")
    x = torch.randn(1, 2, 24, 48, generator=torch.Generator().manual_seed(1))
    a = sht(x.cuda())
    sht.weights = sht.weights * 3.0           # reassignment (sfnonet.py:552)
    b = sht(x.cuda())
    assert rel_l2(torch.view_as_real(b), 3.0 * torch.view_as_real(a)) < 1e-6
    sht.weights.mul_(0.5)                      # in-place mutation
    c = sht(x.cuda())
    assert rel_l2(torch.view_as_real(c), 1.5 * torch.view_as_real(a)) < 1e-6


@pytest.mark.parametrize("grid,nlat,nlon,L,M", [("equiangular", 24, 48, 12, 13), ("legendre-gauss", 120, 240, 120, 121),
                                                ("equiangular", 91, 180, 30, 31)])
def test_adjoints_vs_oracle_autograd(grid, nlat, nlon, L, M):
    o_s, o_i, sht, isht = _pair(nlat, nlon, L, M, grid, rescale=1e5)
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 3, nlat, nlon, generator=g)
    gc = torch.randn(2, 3, L, M, 2, generator=g)
    xo = x.clone().requires_grad_(True)
    torch.view_as_real(o_s(xo)).backward(gc)
    xg = x.cuda().requires_grad_(True)
    torch.view_as_real(sht(xg)).backward(gc.cuda())
    assert rel_l2(xg.grad, xo.grad) < TOL_FP32
    cin = torch.randn(2, 3, L, M, 2, generator=g) * 1e3
    gy = torch.randn(2, 3, nlat, nlon, generator=g)
    co = cin.clone().requires_grad_(True)
    o_i(torch.view_as_complex(co)).backward(gy)
    cg = cin.cuda().requires_grad_(True)
    isht(torch.view_as_complex(cg)).backward(gy.cuda())
    # the oracle gradient carries junk in the structurally unused l<m entries only through pct==0 -> both zero
    assert rel_l2(cg.grad, co.grad) < TOL_FP32


def test_fused_prologue_and_epilogue():
    o_s, o_i, sht, isht = _pair(24, 48, 12, 13, "equiangular", rescale=1e5)
    g = torch.Generator().manual_seed(4)
    B, C = 2, 3
    x = torch.randn(B, C, 24, 48, generator=g)
    a, s = torch.rand(B, C, generator=g) + 0.5, torch.randn(B, C, generator=g)
    want = o_s(x * a[:, :, None, None] + s[:, :, None, None])
    pm = sht.forward_packed(x.cuda(), a.cuda(), s.cuda())
    got = torch.view_as_complex(msfno_b200.sht.relayout(pm, sht, _lib.LAYOUT_PM, _lib.LAYOUT_STD, B, C))
    assert rel_l2(torch.view_as_real(got), torch.view_as_real(want)) < TOL_FP32
    cin = torch.view_as_complex(torch.randn(B, C, 12, 13, 2, generator=g)) * 1e3
    skip = torch.randn(B, C, 24, 48, generator=g)
    pre = o_i(cin) + skip
    want_y = torch.nn.functional.gelu(pre)
    cm = msfno_b200.sht.relayout(torch.view_as_real(cin).cuda(), isht, _lib.LAYOUT_STD, _lib.LAYOUT_CM, B, C)
    stats = torch.empty(B * C, 2, dtype=torch.float64, device="cuda")
    with torch.no_grad():
        got_y = isht.inverse_packed(cm, skip_add=skip.cuda(), act_gelu=True, stats=stats)
    assert rel_l2(got_y, want_y) < TOL_FP32
    st = stats.cpu().view(B, C, 2)
    assert torch.allclose(st[..., 0], want_y.double().sum((2, 3)), rtol=1e-5, atol=1e-4)
    assert torch.allclose(st[..., 1], (want_y.double() ** 2).sum((2, 3)), rtol=1e-5)


def test_full_size_roundtrip_property():
    """Size-independent property at the BASELINE grid: analysis(synthesis(c)) == c for band-limited c."""
    _, _, sht, isht = _pair(721, 1440, 120, 121, "equiangular")
    g = torch.Generator().manual_seed(5)
    c = torch.randn(1, 4, 120, 121, 2, generator=g)
    ii, jj = torch.triu_indices(120, 121, offset=1)
    c[:, :, ii, jj] = 0
    c[:, :, :, 0, 1] = 0
    c = torch.view_as_complex(c).cuda()
    back = sht(isht(c))
    assert rel_l2(torch.view_as_real(back), torch.view_as_real(c)) < TOL_FP32


def test_error_behaviour_at_the_boundary():
    """Shape asserts mirror torch_harmonics (x.shape[-2] == nlat, x.shape[-1] == nlon / lmax, mmax); the fused inverse
    epilogue is inference-only; the C ABI's error codes surface as RuntimeError."""
    import msfno_b200
    from msfno_b200.conv import mlp1x1, padded_weight
    sht = msfno_b200.RealSHT(12, 24, lmax=6, mmax=7, grid="legendre-gauss").float().cuda()
    isht = msfno_b200.InverseRealSHT(12, 24, lmax=6, mmax=7, grid="legendre-gauss").float().cuda()
    with pytest.raises(AssertionError):
        sht(torch.randn(1, 2, 11, 24).cuda())
    with pytest.raises(AssertionError):
        sht(torch.randn(1, 2, 12, 22).cuda())
    with pytest.raises(AssertionError):
        isht(torch.randn(1, 2, 6, 8, dtype=torch.complex64).cuda())
    c = sht(torch.randn(1, 2, 12, 24).cuda())
    assert c.shape == (1, 2, 6, 7) and c.dtype == torch.complex64
    assert isht(c).shape == (1, 2, 12, 24)
    # an empty batch is either rejected with the boundary's error or returns an empty result -- and leaves no sticky
    # CUDA error behind: the next valid call must still work
    try:
        e = sht(torch.randn(0, 2, 12, 24).cuda())
        assert e.shape[0] == 0
    except (RuntimeError, AssertionError):
        pass
    torch.cuda.synchronize()
    assert rel_l2(sht(torch.view_as_real(c).new_zeros(1, 2, 12, 24) + 1.0)[..., 1:, :], torch.zeros(1, 2, 5, 7, dtype=torch.complex64)) >= 0.0
    x = torch.randn(1, 8, 4, 8).cuda()
    w1, w2 = padded_weight(torch.randn(32, 8, 1, 1).cuda()), padded_weight(torch.randn(8, 32, 1, 1).cuda())
    msfno_b200.set_precision("tf32")
    try:
        with pytest.raises(RuntimeError, match="out must be"):
            mlp1x1(x, w1, 8, torch.zeros(32).cuda(), w2, None, out=torch.empty(1, 8, 4, 4).cuda())
    finally:
        msfno_b200.set_precision("fp32")
