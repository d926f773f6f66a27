"""GPU parity of the tensor-core (TF32, tcgen05/TMEM/TMA) tier: <= 2e-3 rel-L2 (north_star)."""
import pytest
import torch

from conftest import TOL_TF32, rel_l2
from oracle import sfno_oracle, th_shim

pytestmark = pytest.mark.gpu

import msfno_b200
from msfno_b200._lib import PREC_TF32, check, lib, ptr


@pytest.mark.parametrize("M,N,K,relu", [(128, 128, 32, 0), (128, 128, 64, 0), (128, 128, 256, 1), (256, 384, 1024, 0),
                                         (7440, 1024, 1024, 1), (130, 72, 736, 0), (5, 8, 12, 0), (1000, 136, 100, 1),
                                         # CTA-pair kernel (M >= 1024, N >= 256): exact, ragged and single-k-block shapes
                                         (1024, 256, 32, 0), (1100, 520, 72, 1), (2048, 512, 512, 0), (7440, 512, 1024, 0)])
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


@pytest.fixture
def tf32_tier():
    msfno_b200.set_precision("tf32")
    yield
    msfno_b200.set_precision("fp32")


@pytest.mark.parametrize("grid,nlat,nlon,L,M,B,C", [("equiangular", 721, 1440, 120, 121, 1, 4),
                                                    ("legendre-gauss", 120, 240, 120, 121, 2, 8),
                                                    ("equiangular", 91, 180, 30, 31, 1, 3)])
def test_sht_tf32_tier(tf32_tier, grid, nlat, nlon, L, M, B, C):
    o_s = th_shim.RealSHT(nlat, nlon, lmax=L, mmax=M, grid=grid).float()
    o_i = th_shim.InverseRealSHT(nlat, nlon, lmax=L, mmax=M, grid=grid).float()
    sht = msfno_b200.RealSHT(nlat, nlon, lmax=L, mmax=M, grid=grid).float().cuda()
    isht = msfno_b200.InverseRealSHT(nlat, nlon, lmax=L, mmax=M, grid=grid).float().cuda()
    for t in (o_s, sht):
        t.weights = t.weights * 1e5
    for t in (o_i, isht):
        t.pct = t.pct / 1e5
    g = torch.Generator().manual_seed(11)
    x = torch.randn(B, C, nlat, nlon, generator=g)
    with torch.no_grad():
        got = sht(x.cuda())
    assert rel_l2(torch.view_as_real(got), torch.view_as_real(o_s(x))) < TOL_TF32
    cin = torch.view_as_complex(torch.randn(B, C, L, M, 2, generator=g)) * 1e3
    with torch.no_grad():
        got_y = isht(cin.cuda())
    assert rel_l2(got_y, o_i(cin)) < TOL_TF32


@pytest.mark.parametrize("name,ftype", [("net_linear_small.pt", "linear"), ("net_nonlinear_small.pt", "non-linear")])
def test_small_net_tf32_tier(tf32_tier, name, ftype):
    import os
    from conftest import GOLD
    d = torch.load(os.path.join(GOLD, name))
    cfg = d["cfg"]
    sd = sfno_oracle.make_state_dict(filter_type=ftype, img_size=cfg["img_size"], scale_factor=cfg["scale_factor"],
                                     in_chans=cfg["in_chans"], out_chans=cfg["out_chans"], embed=cfg["embed_dim_sfno"],
                                     num_layers=cfg["num_layers"], mlp_ratio=cfg["mlp_ratio"],
                                     spectral_layers=cfg["spectral_layers"], seed=d["seed"])
    net = msfno_b200.FourierNeuralOperatorNet("cuda", None, **cfg)
    full = dict(net.state_dict())
    full.update(sd)
    net.load_state_dict(full, strict=True)
    net = net.cuda().eval()
    with torch.no_grad():
        y = net(d["x"].cuda())
    assert rel_l2(y, d["y"]) < TOL_TF32


@pytest.mark.parametrize("tier,tol", [("fp32", 1e-5), ("tf32", TOL_TF32)])
@pytest.mark.parametrize("B,Cin,Cout,H,W,Cin2", [(1, 73, 256, 37, 72, 0), (2, 256, 512, 12, 24, 0), (1, 256, 73, 31, 52, 0),
                                                   (2, 64, 96, 16, 40, 73), (1, 8, 8, 4, 4, 0), (1, 256, 256, 120, 240, 73),
                                                   (2, 512, 256, 12, 24, 0), (1, 1000, 130, 120, 240, 0)])
def test_conv1x1_fused(tier, tol, B, Cin, Cout, H, W, Cin2):
    """y = gelu(conv(x, w) [+ conv(x2, w2)] + bias) + add against torch fp64 (msfno_conv1x1_fwd, both engines)."""
    from msfno_b200.conv import conv1x1, padded_weight
    msfno_b200.set_precision(tier)
    try:
        g = torch.Generator().manual_seed(B * 1000 + Cin + Cout)
        x = torch.randn(B, Cin, H, W, generator=g).cuda()
        w = (torch.randn(Cout, Cin, 1, 1, generator=g) / Cin ** 0.5).cuda()
        bias = torch.randn(Cout, generator=g).cuda()
        add = torch.randn(B, Cout, H, W, generator=g).cuda()
        x2 = w2 = None
        ref = torch.nn.functional.conv2d(x.double(), w.double())
        if Cin2:
            x2 = torch.randn(B, Cin2, H, W, generator=g).cuda()
            w2 = (torch.randn(Cout, Cin2, 1, 1, generator=g) / Cin2 ** 0.5).cuda()
            ref = ref + torch.nn.functional.conv2d(x2.double(), w2.double())
        ref = torch.nn.functional.gelu(ref + bias.double()[None, :, None, None]) + add.double()
        got = conv1x1(x, padded_weight(w), Cin, bias=bias, act_gelu=True, add=add, x2=x2,
                      w2=padded_weight(w2) if w2 is not None else None, cin2=Cin2)
        assert rel_l2(got, ref) < tol
        # per-sample weights and bias (InstanceNorm/FiLM affine folded into the conv), no activation, broadcast add
        A = (torch.rand(B, Cin, generator=g) + 0.5).cuda()
        wp = padded_weight(w)
        Wb = (wp.unsqueeze(0) * torch.nn.functional.pad(A, (0, wp.shape[1] - Cin)).unsqueeze(1)).contiguous()
        bb = torch.randn(B, Cout, generator=g).cuda()
        add1 = torch.randn(1, Cout, H, W, generator=g).cuda()
        got = conv1x1(x, Wb, Cin, bias=bb, add=add1, per_sample_w=True, per_sample_bias=True)
        ref = torch.nn.functional.conv2d((x * A[:, :, None, None]).double(), w.double()) + bb.double()[:, :, None, None] + add1.double()
        assert rel_l2(got, ref) < tol
    finally:
        msfno_b200.set_precision("fp32")


def _mlp_ref(x, w1, b1, w2, b2, add=None, x2=None, w1b=None):
    """fp64 restatement of Conv2d(1x1) -> GELU(erf) -> Conv2d(1x1) (+ big-skip second operand, + add), layers.py:161-168."""
    B, _, H, W = x.shape
    xd = x.double().flatten(2)
    h = torch.einsum("bhc,bcp->bhp", w1.double() if w1.dim() == 3 else w1.double().expand(B, -1, -1), xd)
    if x2 is not None:
        h = h + torch.einsum("hc,bcp->bhp", w1b.double(), x2.double().flatten(2))
    h = torch.nn.functional.gelu(h + (b1.double() if b1.dim() == 2 else b1.double().expand(B, -1)).unsqueeze(-1))
    y = torch.einsum("oh,bhp->bop", w2.double(), h)
    if b2 is not None:
        y = y + b2.double().view(1, -1, 1)
    y = y.reshape(B, -1, H, W)
    return y + add.double() if add is not None else y


@pytest.mark.parametrize("B,cin,cin2,chid,cout,H,W,with_add,per_sample", [
    (1, 73, 0, 256, 256, 36, 100, True, False),     # encoder form: pos_embed add
    (2, 256, 73, 256, 73, 25, 40, False, True),     # decoder form: big-skip second operand, per-sample folded weights
    (1, 40, 0, 64, 24, 9, 12, False, False),        # small / ragged: one partial 128-pixel tile
    (3, 16, 8, 128, 256, 31, 36, True, False),
    (2, 256, 0, 512, 256, 30, 64, True, True)])     # block MLP form: 512 hidden channels in two TMEM chunks, residual add
def test_mlp1x1_fused_tf32(B, cin, cin2, chid, cout, H, W, with_add, per_sample):
    from msfno_b200.conv import mlp1x1, padded_weight
    msfno_b200.set_precision("tf32")
    try:
        g = torch.Generator().manual_seed(B * 1000 + cin + cout)
        x = torch.randn(B, cin, H, W, generator=g).cuda()
        x2 = torch.randn(B, cin2, H, W, generator=g).cuda() if cin2 else None
        w1 = (torch.randn(B, chid, cin, generator=g) if per_sample else torch.randn(chid, cin, generator=g)).cuda() / cin ** 0.5
        w1b = torch.randn(chid, cin2, generator=g).cuda() / cin ** 0.5 if cin2 else None
        b1 = (torch.randn(B, chid, generator=g) if per_sample else torch.randn(chid, generator=g)).cuda()
        w2 = torch.randn(cout, chid, generator=g).cuda() / chid ** 0.5
        b2 = torch.randn(cout, generator=g).cuda()
        add = torch.randn(B if per_sample else 1, cout, H, W, generator=g).cuda() if with_add else None
        pad = lambda w: torch.nn.functional.pad(w, (0, (-w.shape[-1]) % 4)).contiguous()
        stats = torch.full((B * cout, 2), float("nan"), dtype=torch.float64, device="cuda")
        y = mlp1x1(x, pad(w1), cin, b1.contiguous(), pad(w2), b2, add=add, x2=x2, w1b=pad(w1b) if cin2 else None, cin2=cin2,
                   per_sample_w1=per_sample, per_sample_b1=per_sample, final=True, stats=stats)
        torch.cuda.synchronize()
        # fused plane statistics of the output (what msfno_plane_stats(y) would return)
        yd = y.double().reshape(B * cout, -1)
        assert torch.allclose(stats[:, 0], yd.sum(1), rtol=1e-5, atol=1e-3 * yd.abs().sum(1).max().item() / yd.shape[1] ** 0.5)
        assert torch.allclose(stats[:, 1], (yd * yd).sum(1), rtol=1e-5)
        want = _mlp_ref(x, w1, b1, w2, b2, add, x2, w1b)
        assert torch.isfinite(y).all()
        err = rel_l2(y, want)
        assert err < TOL_TF32, err
    finally:
        msfno_b200.set_precision("fp32")


@pytest.mark.parametrize("nlat,nlon,grid,B,C", [(120, 240, "legendre-gauss", 2, 8), (721, 1440, "equiangular", 1, 3)])
def test_sht_isht_tf32_dft_gemm(nlat, nlon, grid, B, C):
    """TF32 tier: longitude transforms as DFT GEMMs on the tensor cores (dft_tc.cu) + Legendre GEMMs, against the
    oracle's rfft/irfft + einsum (th_shim, torch_harmonics semantics), including the fused affine / skip / GELU / stats."""
    from msfno_b200 import _lib
    from msfno_b200.sht import relayout
    L, M = 120, 121
    o_s = th_shim.RealSHT(nlat, nlon, lmax=L, mmax=M, grid=grid).float()
    o_i = th_shim.InverseRealSHT(nlat, nlon, lmax=L, mmax=M, grid=grid).float()
    s = msfno_b200.RealSHT(nlat, nlon, lmax=L, mmax=M, grid=grid).float().cuda()
    i = msfno_b200.InverseRealSHT(nlat, nlon, lmax=L, mmax=M, grid=grid).float().cuda()
    g = torch.Generator().manual_seed(nlat + C)
    x = torch.randn(B, C, nlat, nlon, generator=g)
    sc, sh = torch.rand(B, C, generator=g) + 0.5, torch.randn(B, C, generator=g)
    msfno_b200.set_precision("tf32")
    try:
        with torch.no_grad():
            pm = s.forward_packed(x.cuda(), sc.cuda(), sh.cuda())
            got = torch.view_as_complex(relayout(pm, s, _lib.LAYOUT_PM, _lib.LAYOUT_STD, B, C))
            want = o_s(x.double() * sc.double()[..., None, None] + sh.double()[..., None, None])
            assert rel_l2(torch.view_as_real(got), torch.view_as_real(want)) < TOL_TF32
            # inverse with the fused epilogue
            coef = o_s(x.double()).to(torch.complex64)
            skip = torch.randn(B, C, nlat, nlon, generator=g)
            cm = relayout(torch.view_as_real(coef).contiguous().cuda(), i, _lib.LAYOUT_STD, _lib.LAYOUT_CM, B, C)
            stats = torch.zeros(B * C, 2, dtype=torch.float64, device="cuda")
            y = i.inverse_packed(cm, skip_add=skip.cuda(), act_gelu=True, stats=stats)
            ref = torch.nn.functional.gelu(o_i(coef.to(torch.complex128)) + skip.double())
            assert torch.isfinite(y).all()
            assert rel_l2(y, ref) < TOL_TF32
            yd = y.double().reshape(B * C, -1)
            assert torch.allclose(stats[:, 0], yd.sum(1), rtol=1e-6, atol=1e-3)
            assert torch.allclose(stats[:, 1], (yd * yd).sum(1), rtol=1e-6)
            # plain inverse (no epilogue)
            y0 = i.inverse_packed(cm)
            assert rel_l2(y0, o_i(coef.to(torch.complex128))) < TOL_TF32
    finally:
        msfno_b200.set_precision("fp32")


@pytest.mark.parametrize("nlat,nlon,B,C,gelu", [(721, 1440, 1, 3, False), (200, 288, 2, 2, True), (130, 320, 1, 5, True),
                                                 (257, 512, 3, 1, False)])
def test_isht_tf32_parity_split_kernel(nlat, nlon, B, C, gelu):
    """Full-grid inverse longitude transform of the TF32 tier (idft_eo_kernel: even / odd orders accumulated separately,
    y[j] = E + O, y[j + nlon/2] = E - O, persistent CTAs, TMA tensor stores) against the oracle's einsum + irfft,
    with the GELU / plane-statistics epilogue, ragged last latitude and column tiles, and several items per CTA;
    MSFNO_DFT_NO_EO's plain product must agree with it to rounding."""
    from msfno_b200 import _lib
    from msfno_b200.sht import relayout
    L, M = 120, 121
    o_s = th_shim.RealSHT(nlat, nlon, lmax=L, mmax=M, grid="equiangular").float()
    o_i = th_shim.InverseRealSHT(nlat, nlon, lmax=L, mmax=M, grid="equiangular").float()
    i = msfno_b200.InverseRealSHT(nlat, nlon, lmax=L, mmax=M, grid="equiangular").float().cuda()
    g = torch.Generator().manual_seed(nlat + nlon)
    x = torch.randn(B, C, nlat, nlon, generator=g)
    msfno_b200.set_precision("tf32")
    try:
        with torch.no_grad():
            coef = o_s(x.double()).to(torch.complex64)
            cm = relayout(torch.view_as_real(coef).contiguous().cuda(), i, _lib.LAYOUT_STD, _lib.LAYOUT_CM, B, C)
            stats = torch.zeros(B * C, 2, dtype=torch.float64, device="cuda")
            y = i.inverse_packed(cm, act_gelu=gelu, stats=stats)
            ref = o_i(coef.to(torch.complex128))
            if gelu:
                ref = torch.nn.functional.gelu(ref)
            assert torch.isfinite(y).all()
            assert rel_l2(y, ref) < TOL_TF32
            # both halves of the longitude circle separately (a wrong sign of the odd part shows up in one of them only)
            assert rel_l2(y[..., : nlon // 2], ref[..., : nlon // 2]) < TOL_TF32
            assert rel_l2(y[..., nlon // 2:], ref[..., nlon // 2:]) < TOL_TF32
            yd = y.double().reshape(B * C, -1)
            assert torch.allclose(stats[:, 0], yd.sum(1), rtol=1e-6, atol=1e-3)
            assert torch.allclose(stats[:, 1], (yd * yd).sum(1), rtol=1e-6)
    finally:
        msfno_b200.set_precision("fp32")
