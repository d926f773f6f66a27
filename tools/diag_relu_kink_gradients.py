"""diagnostic: SpectralAttentionS2 fwd+bwd at C=256 on the inner grid and with the full-grid inverse, both fp32 engines, vs oracle autograd"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import msfno_b200
from oracle import sfno_oracle, th_shim
from conftest import rel_l2

def case(nlat_o, nlon_o, grid_o, B, C):
    L, M = 120, 121
    o_s = th_shim.RealSHT(120, 240, lmax=L, mmax=M, grid="legendre-gauss").float()
    o_i = th_shim.InverseRealSHT(nlat_o, nlon_o, lmax=L, mmax=M, grid=grid_o).float()
    sht = msfno_b200.RealSHT(120, 240, lmax=L, mmax=M, grid="legendre-gauss").float().cuda()
    isht = msfno_b200.InverseRealSHT(nlat_o, nlon_o, lmax=L, mmax=M, grid=grid_o).float().cuda()
    for t in (o_s, sht): t.weights = t.weights * 1e5
    for t in (o_i, isht): t.pct = t.pct / 1e5
    g = torch.Generator().manual_seed(4)
    x = torch.randn(B, C, 120, 240, generator=g)
    gy = torch.randn(B, C, nlat_o, nlon_o, generator=g)
    ws = [0.02 * torch.randn(C, 2 * C, 2, generator=g), 0.02 * torch.randn(2 * C, 2 * C, 2, generator=g), 0.02 * torch.randn(2 * C, 2 * C, 2, generator=g)]
    wout = 0.02 * torch.randn(2 * C, C, 2, generator=g)
    xo = x.clone().requires_grad_(True)
    wso = [w.clone().requires_grad_(True) for w in ws]
    wouto = wout.clone().requires_grad_(True)
    yo = sfno_oracle.spectral_attention_s2(xo, wso, wouto, o_s, o_i)
    yo.backward(gy)
    for engine in ("tc3x", "ffma"):
        msfno_b200.set_fp32_engine(engine)
        mod = msfno_b200.SpectralAttentionS2(sht, isht, C, hidden_size_factor=2, spectral_layers=3).cuda()
        with torch.no_grad():
            for p, w in zip(mod.w, ws): p.copy_(w)
            mod.wout.copy_(wout)
        xg = x.cuda().requires_grad_(True)
        y = mod(xg)
        y.backward(gy.cuda())
        print(nlat_o, B, C, engine, dict(y=rel_l2(y, yo), gx=rel_l2(xg.grad, xo.grad), gw0=rel_l2(mod.w[0].grad, wso[0].grad),
              gw1=rel_l2(mod.w[1].grad, wso[1].grad), gw2=rel_l2(mod.w[2].grad, wso[2].grad), gwout=rel_l2(mod.wout.grad, wouto.grad)), flush=True)
    msfno_b200.set_fp32_engine("tc3x")

case(120, 240, "legendre-gauss", 2, 256)
case(721, 1440, "equiangular", 1, 256)
case(120, 240, "legendre-gauss", 1, 64)
