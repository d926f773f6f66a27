"""One pass over the round's top kernels for `ncu --set full`: fused MLP (encoder / decoder forms), DFT GEMMs on both grids,
the CTA-pair GEMM (spectral MLP) and the SpectralConvS2 TMA stream."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import msfno_b200
from msfno_b200 import _lib
from msfno_b200._lib import lib, ptr, check
from msfno_b200.conv import mlp1x1, round_tf32
msfno_b200.set_precision("tf32")
dev = torch.device("cuda:0")
B, C, L, M = 1, 256, 120, 121
H, W = 721, 1440
pad = lambda w: round_tf32(torch.nn.functional.pad(w, (0, (-w.shape[-1]) % 4)).contiguous())
with torch.no_grad():
    x = torch.randn(1, 73, H, W, device=dev)
    w1, b1 = pad(torch.randn(256, 73, device=dev) / 8), torch.randn(256, device=dev)
    w2, b2 = pad(torch.randn(256, 256, device=dev) / 16), torch.randn(256, device=dev)
    pos = torch.randn(1, 256, H, W, device=dev)
    stats = torch.zeros(256, 2, dtype=torch.float64, device=dev)
    mlp1x1(x, w1, 73, b1, w2, b2, add=pos, stats=stats)                                    # encoder form
    y = torch.randn(1, 256, H, W, device=dev)
    w1a, w1b = pad(torch.randn(256, 256, device=dev) / 16), pad(torch.randn(256, 73, device=dev) / 16)
    w3, b3 = pad(torch.randn(73, 256, device=dev) / 16), torch.randn(73, device=dev)
    mlp1x1(y, w1a, 256, b1, w3, b3, x2=x, w1b=w1b, cin2=73, final=True)                   # decoder form
    del pos
    for nlat, nlon, grid in ((721, 1440, "equiangular"), (120, 240, "legendre-gauss")):
        s = msfno_b200.RealSHT(nlat, nlon, lmax=L, mmax=M, grid=grid).float().to(dev)
        i = msfno_b200.InverseRealSHT(nlat, nlon, lmax=L, mmax=M, grid=grid).float().to(dev)
        xx = torch.randn(B, C, nlat, nlon, device=dev)
        st2 = torch.zeros(B * C, 2, dtype=torch.float64, device=dev)
        pm = s.forward_packed(xx)
        cm = msfno_b200.sht.relayout(pm, s, _lib.LAYOUT_PM, _lib.LAYOUT_CM, B, C)
        i.inverse_packed(cm, act_gelu=True, stats=st2)
    att = msfno_b200.SpectralAttentionS2(s, i, C, hidden_size_factor=2, spectral_layers=3, precision="tf32").to(dev)
    att.spectral(pm)
    plan = s._get_plan(dev)
    stc = torch.cuda.current_stream().cuda_stream
    w = torch.randn(C, C, 7260, 2, device=dev)
    a = torch.randn(1, plan.P, 2 * C, device=dev); o = torch.empty(1, plan.P, 2 * C, device=dev)
    ws = torch.empty(lib.msfno_specconv_ws_floats(plan.h, 1, C, C), device=dev)
    check(lib.msfno_specconv_fwd(plan.h, ptr(a), ptr(w), ptr(o), ptr(ws), 1, C, C, stc))
torch.cuda.synchronize()
print("ok")
