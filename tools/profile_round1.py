"""One pass over the kernels added late in round 1 (for `ncu --set full -k regex:...`): DFT GEMMs on both grids and the
SpectralConvS2 TMA stream."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import msfno_b200
from msfno_b200 import _lib
from msfno_b200._lib import lib, ptr, check
msfno_b200.set_precision("tf32")
dev = torch.device("cuda:0")
B, C, L, M = 1, 256, 120, 121
with torch.no_grad():
    for nlat, nlon, grid in ((721, 1440, "equiangular"), (120, 240, "legendre-gauss")):
        s = msfno_b200.RealSHT(nlat, nlon, lmax=L, mmax=M, grid=grid).float().to(dev)
        i = msfno_b200.InverseRealSHT(nlat, nlon, lmax=L, mmax=M, grid=grid).float().to(dev)
        x = torch.randn(B, C, nlat, nlon, device=dev)
        skip = torch.randn(B, C, nlat, nlon, device=dev)
        stats = torch.zeros(B * C, 2, dtype=torch.float64, device=dev)
        for _ in range(2):
            pm = s.forward_packed(x)
            cm = msfno_b200.sht.relayout(pm, s, _lib.LAYOUT_PM, _lib.LAYOUT_CM, B, C)
            y = i.inverse_packed(cm, skip_add=skip, act_gelu=True, stats=stats)
    plan = s._get_plan(dev)
    st = torch.cuda.current_stream().cuda_stream
    w = torch.randn(C, C, 7260, 2, device=dev)
    a = torch.randn(1, plan.P, 2 * C, device=dev); o = torch.empty(1, plan.P, 2 * C, device=dev)
    ws = torch.empty(lib.msfno_specconv_ws_floats(plan.h, 1, C, C), device=dev)
    for _ in range(2):
        check(lib.msfno_specconv_fwd(plan.h, ptr(a), ptr(w), ptr(o), ptr(ws), 1, C, C, st))
torch.cuda.synchronize()
print("ok")
