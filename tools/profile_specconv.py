"""One SpectralConvS2 contraction per batch size (for ncu captures): python tools/profile_specconv.py [B ...]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import msfno_b200
from msfno_b200._lib import lib, ptr, check
dev = torch.device("cuda:0")
sht = msfno_b200.RealSHT(120, 240, lmax=120, mmax=121, grid="legendre-gauss").float().to(dev)
plan = sht._get_plan(dev)
st = torch.cuda.current_stream().cuda_stream
C = 256
w = torch.randn(C, C, 7260, 2, device=dev)
for B in [int(v) for v in sys.argv[1:]] or [1]:
    a = torch.randn(B, plan.P, 2 * C, device=dev); o = torch.empty(B, plan.P, 2 * C, device=dev)
    ws = torch.empty(lib.msfno_specconv_ws_floats(plan.h, B, C, C), device=dev)
    for _ in range(2):
        check(lib.msfno_specconv_fwd(plan.h, ptr(a), ptr(w), ptr(o), ptr(ws), B, C, C, st))
    torch.cuda.synchronize()
print("ok")
