"""Tiny driver for `ncu --set full`: runs each hot-path stage a few times at the BASELINE shapes."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import msfno_b200
from msfno_b200 import _lib

dev = torch.device("cuda:0")
B, C, L, M = 1, 256, 120, 121
prec = sys.argv[1] if len(sys.argv) > 1 else "tf32"
with torch.no_grad():
    for nlat, nlon, grid in ((721, 1440, "equiangular"), (120, 240, "legendre-gauss")):
        s = msfno_b200.RealSHT(nlat, nlon, lmax=L, mmax=M, grid=grid).float().to(dev)
        i = msfno_b200.InverseRealSHT(nlat, nlon, lmax=L, mmax=M, grid=grid).float().to(dev)
        s.weights = s.weights * 1e5
        i.pct = i.pct / 1e5
        x = torch.randn(B, C, nlat, nlon, device=dev)
        for _ in range(3):
            pm = s.forward_packed(x)
            cm = msfno_b200.sht.relayout(pm, s, _lib.LAYOUT_PM, _lib.LAYOUT_CM, B, C)
            y = i.inverse_packed(cm)
    att = msfno_b200.SpectralAttentionS2(s, i, C, hidden_size_factor=2, spectral_layers=3, precision=prec).to(dev)
    for _ in range(3):
        att.spectral(pm)
torch.cuda.synchronize()
print("ok")
