import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import msfno_b200
from msfno_b200 import precision as P, _lib
from oracle import th_shim, sfno_oracle
def rel(a, b): return float((a.double().cpu() - b.double().cpu()).norm() / b.double().cpu().norm())
g = torch.Generator().manual_seed(0)
for (nlat, nlon, grid) in ((721, 1440, "equiangular"), (120, 240, "legendre-gauss")):
    o_s = th_shim.RealSHT(nlat, nlon, lmax=120, mmax=121, grid=grid).float(); o_s.weights = o_s.weights * 1e5
    o_i = th_shim.InverseRealSHT(nlat, nlon, lmax=120, mmax=121, grid=grid).float(); o_i.pct = o_i.pct / 1e5
    x = torch.randn(1, 8, nlat, nlon, generator=g)
    cin = torch.view_as_complex(torch.randn(1, 8, 120, 121, 2, generator=g)) * 1e3
    want, want_y = o_s(x), o_i(cin)
    for mode in ("fp32", "tf32", "tf32_legfp32", "fp32"):
        msfno_b200.set_precision(mode[:4]); P.set_legendre_on_tensor_cores(mode == "tf32")
        sht = msfno_b200.RealSHT(nlat, nlon, lmax=120, mmax=121, grid=grid).float().cuda(); sht.weights = sht.weights * 1e5
        isht = msfno_b200.InverseRealSHT(nlat, nlon, lmax=120, mmax=121, grid=grid).float().cuda(); isht.pct = isht.pct / 1e5
        with torch.no_grad():
            e1 = rel(torch.view_as_real(sht(x.cuda())), torch.view_as_real(want))
            e2 = rel(isht(cin.cuda()), want_y)
            # same module, toggled tier afterwards
            msfno_b200.set_precision("fp32"); P.set_legendre_on_tensor_cores(True)
            e3 = rel(torch.view_as_real(sht(x.cuda())), torch.view_as_real(want))
        print(nlat, mode, "sht %.2e isht %.2e  after-toggle-to-fp32 sht %.2e" % (e1, e2, e3), flush=True)
