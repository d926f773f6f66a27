"""Legendre contraction stages alone (tf32 tier): analysis (Xt -> PM) and synthesis (CM -> Yt) on both grids."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import msfno_b200
from msfno_b200 import _lib
from msfno_b200._lib import lib, ptr, check
msfno_b200.set_precision("tf32")
dev = torch.device("cuda:0")
B, C, L, M = 1, 256, 120, 121
flush = torch.empty(192 * 1024 * 1024 // 4, device=dev)
def timeit(fn, iters=8):
    for _ in range(3): fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]
out = {"persistent": os.environ.get("MSFNO_GEMM_NO_PERSIST") is None}
st = torch.cuda.current_stream().cuda_stream
for nlat, nlon, grid in ((120, 240, "legendre-gauss"), (721, 1440, "equiangular")):
    s = msfno_b200.RealSHT(nlat, nlon, lmax=L, mmax=M, grid=grid).float().to(dev)
    i = msfno_b200.InverseRealSHT(nlat, nlon, lmax=L, mmax=M, grid=grid).float().to(dev)
    x = torch.randn(B, C, nlat, nlon, device=dev)
    with torch.no_grad():
        pm = s.forward_packed(x)                      # sets tables / precision on the plans
        cm = msfno_b200.sht.relayout(pm, s, _lib.LAYOUT_PM, _lib.LAYOUT_CM, B, C)
        i.inverse_packed(cm)
    ps, pi = s._get_plan(dev), i._get_plan(dev)
    kpad, mlim, P = lib.msfno_plan_query(ps.h, _lib.Q_KPAD), lib.msfno_plan_query(ps.h, _lib.Q_MLIM), lib.msfno_plan_query(ps.h, _lib.Q_NPACK)
    xt = torch.randn(B, mlim, 2 * C, kpad, device=dev)
    opm = torch.empty(B, P, 2 * C, device=dev)
    out["analysis_%d" % nlat] = timeit(lambda: check(lib.msfno_legendre_stage(ps.h, 0, ptr(xt), ptr(opm), 0, mlim, B, C, st)))
    yt = torch.empty(B, mlim, 2 * C, kpad, device=dev)
    out["synthesis_%d" % nlat] = timeit(lambda: check(lib.msfno_legendre_stage(pi.h, 2, ptr(cm), ptr(yt), 0, mlim, B, C, st)))
print(json.dumps(out))
