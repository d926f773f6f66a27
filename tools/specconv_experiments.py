import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import msfno_b200
from msfno_b200._lib import lib, ptr, check
dev = torch.device("cuda:0")
sht = msfno_b200.RealSHT(120, 240, lmax=120, mmax=121, grid="legendre-gauss").float().to(dev)
plan = sht._get_plan(dev)
flush = torch.empty(192 * 1024 * 1024 // 4, device=dev)
def timeit(fn, iters=6):
    for _ in range(2): fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]
out = {}
st = torch.cuda.current_stream().cuda_stream
for C in (64, 128, 256):
    for B in (1, 4, 8):
        a = torch.randn(B, plan.P, 2 * C, device=dev); w = torch.randn(C, C, 7260, 2, device=dev); o = torch.empty(B, plan.P, 2 * C, device=dev)
        ws = torch.empty(lib.msfno_specconv_ws_floats(plan.h, B, C, C), device=dev)
        ms = timeit(lambda: check(lib.msfno_specconv_fwd(plan.h, ptr(a), ptr(w), ptr(o), ptr(ws), B, C, C, st)))
        gb = (8.0 * C * C * 7260 + 16.0 * B * C * 7260) / 1e9
        ent = {"ms": ms, "GBps": gb / ms * 1e3, "weight_GB": 8.0 * C * C * 7260 / 1e9}
        g = torch.randn_like(o); ga = torch.empty_like(a); gw = torch.empty_like(w)
        ent["bwd_x_ms"] = timeit(lambda: check(lib.msfno_specconv_bwd_x(plan.h, ptr(g), ptr(w), ptr(ga), ptr(ws), B, C, C, st)))
        ent["bwd_w_ms"] = timeit(lambda: check(lib.msfno_specconv_bwd_w(plan.h, ptr(a), ptr(g), ptr(gw), ptr(ws), B, C, C, st)))
        ent["fallback_ms"] = timeit(lambda: check(lib.msfno_specconv_fwd(plan.h, ptr(a), ptr(w), ptr(o), None, B, C, C, st)))
        out["C%d_B%d" % (C, B)] = ent
        del g, ga, gw
        del a, w, o
x = torch.empty(int(3.8e9 // 4), device=dev); y = torch.empty_like(x)
ms = timeit(lambda: y.copy_(x)); out["copy_3.8GB"] = {"ms": ms, "GBps_rw": 2 * 3.8 / ms * 1e3}
ms = timeit(lambda: x.sum()); out["read_3.8GB_sum"] = {"ms": ms, "GBps": 3.8 / ms * 1e3}
print(json.dumps(out, indent=1))
