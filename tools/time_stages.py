"""Stage-level CUDA-event timings of the hot path at the BASELINE shapes (development aid; bench.py is the
contract).  Usage: python tools/time_stages.py [--C 256] [--B 1] [--linear]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import msfno_b200
from msfno_b200 import _lib


def timeit(fn, iters=5, warm=2, flush=None):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--C", type=int, default=256)
    ap.add_argument("--B", type=int, default=1)
    ap.add_argument("--linear", action="store_true")
    ap.add_argument("--precision", default="fp32")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    msfno_b200.set_precision(a.precision)
    B, C = a.B, a.C
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
    out = {}
    grids = {"full": (721, 1440, "equiangular"), "inner": (120, 240, "legendre-gauss")}
    L, M = 120, 121
    tr = {}
    for name, (nlat, nlon, grid) in grids.items():
        s = msfno_b200.RealSHT(nlat, nlon, lmax=L, mmax=M, grid=grid).float().to(dev)
        i = msfno_b200.InverseRealSHT(nlat, nlon, lmax=L, mmax=M, grid=grid).float().to(dev)
        s.weights = s.weights * 1e5
        i.pct = i.pct / 1e5
        tr[name] = (s, i)
    with torch.no_grad():
        for name, (nlat, nlon, grid) in grids.items():
            s, i = tr[name]
            x = torch.randn(B, C, nlat, nlon, device=dev)
            pm = s.forward_packed(x)
            cm = msfno_b200.sht.relayout(pm, s, _lib.LAYOUT_PM, _lib.LAYOUT_CM, B, C)
            out["sht_fwd_" + name] = timeit(lambda: s.forward_packed(x), flush=flush)
            out["isht_fwd_" + name] = timeit(lambda: i.inverse_packed(cm), flush=flush)
            gb = (4 * B * C * nlat * nlon + 8 * B * C * L * M + 4 * M * L * nlat) / 1e9
            out["sht_fwd_%s_GBps" % name] = gb / out["sht_fwd_" + name] * 1e3
            out["isht_fwd_%s_GBps" % name] = gb / out["isht_fwd_" + name] * 1e3
        s, i = tr["inner"]
        x = torch.randn(B, C, 120, 240, device=dev)
        pm = s.forward_packed(x)
        att = msfno_b200.SpectralAttentionS2(s, i, C, hidden_size_factor=2, spectral_layers=3, precision=a.precision).to(dev)
        out["specattn_mlp"] = timeit(lambda: att.spectral(pm), flush=flush)
        out["specattn_mlp_TFLOPs_dense_equiv"] = 8 * 786432 * (C / 256) ** 2 * B * 7260 / out["specattn_mlp"] / 1e9
        out["specattn_filter_inner"] = timeit(lambda: att(x), flush=flush)
        if a.linear:
            conv = msfno_b200.SpectralConvS2(s, i, C).to(dev)
            out["specconv"] = timeit(lambda: conv.spectral(pm), flush=flush)
            out["specconv_GBps"] = (8 * C * C * 7260 + 16 * B * C * 7260) / 1e9 / out["specconv"] * 1e3
            out["specconv_filter_inner"] = timeit(lambda: conv(x), flush=flush)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
