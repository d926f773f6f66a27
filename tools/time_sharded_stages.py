#!/usr/bin/env python
"""Per-stage device times of the spatially sharded SHT (BASELINE configs[4] (A): 1441 x 2880, C = 256, lmax 240) on ONE GPU:
longitude FFT, Legendre analysis, Legendre synthesis, inverse FFT -- the four kernels a rank runs either side of the
all-to-all.  usage: python tools/time_sharded_stages.py > gpurun_out/sharded_stages.json"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import msfno_b200  # noqa: E402
from msfno_b200 import distributed as D  # noqa: E402


def timed(fn, iters=5):
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    dev = torch.device("cuda", 0)
    nlat, nlon, L, M, B, C = 1441, 2880, 240, 241, 1, 256
    sht = msfno_b200.RealSHT(nlat, nlon, lmax=L, mmax=M, grid="equiangular").float().to(dev)
    isht = msfno_b200.InverseRealSHT(nlat, nlon, lmax=L, mmax=M, grid="equiangular").float().to(dev)
    st = D.CudaStages(nlat, nlat, nlon, L, M, sht.weights, isht.pct, dev)
    x = torch.randn(B, C, nlat, nlon, device=dev)
    P = st.leg_a.P
    out = {"config": "1441 x 2880, C = 256, lmax = 240, one GPU holds every latitude and every order"}
    with torch.no_grad():
        Xt = st.fft_fwd(x)
        pm = st.legendre_fwd(Xt, 0, st.mlim, P)
        cm = pm.transpose(1, 2).contiguous()
        Yt = st.legendre_inv(cm, 0, st.mlim)
        y = st.fft_inv(Yt, B, C)
        gb_grid = 4.0 * B * C * nlat * nlon / 1e9
        gb_xt = 4.0 * Xt.numel() / 1e9
        ms = timed(lambda: st.fft_fwd(x))
        out["fft_fwd"] = {"ms": ms, "GB": gb_grid + gb_xt, "GBps": (gb_grid + gb_xt) / ms * 1e3}
        ms = timed(lambda: st.legendre_fwd(Xt, 0, st.mlim, P))
        gf = 2.0 * 2 * C * nlat * sum(L - m for m in range(st.mlim)) / 1e9
        out["legendre_analysis"] = {"ms": ms, "GFLOP": gf, "TFLOPs": gf / ms}
        ms = timed(lambda: st.legendre_inv(cm, 0, st.mlim))
        out["legendre_synthesis"] = {"ms": ms, "GFLOP": gf, "TFLOPs": gf / ms}
        ms = timed(lambda: st.fft_inv(Yt, B, C))
        out["fft_inv"] = {"ms": ms, "GB": gb_grid + gb_xt, "GBps": (gb_grid + gb_xt) / ms * 1e3}
        ms = timed(lambda: pm.transpose(1, 2).contiguous())
        out["pm_to_cm_transpose"] = {"ms": ms}
    out["roundtrip_rel_l2_vs_input_band_limited"] = None
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
