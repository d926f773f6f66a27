#!/usr/bin/env python
"""Times the three GEMM engines (FFMA, plain TF32, 3xTF32) on the shapes of the spectral MLP, the Legendre stages and the
1x1 convs; prints one JSON object.  usage: python tools/x3_bench.py > gpurun_out/x3_bench.json"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import msfno_b200  # noqa: E402
from msfno_b200._lib import check, lib, ptr  # noqa: E402


def timed(fn, iters=10):
    for _ in range(3):
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
    st = torch.cuda.current_stream().cuda_stream
    out = {}
    for (M, N, K, a_k, b_k) in [(7440, 1024, 512, 1, 1), (7440, 1024, 1024, 1, 1), (512, 7440, 1024, 1, 1), (7440, 1024, 1024, 1, 0),
                                (7440, 512, 1024, 0, 0), (1024, 1024, 7440, 0, 0), (256, 28800, 256, 1, 0), (256, 1038240, 76, 1, 0),
                                (256, 1038240, 256, 1, 0)]:
        A = torch.randn((M, K) if a_k else (K, M), device="cuda")
        B = torch.randn((N, K) if b_k else (K, N), device="cuda")
        D = torch.empty(M, N, device="cuda")
        row = {}
        for name, eng in (("ffma", 0), ("tf32", 1), ("x3", 3)):
            if eng == 0 and M * N * K > 3e11:
                continue
            ms = timed(lambda: check(lib.msfno_gemm_ex(ptr(A), A.shape[1], a_k, ptr(B), B.shape[1], b_k, ptr(D), N, M, N, K, 0, None, 0,
                                                        0, eng, st)))
            row[name] = {"ms": round(ms, 4), "TFLOPs": round(2.0 * M * N * K / ms / 1e9, 1)}
        out["M%d_N%d_K%d_a%d_b%d" % (M, N, K, a_k, b_k)] = row
        del A, B, D
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
