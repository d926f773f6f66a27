#!/usr/bin/env python
"""One 3xTF32 GEMM shape a few times (the command ncu profiles): python tools/x3_one.py [M N K a_k b_k engine]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import msfno_b200  # noqa: E402
from msfno_b200._lib import check, lib, ptr  # noqa: E402

a = [int(v) for v in sys.argv[1:]] + [7440, 1024, 1024, 1, 1, 3][len(sys.argv) - 1:]
M, N, K, a_k, b_k, eng = a
A = torch.randn((M, K) if a_k else (K, M), device="cuda")
B = torch.randn((N, K) if b_k else (K, N), device="cuda")
D = torch.empty(M, N, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for _ in range(4):
    check(lib.msfno_gemm_ex(ptr(A), A.shape[1], a_k, ptr(B), B.shape[1], b_k, ptr(D), N, M, N, K, 0, None, 0, 0, eng, st))
torch.cuda.synchronize()
print("ok", float(D.abs().mean()))
