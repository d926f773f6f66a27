"""TF32 GEMM throughput on the spectral-MLP shapes (msfno_gemm_nt). MSFNO_GEMM_NO_PAIR=1 selects the single-CTA kernel."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import msfno_b200
from msfno_b200._lib import lib, ptr, check, PREC_TF32
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(192 * 1024 * 1024 // 4, device=dev)
out = {"pair_kernel": os.environ.get("MSFNO_GEMM_NO_PAIR") is None}
for (M, N, K) in [(7440, 1024, 512), (7440, 1024, 1024), (7440, 512, 1024), (14880, 1024, 1024)]:
    A = torch.randn(M, K, device=dev); Bm = torch.randn(N, K, device=dev); D = torch.empty(M, N, device=dev)
    fn = lambda: check(lib.msfno_gemm_nt(ptr(A), K, ptr(Bm), K, ptr(D), N, M, N, K, 1, PREC_TF32, st))
    for _ in range(3): fn()
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    out["%dx%dx%d" % (M, N, K)] = {"ms": round(ms, 4), "TFLOPs": round(2.0 * M * N * K / ms / 1e9, 1)}
print(json.dumps(out))
