"""Encoder / decoder channel MLPs at 721x1440: fused msfno_mlp1x1_fwd vs two msfno_conv1x1_fwd calls (tf32 tier)."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import msfno_b200
from msfno_b200.conv import conv1x1, mlp1x1, round_tf32
msfno_b200.set_precision("tf32")
dev = torch.device("cuda:0")
H, W = 721, 1440
quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
flush = torch.empty(192 * 1024 * 1024 // 4, device=dev)
def timeit(fn, iters=5):
    for _ in range(2): fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]
pad = lambda w: round_tf32(torch.nn.functional.pad(w, (0, (-w.shape[-1]) % 4)).contiguous())
out = {}
with torch.no_grad():
    # encoder: 73 -> 256 -> 256 + pos_embed
    x = torch.randn(1, 73, H, W, device=dev)
    w1, b1 = pad(torch.randn(256, 73, device=dev) / 8), torch.randn(256, device=dev)
    w2, b2 = pad(torch.randn(256, 256, device=dev) / 16), torch.randn(256, device=dev)
    pos = torch.randn(1, 256, H, W, device=dev)
    fused = lambda: mlp1x1(x, w1, 73, b1, w2, b2, add=pos)
    if quick:
        fused(); torch.cuda.synchronize()
    else:
        out["enc_fused_ms"] = timeit(fused)
        out["enc_two_convs_ms"] = timeit(lambda: conv1x1(conv1x1(x, w1, 73, bias=b1, act_gelu=True), w2, 256, bias=b2, add=pos))
        out["enc_fused_no_add_ms"] = timeit(lambda: mlp1x1(x, w1, 73, b1, w2, b2))
        out["enc_GB_min"] = (73 + 256 + 256) * H * W * 4 / 1e9
    del pos
    # decoder: (256 | 73) -> 256 -> 73
    y = torch.randn(1, 256, H, W, device=dev)
    w1a, w1b = pad(torch.randn(256, 256, device=dev) / 16), pad(torch.randn(256, 73, device=dev) / 16)
    w3, b3 = pad(torch.randn(73, 256, device=dev) / 16), torch.randn(73, device=dev)
    fused = lambda: mlp1x1(y, w1a, 256, b1, w3, b3, x2=x, w1b=w1b, cin2=73, final=True)
    if quick:
        fused(); torch.cuda.synchronize()
    else:
        out["dec_fused_ms"] = timeit(fused)
        out["dec_two_convs_ms"] = timeit(lambda: conv1x1(conv1x1(y, w1a, 256, bias=b1, act_gelu=True, x2=x, w2=w1b, cin2=73), w3, 256, bias=b3, final=True))
        out["dec_GB_min"] = (256 + 73 + 73) * H * W * 4 / 1e9
print(json.dumps(out))
