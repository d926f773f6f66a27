"""Event timings of the fused 1x1-conv GEMM at the SFNO shapes vs the PyTorch library conv (development aid)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import msfno_b200
from msfno_b200.conv import conv1x1, padded_weight

def timeit(fn, iters=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]

dev = torch.device("cuda:0"); out = {}
torch.backends.cudnn.allow_tf32 = True
for tier in ("tf32", "fp32"):
    msfno_b200.set_precision(tier)
    for name, (cin, cout, H, W) in {"enc_fc1_73_256_full": (73, 256, 721, 1440), "enc_fc2_256_256_full": (256, 256, 721, 1440),
                                     "dec_fc2_256_73_full": (256, 73, 721, 1440), "mlp_fc1_256_512_inner": (256, 512, 120, 240),
                                     "mlp_fc2_512_256_inner": (512, 256, 120, 240)}.items():
        x = torch.randn(1, cin, H, W, device=dev); w = torch.randn(cout, cin, 1, 1, device=dev) / cin ** 0.5
        b = torch.randn(cout, device=dev); wp = padded_weight(w)
        out["%s_%s_ours_bias_gelu" % (name, tier)] = timeit(lambda: conv1x1(x, wp, cin, bias=b, act_gelu=True))
        if tier == "tf32":
            out["%s_cudnn_tf32_conv_only" % name] = timeit(lambda: torch.nn.functional.conv2d(x, w, b))
            out["%s_cudnn_tf32_conv_gelu" % name] = timeit(lambda: torch.nn.functional.gelu(torch.nn.functional.conv2d(x, w, b)))
        del x
print(json.dumps(out, indent=1))
