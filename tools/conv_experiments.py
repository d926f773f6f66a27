import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import msfno_b200
from msfno_b200.conv import conv1x1, padded_weight
msfno_b200.set_precision("tf32")
dev = torch.device("cuda:0")
def timeit(fn, iters=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]
out = {}
def run(name, B, cin, cout, H, W, gelu=True):
    x = torch.randn(B, cin, H, W, device=dev); w = torch.randn(cout, cin, 1, 1, device=dev) / cin ** 0.5
    b = torch.randn(cout, device=dev); wp = padded_weight(w)
    out[name] = timeit(lambda: conv1x1(x, wp, cin, bias=b, act_gelu=gelu))
run("full_256_256_gelu", 1, 256, 256, 721, 1440)
run("full_256_256_nogelu", 1, 256, 256, 721, 1440, gelu=False)
run("E1_full_32_256", 1, 32, 256, 721, 1440)
run("E2_full_256_128", 1, 256, 128, 721, 1440)
run("E3_36x_inner_256_256", 36, 256, 256, 120, 240)
run("E4_full_64_256", 1, 64, 256, 721, 1440)
run("E5_full_128_256", 1, 128, 256, 721, 1440)
y = torch.empty(1, 256, 721, 1440, device=dev); x = torch.randn(1, 256, 721, 1440, device=dev)
out["copy_1GB"] = timeit(lambda: y.copy_(x))
print(json.dumps(out, indent=1))
