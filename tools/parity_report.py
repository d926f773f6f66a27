"""Measured rel-L2 of the CUDA path vs the oracle at BASELINE sizes, both tiers (development aid -> profiles/)."""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import msfno_b200
from oracle import sfno_oracle

def rel(a, b):
    return float((a.double().cpu() - b.double()).norm() / b.double().norm())

out = {}
sd = sfno_oracle.make_state_dict(filter_type="non-linear", seed=0)
tr = sfno_oracle.Transforms()
g = torch.Generator().manual_seed(3)
x = torch.randn(1, 73, 721, 1440, generator=g)
t0 = time.time()
with torch.no_grad():
    want = sfno_oracle.sfno_forward(x, sd, tr, "non-linear", 12)
out["oracle_seconds"] = time.time() - t0
from msfno_b200 import precision as _prec
for tier in ("fp32", "tf32", "tf32_legendre_fp32"):
    msfno_b200.set_precision(tier[:4])
    _prec.set_legendre_on_tensor_cores(tier == "tf32")
    net = msfno_b200.FourierNeuralOperatorNet("cuda", None, filter_type="non-linear")
    full = dict(net.state_dict()); full.update(sd); net.load_state_dict(full, strict=True)
    net = net.cuda().eval()
    with torch.no_grad():
        got = net(x.cuda())
        torch.cuda.synchronize(); t1 = time.time()
        for _ in range(5): net(x.cuda())
        torch.cuda.synchronize(); out["eager_ms_%s" % tier] = (time.time() - t1) / 5 * 1e3
    out["full_sfno12_nonlinear_%s_rel_l2" % tier] = rel(got, want)
    del net
print(json.dumps(out, indent=1))
