import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import msfno_b200
from msfno_b200 import precision as P
from oracle import sfno_oracle
def rel(a, b): return float((a.double().cpu() - b.double().cpu()).norm() / b.double().cpu().norm())
order = sys.argv[1].split(",")
d = torch.load("tests/golden/net_nonlinear_small.pt"); cfg = d["cfg"]
sd = sfno_oracle.make_state_dict(filter_type="non-linear", img_size=cfg["img_size"], scale_factor=cfg["scale_factor"], in_chans=cfg["in_chans"], out_chans=cfg["out_chans"], embed=cfg["embed_dim_sfno"], num_layers=cfg["num_layers"], mlp_ratio=cfg["mlp_ratio"], spectral_layers=cfg["spectral_layers"], seed=d["seed"])
sdF = sfno_oracle.make_state_dict(filter_type="non-linear", seed=0, num_layers=2)
trF = sfno_oracle.Transforms()
xF = torch.randn(1, 73, 721, 1440, generator=torch.Generator().manual_seed(3))
with torch.no_grad(): wantF = sfno_oracle.sfno_forward(xF, sdF, trF, "non-linear", 2)
for mode in order:
    msfno_b200.set_precision(mode[:4]); P.set_legendre_on_tensor_cores(mode != "tf32_legfp32")
    net = msfno_b200.FourierNeuralOperatorNet("cuda", None, **cfg); full = dict(net.state_dict()); full.update(sd); net.load_state_dict(full, strict=True); net = net.cuda().eval()
    with torch.no_grad(): e = rel(net(d["x"].cuda()), d["y"])
    netF = msfno_b200.FourierNeuralOperatorNet("cuda", None, filter_type="non-linear", num_layers=2); full = dict(netF.state_dict()); full.update(sdF); netF.load_state_dict(full, strict=True); netF = netF.cuda().eval()
    with torch.no_grad(): eF = rel(netF(xF.cuda()), wantF)
    print(mode, "small %.2e  full-size 2-block %.2e" % (e, eF), flush=True)
    del net, netF
