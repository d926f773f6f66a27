import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import msfno_b200
from msfno_b200.conv import conv1x1, padded_weight
msfno_b200.set_precision("tf32")
dev = torch.device("cuda:0")
x = torch.randn(1, 256, 721, 1440, device=dev); w = torch.randn(256, 256, 1, 1, device=dev) / 16; b = torch.randn(256, device=dev)
wp = padded_weight(w)
for _ in range(3):
    y = conv1x1(x, wp, 256, bias=b, act_gelu=True)
torch.cuda.synchronize(); print("ok")
