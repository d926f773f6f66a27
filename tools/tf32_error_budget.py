#!/usr/bin/env python
"""Which stages carry the tf32 tier's error on the full 12-block net?  Runs the net against the CPU oracle with single
stage classes switched to the fp32-grade engine, prints rel-L2 per variant (JSON).  usage: python tools/tf32_error_budget.py [--perturb]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import msfno_b200  # noqa: E402
from msfno_b200 import _lib, conv as mconv, precision as mprec, sfnonet as msf  # noqa: E402
from test_gpu_parity_r2 import _oracle_and_net  # noqa: E402
from conftest import rel_l2  # noqa: E402


def main():
    perturb = "--perturb" in sys.argv
    torch.set_num_threads(os.cpu_count() or 1)
    x, want, net = _oracle_and_net(21 if perturb else 0, 1, perturb=perturb)
    xc = x.cuda()
    out = {}

    def run(name):
        with torch.no_grad():
            out[name] = rel_l2(net(xc), want)

    msfno_b200.set_precision("fp32")
    run("all_fp32")
    msfno_b200.set_precision("tf32")
    run("all_tf32")
    # (i) spectral MLP at fp32 grade
    for b in net.blocks:
        b.filter_layer.filter.precision = "fp32"
    run("tf32_but_spectral_mlp_fp32")
    for b in net.blocks:
        b.filter_layer.filter.precision = None
    # (ii) Legendre at fp32 grade
    mprec.set_legendre_on_tensor_cores(False)
    run("tf32_but_legendre_fp32")
    mprec.set_legendre_on_tensor_cores(True)
    # (iii) channel MLPs / convs at fp32 grade (fused kernel off, conv1x1 with the fp32 engine, no TF32 rounding of weights)
    keep = (mconv.mlp1x1_supported, msf.mlp1x1_supported, mconv.conv1x1, msf.conv1x1, mconv.padded_weight, msf.padded_weight,
            msf.fold_norm_affine, msf.fold_affine)

    def conv_fp32(*a, **k):
        mprec._TIER = "fp32"
        try:
            return keep[2](*a, **k)
        finally:
            mprec._TIER = "tf32"

    def wrap_fp32(fn):
        def f(*a, **k):
            mprec._TIER = "fp32"
            try:
                return fn(*a, **k)
            finally:
                mprec._TIER = "tf32"
        return f

    msf.mlp1x1_supported = mconv.mlp1x1_supported = lambda *a, **k: False
    msf.conv1x1 = conv_fp32
    msf.padded_weight = wrap_fp32(keep[4])
    msf.fold_norm_affine = wrap_fp32(keep[6])
    msf.fold_affine = wrap_fp32(keep[7])
    _lib.invalidate_caches()
    run("tf32_but_channel_mlps_fp32")
    (mconv.mlp1x1_supported, msf.mlp1x1_supported, mconv.conv1x1, msf.conv1x1, mconv.padded_weight, msf.padded_weight,
     msf.fold_norm_affine, msf.fold_affine) = keep
    _lib.invalidate_caches()
    run("all_tf32_again")
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
