"""GPU tests of the serving helpers: CUDA-graph replay (GraphedForward, rollout) and the host-buffer pipeline
(HostPipeline) must reproduce the eager forward exactly (same kernels, same order)."""
import os

import pytest
import torch

from conftest import GOLD, TOL_FP32, rel_l2
from oracle import sfno_oracle

pytestmark = pytest.mark.gpu

import msfno_b200


def _small_net(ftype="non-linear"):
    d = torch.load(os.path.join(GOLD, "net_nonlinear_small.pt" if ftype == "non-linear" else "net_linear_small.pt"))
    cfg = d["cfg"]
    sd = sfno_oracle.make_state_dict(filter_type=ftype, img_size=cfg["img_size"], scale_factor=cfg["scale_factor"],
                                     in_chans=cfg["in_chans"], out_chans=cfg["out_chans"], embed=cfg["embed_dim_sfno"],
                                     num_layers=cfg["num_layers"], mlp_ratio=cfg["mlp_ratio"],
                                     spectral_layers=cfg["spectral_layers"], seed=d["seed"])
    net = msfno_b200.FourierNeuralOperatorNet("cuda", None, **cfg)
    full = dict(net.state_dict())
    full.update(sd)
    net.load_state_dict(full, strict=True)
    return net.cuda().eval(), d


@pytest.mark.parametrize("ftype", ["non-linear", "linear"])
def test_graphed_forward_matches_eager_and_reference(ftype):
    net, d = _small_net(ftype)
    x = d["x"].cuda()
    with torch.no_grad():
        y_eager = net(x).clone()
    g = msfno_b200.GraphedForward(net, x)
    y1 = g(x).clone()
    y2 = g(x * 2.0).clone()           # new input through the static buffer
    y3 = g(x).clone()
    assert torch.equal(y1, y_eager) and torch.equal(y3, y_eager)
    assert not torch.equal(y2, y_eager)
    assert rel_l2(y1, d["y"]) < TOL_FP32


def test_rollout_matches_step_by_step():
    net, d = _small_net()
    x = d["x"].cuda()
    with torch.no_grad():
        ref = x
        for _ in range(4):
            ref = net(ref)                      # x <- net(x): in_chans == out_chans (sfno/model.py:327-331)
    g = msfno_b200.GraphedForward(net, x)
    out = g.rollout(x, 4)
    assert torch.equal(out, ref)


def test_host_pipeline_matches_eager():
    net, d = _small_net()
    gen = torch.Generator().manual_seed(3)
    xs = [torch.randn(d["x"].shape, generator=gen).pin_memory() for _ in range(5)]
    ys = [torch.empty(d["y"].shape).pin_memory() for _ in range(5)]
    with torch.no_grad():
        want = [net(x.cuda()).cpu() for x in xs]
    pipe = msfno_b200.HostPipeline(net, torch.device("cuda"))
    pipe.run(xs, ys)
    torch.cuda.synchronize()
    for a, b in zip(ys, want):
        assert torch.equal(a, b)
    # graph-replayed step inside the pipeline
    pipe2 = msfno_b200.HostPipeline(msfno_b200.GraphedForward(net, xs[0].cuda()), torch.device("cuda"))
    ys2 = [torch.empty(d["y"].shape).pin_memory() for _ in range(5)]
    pipe2.run(xs, ys2)
    torch.cuda.synchronize()
    for a, b in zip(ys2, want):
        assert torch.equal(a, b)
