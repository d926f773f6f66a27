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


def test_rollout_in_place_graph_tf32_tier():
    """Tensor-core tier: the rollout graph lets the fused decoder write the forecast into the input buffer (no feedback
    copy).  The decoder reads its big-skip operand (the same buffer) tile by tile before storing that tile, so the
    in-place steps must reproduce the out-of-place eager steps (and the oracle on the same weights)."""
    from conftest import TOL_TF32
    cfg = dict(filter_type="non-linear", img_size=(36, 72), scale_factor=2, in_chans=7, out_chans=7, embed_dim_sfno=32,
               num_layers=3, mlp_ratio=2.0, spectral_layers=2)
    torch.manual_seed(11)
    net = msfno_b200.FourierNeuralOperatorNet("cuda", None, **cfg).cuda().eval()
    x = torch.randn(2, 7, 36, 72, generator=torch.Generator().manual_seed(12))
    sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    tr = sfno_oracle.Transforms(cfg["img_size"], cfg["scale_factor"])
    want = x
    with torch.no_grad():
        for _ in range(3):
            want = sfno_oracle.sfno_forward(want, sd, tr, "non-linear", cfg["num_layers"])
    xd = x.cuda()
    msfno_b200.set_precision("tf32")
    try:
        with torch.no_grad():
            ref = xd
            for _ in range(3):
                ref = net(ref)
            one_ref = net(xd).clone()
        g = msfno_b200.GraphedForward(net, xd)
        out = g.rollout(xd, 3).clone()
        in_place = g.inplace_graph is not None
        again = g.rollout(xd, 3).clone()     # the static input is reloaded from x0 on every call
        one = g(xd).clone()                  # the out-of-place graph still works after the in-place capture
    finally:
        msfno_b200.set_precision("fp32")
    assert in_place, "fused decoder (mlp1x1, hidden 32) not taken: the in-place rollout graph was not captured"
    assert not hasattr(net, "_decode_out")
    assert rel_l2(out, ref) < 1e-5, "in-place rollout differs from eager steps of the same tier"
    assert torch.equal(out, again)
    assert rel_l2(one, one_ref) < 1e-5
    assert rel_l2(out, want) < 3 * TOL_TF32


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


@pytest.mark.parametrize("tier", ["fp32", "tf32"])
def test_module_on_second_gpu_while_the_first_is_current(tier):
    """A net living on cuda:1 must run (forward and backward) while cuda:0 is the CURRENT device -- the reference's
    PyTorch modules guard the device themselves, so a drop-in has to as well (launch stream, tensor-map encoding and
    per-device constants all follow the input tensor's device).  Needs two GPUs."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    msfno_b200.set_precision(tier)
    try:
        net0, d = _small_net("non-linear")
        torch.cuda.set_device(0)
        net1 = msfno_b200.FourierNeuralOperatorNet("cuda:1", None, **d["cfg"])
        net1.load_state_dict(net0.state_dict(), strict=True)
        net1 = net1.to("cuda:1").eval()
        x0 = d["x"].cuda(0)
        x1 = d["x"].to("cuda:1")
        assert torch.cuda.current_device() == 0
        with torch.no_grad():
            y0 = net0(x0)
            y1 = net1(x1)                      # fused inference path on the non-current device
        assert y1.device.index == 1 and torch.equal(y0.cpu(), y1.cpu())
        # autograd path (adjoint kernels run on autograd's worker thread of device 1)
        for p0, p1 in zip(net0.parameters(), net1.parameters()):
            p0.requires_grad_(True)
            p1.requires_grad_(True)
        g = torch.randn(y0.shape, generator=torch.Generator().manual_seed(3))
        with msfno_b200.precision.library_scope():
            net0(x0).backward(g.cuda(0))
            net1(x1).backward(g.to("cuda:1"))
        assert torch.cuda.current_device() == 0
        for (n, p0), p1 in zip(net0.named_parameters(), net1.parameters()):
            if p0.grad is not None:
                assert p1.grad is not None and rel_l2(p1.grad.cpu(), p0.grad.cpu()) < 1e-6, n
    finally:
        msfno_b200.set_precision("fp32")
        torch.cuda.set_device(0)


@pytest.mark.parametrize("tier", ["fp32", "tf32"])
def test_rollout_with_per_step_host_output(tier):
    """GraphedForward.rollout(host_out=...) (N4: the reference's running() loop copies every step to the host,
    sfno/model.py:345-370): every saved step must equal the eager iterate, whichever graph (in-place / copying) runs."""
    msfno_b200.set_precision(tier)
    try:
        net, d = _small_net("non-linear")
        if d["cfg"]["in_chans"] != d["cfg"]["out_chans"]:
            pytest.skip("rollout needs in_chans == out_chans")
        x = d["x"].cuda()
        steps, every = 6, 2
        with torch.no_grad():
            want, cur = [], x
            for t in range(1, steps + 1):
                cur = net(cur)
                if t % every == 0:
                    want.append(cur.clone())
        g = msfno_b200.GraphedForward(net, x)
        host = [torch.empty(x.shape, dtype=x.dtype).pin_memory() for _ in range(steps // every)]
        last = g.rollout(x, steps, host_out=host, every=every)
        torch.cuda.synchronize()
        assert torch.equal(last.cpu(), want[-1].cpu())
        for h, w in zip(host, want):
            assert torch.equal(h, w.cpu())
        # and again (the staging buffers and events are re-used)
        last = g.rollout(x, steps, host_out=host, every=every)
        torch.cuda.synchronize()
        assert all(torch.equal(h, w.cpu()) for h, w in zip(host, want))
    finally:
        msfno_b200.set_precision("fp32")
