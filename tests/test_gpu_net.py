"""GPU parity: whole SFNO / MSFNO nets (fused inference path and autograd path) against the golden outputs of
the unmodified reference, plus training-step gradients against the oracle's autograd."""
import os

import pytest
import torch

from conftest import GOLD, TOL_FP32, rel_l2
from oracle import sfno_oracle

pytestmark = pytest.mark.gpu

import msfno_b200


def _load(net, sd):
    full = dict(net.state_dict())
    for k, v in sd.items():
        assert k in full and full[k].shape == v.shape, k
        full[k] = v
    net.load_state_dict(full, strict=True)
    return net


def _oracle_sd(cfg, ftype, seed, film_layers=0):
    return sfno_oracle.make_state_dict(filter_type=ftype, img_size=cfg["img_size"], scale_factor=cfg["scale_factor"],
                                       in_chans=cfg["in_chans"], out_chans=cfg["out_chans"], embed=cfg["embed_dim_sfno"],
                                       num_layers=cfg["num_layers"], mlp_ratio=cfg["mlp_ratio"],
                                       spectral_layers=cfg["spectral_layers"], seed=seed, film_layers=film_layers)


@pytest.mark.parametrize("name,ftype", [("net_linear_small.pt", "linear"), ("net_nonlinear_small.pt", "non-linear")])
def test_small_net_matches_reference(name, ftype):
    d = torch.load(os.path.join(GOLD, name))
    cfg = d["cfg"]
    net = _load(msfno_b200.FourierNeuralOperatorNet("cuda", None, **cfg), _oracle_sd(cfg, ftype, d["seed"])).cuda().eval()
    x = d["x"].cuda()
    with torch.no_grad():
        y_fused = net(x)
    assert rel_l2(y_fused, d["y"]) < TOL_FP32
    y_auto = net(x.clone().requires_grad_(True))   # autograd (unfused) path
    assert rel_l2(y_auto, d["y"]) < TOL_FP32


class _Cfg:
    film_gen_type, cls, embed_dim, mlp_dim, dropout, scale_weight, repeat_film = "mae", "x", 512, 1024, 0.0, 1, False


@pytest.mark.parametrize("fl", [1, 3])
def test_filmed_net_matches_reference(fl):
    d = torch.load(os.path.join(GOLD, "filmed_fl%d_small.pt" % fl))
    cfg = dict(d["cfg"])
    c = _Cfg()
    c.film_layers, c.batch_size = fl, d["x"].shape[0]
    mlp_ratio = cfg.pop("mlp_ratio")
    net = msfno_b200.FourierNeuralOperatorNet_Filmed("cuda", c, mlp_ratio=mlp_ratio, advanced_logging=True,
                                                     film_layers=fl, model_depth=6, **cfg)
    cfg["mlp_ratio"] = mlp_ratio
    sd = _oracle_sd(cfg, "non-linear", d["seed"], film_layers=fl)
    net = _load(net, sd).cuda().eval()
    x, cond = d["x"].cuda(), d["cond"].cuda()
    with torch.no_grad():
        y = net(x, cond, d["scale"])
    assert rel_l2(net.gamma, d["gamma"]) < 1e-5
    assert rel_l2(y, d["y"]) < TOL_FP32

    # training step: gradients w.r.t. the FiLM head and (fl=3) through SHT / spectral MLP / ISHT
    y2 = net(x, cond, d["scale"])
    assert rel_l2(y2, d["y"]) < TOL_FP32
    gy = torch.randn(d["y"].shape, generator=torch.Generator().manual_seed(1))
    with msfno_b200.precision.library_scope():
        y2.backward(gy.cuda())

    sdo = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
    tr = sfno_oracle.Transforms(cfg["img_size"], cfg["scale_factor"])
    B = x.shape[0]
    fm = sfno_oracle.film_head(d["cond"], sdo).reshape(B, 2, fl, 256)
    # reference semantics: encoder and un-FiLMed blocks run under no_grad (sfnonet.py:817-827,843-844); compare
    # only parameters that receive gradient in both: the film head and the FiLMed blocks' spectral weights
    yo = sfno_oracle.sfno_forward(d["x"], sdo, tr, "non-linear", cfg["num_layers"], film_mod=fm, film_layers=fl,
                                  scale=d["scale"])
    yo.backward(gy)
    got = dict(net.named_parameters())
    for k in ("film_gen.film_head.net.4.weight", "film_gen.film_head.net.1.weight", "decoder.fwd.0.weight"):
        assert rel_l2(got[k].grad, sdo[k].grad) < TOL_FP32, (k, rel_l2(got[k].grad, sdo[k].grad))
    last = cfg["num_layers"] - 1
    k = "blocks.%d.filter_layer.filter.wout" % last
    assert rel_l2(got[k].grad, sdo[k].grad) < TOL_FP32, (k, rel_l2(got[k].grad, sdo[k].grad))
    k = "blocks.%d.filter_layer.filter.w.0" % last
    assert rel_l2(got[k].grad, sdo[k].grad) < TOL_FP32, (k, rel_l2(got[k].grad, sdo[k].grad))


def test_filmed_net_frozen_backbone_training_shortcut():
    """Reference training freezes everything but the FiLM generator (train.py): the FiLMed block then runs its pre-FiLM
    part on the fused no-grad kernels and only FiLM -> decoder is recorded by autograd.  The film-head gradients must
    match the oracle's (which differentiates the whole graph)."""
    fl = 1
    d = torch.load(os.path.join(GOLD, "filmed_fl%d_small.pt" % fl))
    cfg = dict(d["cfg"])
    c = _Cfg()
    c.film_layers, c.batch_size = fl, d["x"].shape[0]
    mlp_ratio = cfg.pop("mlp_ratio")
    net = msfno_b200.FourierNeuralOperatorNet_Filmed("cuda", c, mlp_ratio=mlp_ratio, advanced_logging=True,
                                                     film_layers=fl, model_depth=6, **cfg)
    cfg["mlp_ratio"] = mlp_ratio
    sd = _oracle_sd(cfg, "non-linear", d["seed"], film_layers=fl)
    net = _load(net, sd).cuda().eval()
    for n, p in net.named_parameters():
        p.requires_grad_(n.startswith("film_gen"))
    x, cond = d["x"].cuda(), d["cond"].cuda()
    y = net(x, cond, d["scale"])
    assert rel_l2(y, d["y"]) < TOL_FP32
    gy = torch.randn(d["y"].shape, generator=torch.Generator().manual_seed(1))
    with msfno_b200.precision.library_scope():
        y.backward(gy.cuda())
    sdo = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
    tr = sfno_oracle.Transforms(cfg["img_size"], cfg["scale_factor"])
    fm = sfno_oracle.film_head(d["cond"], sdo).reshape(x.shape[0], 2, fl, 256)
    yo = sfno_oracle.sfno_forward(d["x"], sdo, tr, "non-linear", cfg["num_layers"], film_mod=fm, film_layers=fl, scale=d["scale"])
    yo.backward(gy)
    got = dict(net.named_parameters())
    for k in ("film_gen.film_head.net.4.weight", "film_gen.film_head.net.1.weight", "film_gen.film_head.net.4.bias"):
        assert rel_l2(got[k].grad, sdo[k].grad) < TOL_FP32, (k, rel_l2(got[k].grad, sdo[k].grad))
    assert got["decoder.fwd.0.weight"].grad is None


def test_weight_caches_survive_address_reuse():
    """Regression: caches of derived weights must be keyed on tensor identity, not on addresses -- a second model built
    after the first one is freed reuses the same device addresses (in a different order) with different values."""
    import gc
    d = torch.load(os.path.join(GOLD, "net_nonlinear_small.pt"))
    cfg = d["cfg"]
    tr = sfno_oracle.Transforms(cfg["img_size"], cfg["scale_factor"])
    x = d["x"]
    for seed in (11, 12, 13):
        sd = _oracle_sd(cfg, "non-linear", seed)
        net = _load(msfno_b200.FourierNeuralOperatorNet("cuda", None, **cfg), sd).cuda().eval()
        with torch.no_grad():
            got = net(x.cuda())
            want = sfno_oracle.sfno_forward(x, sd, tr, "non-linear", cfg["num_layers"])
        assert rel_l2(got, want) < TOL_FP32, seed
        del net
        gc.collect()
        torch.cuda.empty_cache()


def test_fold_affine_matches_algebra():
    """msfno_fold_affine: conv1x1(A*y + S, W) + bias == conv1x1(y, Wb) + bb (norm / FiLM affine folded into the conv)."""
    from msfno_b200.sfnonet import fold_affine
    g = torch.Generator().manual_seed(3)
    B, C, O = 3, 73, 40
    W = torch.zeros(O, 76)
    W[:, :C] = torch.randn(O, C, generator=g)
    A, S, bias = torch.rand(B, C, generator=g) + 0.5, torch.randn(B, C, generator=g), torch.randn(O, generator=g)
    Wb, bb = fold_affine(W.cuda(), A.cuda(), S.cuda(), bias.cuda())
    assert rel_l2(Wb[:, :, :C], W[None, :, :C].double() * A[:, None, :].double()) < 1e-6
    assert float(Wb[:, :, C:].abs().max()) == 0.0
    assert rel_l2(bb, S.double() @ W[:, :C].double().T + bias.double()) < 1e-6
    Wb2, bb2 = fold_affine(W.cuda(), A.cuda(), S.cuda(), None)
    assert rel_l2(bb2, S.double() @ W[:, :C].double().T) < 1e-6


@pytest.mark.parametrize("tier", ["fp32", "tf32"])
@pytest.mark.parametrize("film", [False, True])
def test_fold_norm_affine_is_the_two_launch_result(tier, film):
    """msfno_fold_norm_affine == msfno_norm_film_coeffs followed by msfno_fold_affine, bit for bit (the block's
    InstanceNorm -> FiLM -> fc1 hand-over, sfnonet.py:380-386), and both match the fp64 algebra."""
    from msfno_b200.sfnonet import fold_affine, fold_norm_affine, norm_film_coeffs, plane_stats
    from msfno_b200.conv import padded_weight
    msfno_b200.set_precision(tier)
    try:
        g = torch.Generator().manual_seed(21)
        B, C, O, H, W = 2, 37, 50, 12, 20
        y = (torch.randn(B, C, H, W, generator=g) * 3 + 1).cuda()
        norm = torch.nn.InstanceNorm2d(C, eps=1e-6, affine=True).cuda()
        with torch.no_grad():
            norm.weight.copy_(torch.randn(C, generator=g)); norm.bias.copy_(torch.randn(C, generator=g))
        fc = torch.nn.Conv2d(C, O, 1).cuda()
        gam = torch.randn(B, C, generator=g).cuda() if film else None
        bet = torch.randn(B, C, generator=g).cuda() if film else None
        st = plane_stats(y)
        Wp = padded_weight(fc.weight)
        with torch.no_grad():
            A, S = norm_film_coeffs(st, norm, B, C, H * W, gam, bet, 0.7)
            Wb0, bb0 = fold_affine(Wp, A, S, fc.bias)
            Wb1, bb1 = fold_norm_affine(Wp, st, norm, H * W, fc.bias, gam, bet, 0.7)
        assert torch.equal(Wb0, Wb1) and torch.equal(bb0, bb1)
        # fp64 algebra: conv(film(norm(y))) == conv_b(y) with the folded per-sample weights
        with torch.no_grad():
            z = norm(y).double()
            if film:
                z = (1 + gam.double()[:, :, None, None] * 0.7) * z + bet.double()[:, :, None, None] * 0.7
            want = torch.nn.functional.conv2d(z, fc.weight.double(), fc.bias.double())
            got = torch.einsum("boc,bchw->bohw", Wb1[:, :, :C].double(), y.double()) + bb1.double()[:, :, None, None]
        assert rel_l2(got, want) < (1e-5 if tier == "fp32" else 2e-3)
    finally:
        msfno_b200.set_precision("fp32")


def test_mean_carry_matches_algebra():
    """msfno_mean_carry: the per-block coefficients of the mean-carrying stream (tensor-core tier) against fp64 algebra --
    b2 = bias2 - mean(X), mu_out = mu + mean(X), sb = skip_bias + Wskip mu -- with and without an incoming offset."""
    from msfno_b200._lib import check, lib, ptr
    from msfno_b200.sfnonet import plane_stats
    g = torch.Generator().manual_seed(5)
    B, C, H, W = 3, 256, 9, 16
    X = (torch.randn(B, C, H, W, generator=g) + torch.randn(B, C, 1, 1, generator=g) * 3).cuda()
    st = plane_stats(X)
    mu = torch.randn(B, C, generator=g).cuda()
    bias2, sbias = torch.randn(C, generator=g).cuda(), torch.randn(C, generator=g).cuda()
    Wsk = torch.randn(C, C, generator=g).cuda()
    stream = torch.cuda.current_stream().cuda_stream
    mean = X.double().mean(dim=(2, 3))
    for use_mu in (True, False):
        b2 = torch.full((B, C), float("nan"), device="cuda")
        mo = torch.full((B, C), float("nan"), device="cuda")
        sb = torch.full((B, C), float("nan"), device="cuda") if use_mu else None
        check(lib.msfno_mean_carry(ptr(st), H * W, ptr(mu) if use_mu else None, ptr(bias2), ptr(Wsk) if use_mu else None, C,
                                   ptr(sbias) if use_mu else None, ptr(b2), ptr(mo), ptr(sb), B, C, stream), "mean_carry")
        assert rel_l2(b2, bias2.double()[None, :] - mean) < 1e-6
        assert rel_l2(mo, (mu.double() if use_mu else 0) + mean) < 1e-6
        if use_mu:
            assert rel_l2(sb, sbias.double()[None, :] + mu.double() @ Wsk.double().T) < 1e-6
    # skip-bias only (a block that receives an offset but cannot carry it on): stats may be NULL
    sb = torch.empty((B, C), device="cuda")
    check(lib.msfno_mean_carry(None, 0, ptr(mu), None, ptr(Wsk), C, ptr(sbias), None, None, ptr(sb), B, C, stream), "mean_carry")
    assert rel_l2(sb, sbias.double()[None, :] + mu.double() @ Wsk.double().T) < 1e-6
