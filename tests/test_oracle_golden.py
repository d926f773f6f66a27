"""The oracle restatement reproduces the golden vectors written by oracle/gen_golden.py from the
UNMODIFIED reference modules (bit-for-bit on the generating machine; 1e-6 allowed across BLAS builds)."""
import json
import os

import torch

from conftest import GOLD, rel_l2
from oracle import sfno_oracle, th_shim


def _sd_checksum(sd):
    return float(sum(v.double().abs().sum() for v in sd.values() if v.is_floating_point()))


def _transforms_24x48():
    sht = th_shim.RealSHT(24, 48, lmax=12, mmax=13, grid="equiangular").float()
    isht = th_shim.InverseRealSHT(24, 48, lmax=12, mmax=13, grid="equiangular").float()
    sht.weights = sht.weights * 1e5
    isht.pct = isht.pct / 1e5
    return sht, isht


def test_reference_reported_zero_deviation():
    rep = json.load(open(os.path.join(GOLD, "oracle_vs_reference.json")))
    assert all(v < 1e-6 for v in rep.values()), rep


def test_filters():
    sht, isht = _transforms_24x48()
    d = torch.load(os.path.join(GOLD, "filter_linear_24x48.pt"))
    assert rel_l2(sfno_oracle.spectral_conv_s2(d["x"], d["w"], sht, isht), d["y"]) < 1e-6
    d = torch.load(os.path.join(GOLD, "filter_nonlinear_24x48.pt"))
    assert rel_l2(sfno_oracle.spectral_attention_s2(d["x"], d["ws"], d["wout"], sht, isht), d["y"]) < 1e-6


def test_filter_backward_matches_reference_autograd():
    sht, isht = _transforms_24x48()
    d = torch.load(os.path.join(GOLD, "filter_linear_24x48.pt"))
    x = d["x"].clone().requires_grad_(True)
    w = d["w"].clone().requires_grad_(True)
    sfno_oracle.spectral_conv_s2(x, w, sht, isht).backward(d["gy"])
    assert rel_l2(x.grad, d["gx"]) < 1e-5 and rel_l2(w.grad, d["gw"]) < 1e-5
    d = torch.load(os.path.join(GOLD, "filter_nonlinear_24x48.pt"))
    x = d["x"].clone().requires_grad_(True)
    ws = [t.clone().requires_grad_(True) for t in d["ws"]]
    wout = d["wout"].clone().requires_grad_(True)
    sfno_oracle.spectral_attention_s2(x, ws, wout, sht, isht).backward(d["gy"])
    assert rel_l2(x.grad, d["gx"]) < 1e-5 and rel_l2(wout.grad, d["gwout"]) < 1e-5
    for a, b in zip(ws, d["gws"]):
        assert rel_l2(a.grad, b) < 1e-5


def test_small_nets():
    for name, ftype in (("net_linear_small.pt", "linear"), ("net_nonlinear_small.pt", "non-linear")):
        d = torch.load(os.path.join(GOLD, name))
        cfg = d["cfg"]
        sd = sfno_oracle.make_state_dict(filter_type=ftype, img_size=cfg["img_size"], scale_factor=cfg["scale_factor"],
                                         in_chans=cfg["in_chans"], out_chans=cfg["out_chans"], embed=cfg["embed_dim_sfno"],
                                         num_layers=cfg["num_layers"], mlp_ratio=cfg["mlp_ratio"],
                                         spectral_layers=cfg["spectral_layers"], seed=d["seed"])
        assert abs(_sd_checksum(sd) - d["sd_checksum"]) < 1e-6 * d["sd_checksum"], "RNG drift: weights differ"
        tr = sfno_oracle.Transforms(cfg["img_size"], cfg["scale_factor"])
        with torch.no_grad():
            y = sfno_oracle.sfno_forward(d["x"], sd, tr, ftype, cfg["num_layers"])
        assert rel_l2(y, d["y"]) < 1e-6


def test_filmed_nets():
    for fl in (1, 3):
        d = torch.load(os.path.join(GOLD, "filmed_fl%d_small.pt" % fl))
        cfg = d["cfg"]
        sd = sfno_oracle.make_state_dict(filter_type="non-linear", img_size=cfg["img_size"],
                                         scale_factor=cfg["scale_factor"], in_chans=cfg["in_chans"],
                                         out_chans=cfg["out_chans"], embed=cfg["embed_dim_sfno"],
                                         num_layers=cfg["num_layers"], mlp_ratio=cfg["mlp_ratio"],
                                         spectral_layers=cfg["spectral_layers"], seed=d["seed"], film_layers=fl)
        assert abs(_sd_checksum(sd) - d["sd_checksum"]) < 1e-6 * d["sd_checksum"]
        tr = sfno_oracle.Transforms(cfg["img_size"], cfg["scale_factor"])
        B = d["x"].shape[0]
        with torch.no_grad():
            fm = sfno_oracle.film_head(d["cond"], sd).reshape(B, 2, fl, 256)
            y = sfno_oracle.sfno_forward(d["x"], sd, tr, "non-linear", cfg["num_layers"], film_mod=fm, film_layers=fl,
                                         scale=d["scale"])
        assert rel_l2(fm[:, 0], d["gamma"]) < 1e-6
        assert rel_l2(y, d["y"]) < 1e-6


def test_film_and_transform_goldens():
    d = torch.load(os.path.join(GOLD, "film_small.pt"))
    assert rel_l2(sfno_oracle.film(d["x"], d["gamma"], d["beta"], d["scale"]), d["y"]) < 1e-7
    for name in ("sht_equi_24x48", "sht_lg_12x24", "sht_equi_37x72_l10"):
        d = torch.load(os.path.join(GOLD, name + ".pt"))
        sht = th_shim.RealSHT(d["nlat"], d["nlon"], lmax=d["lmax"], mmax=d["mmax"], grid=d["grid"]).float()
        isht = th_shim.InverseRealSHT(d["nlat"], d["nlon"], lmax=d["lmax"], mmax=d["mmax"], grid=d["grid"]).float()
        assert rel_l2(torch.view_as_real(sht(d["x"])), torch.view_as_real(d["coeffs"])) < 1e-6
        assert rel_l2(isht(d["cin"]), d["y"]) < 1e-6


def test_loss_oracle_matches_reference_losses():
    """oracle.sfno_oracle.l2_sphere / cosine_mse vs the UNMODIFIED reference losses.py (mounted tree or its staged copy)."""
    import pytest
    import torch
    from oracle import ref_import, sfno_oracle
    if not ref_import.available():
        pytest.skip("no copy of the reference")
    ref = ref_import.load()
    ref_import.use_harmonics(ref.th_shim)
    g = torch.Generator().manual_seed(1)
    a, b = torch.randn(2, 4, 24, 48, generator=g), torch.randn(2, 4, 24, 48, generator=g)
    for rel, sq in ((True, False), (False, True), (True, True)):
        assert abs(float(sfno_oracle.l2_sphere(a, b, rel, sq, True)) - float(ref.losses.L2Sphere(rel, sq)(a, b))) < 1e-6
        assert abs(float(sfno_oracle.l2_sphere(a, b, rel, sq, False)) - float(ref.losses.L2Sphere_noSine(rel, sq)(a, b))) < 1e-6
    for red in ("mean", "sum"):
        assert abs(float(sfno_oracle.cosine_mse(a, b, red)) - float(ref.losses.CosineMSELoss(red)(a, b))) < 1e-6
