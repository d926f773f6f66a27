"""ORACLE tooling (build container only): generate tests/golden/*.pt from the UNMODIFIED
reference modules (imported via oracle/ref_import.py) and assert that the restatement in
oracle/sfno_oracle.py reproduces them.

    python -m oracle.gen_golden            # writes tests/golden/, prints max deviations

Weights are NOT stored: they are regenerated from a seed by sfno_oracle.make_state_dict and
loaded into the reference modules with load_state_dict(strict=True) (which also pins
state_dict key/shape compatibility).  Each golden file carries a float64 checksum of the
weights so RNG drift is detected instead of silently comparing different models.
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_import, sfno_oracle, th_shim  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-300))


def sd_checksum(sd):
    return float(sum(v.double().abs().sum() for v in sd.values() if v.is_floating_point()))


def small_cfg(filter_type, embed, num_layers, mlp_ratio, spectral_layers, img=(24, 48), scale=2, in_chans=5):
    return dict(filter_type=filter_type, img_size=img, scale_factor=scale, in_chans=in_chans, out_chans=in_chans,
                embed_dim_sfno=embed, num_layers=num_layers, mlp_ratio=mlp_ratio, spectral_layers=spectral_layers)


def oracle_kwargs(cfg, seed, film_layers=0):
    return dict(filter_type=cfg["filter_type"], img_size=cfg["img_size"], scale_factor=cfg["scale_factor"],
                in_chans=cfg["in_chans"], out_chans=cfg["out_chans"], embed=cfg["embed_dim_sfno"],
                num_layers=cfg["num_layers"], mlp_ratio=cfg["mlp_ratio"], spectral_layers=cfg["spectral_layers"],
                seed=seed, film_layers=film_layers)


def load_into(net, sd):
    """Load oracle-generated weights into a reference module; non-float buffers (ii, jj,
    activation.bias) keep the reference's own values but must exist with the same keys."""
    ref_sd = net.state_dict()
    extra = {k: v for k, v in ref_sd.items() if k not in sd}
    for k in extra:
        assert k.endswith((".ii", ".jj", "activation.bias")), "unexpected reference key %s" % k
    full = dict(sd)
    full.update(extra)
    net.load_state_dict(full, strict=True)
    return {k: list(v.shape) for k, v in ref_sd.items()}


def main():
    os.makedirs(GOLD, exist_ok=True)
    ref = ref_import.load()
    report = {}
    torch.manual_seed(1234)

    # ---- 1. transforms (th_shim; restated library) ------------------------------------
    for name, (nlat, nlon, L, M, grid) in {
        "sht_equi_24x48": (24, 48, 12, 13, "equiangular"),
        "sht_lg_12x24": (12, 24, 12, 13, "legendre-gauss"),
        "sht_equi_37x72_l10": (37, 72, 10, 9, "equiangular"),
    }.items():
        g = torch.Generator().manual_seed(7)
        sht = th_shim.RealSHT(nlat, nlon, lmax=L, mmax=M, grid=grid).float()
        isht = th_shim.InverseRealSHT(nlat, nlon, lmax=L, mmax=M, grid=grid).float()
        x = torch.randn(2, 3, nlat, nlon, generator=g)
        c = sht(x)
        cin = torch.view_as_complex(torch.randn(2, 3, L, M, 2, generator=g))
        y = isht(cin)
        torch.save(dict(nlat=nlat, nlon=nlon, lmax=L, mmax=M, grid=grid, x=x, coeffs=c, cin=cin, y=y,
                        weights=sht.weights.clone(), pct=isht.pct.clone()), os.path.join(GOLD, name + ".pt"))

    # ---- 2. reference spectral modules (unmodified code) on small grids ----------------
    for ftype in ("linear", "non-linear"):
        g = torch.Generator().manual_seed(11)
        C = 8
        sht = th_shim.RealSHT(24, 48, lmax=12, mmax=13, grid="equiangular").float()
        isht = th_shim.InverseRealSHT(24, 48, lmax=12, mmax=13, grid="equiangular").float()
        sht.weights = sht.weights * 1e5
        isht.pct = isht.pct / 1e5
        x = torch.randn(2, C, 24, 48, generator=g)
        if ftype == "linear":
            mod = ref.layers.SpectralConvS2(sht, isht, C, use_complex_kernels=True)
            w = 0.02 * torch.randn(C, C, mod.w.shape[2], 2, generator=g)
            with torch.no_grad():
                mod.w.copy_(w)
            y_ref = mod(x).detach()
            y_or = sfno_oracle.spectral_conv_s2(x, w, sht, isht)
            payload = dict(x=x, w=w, y=y_ref)
        else:
            mod = ref.layers.SpectralAttentionS2(sht, isht, C, use_complex_kernels=True, hidden_size_factor=2,
                                                 complex_activation="real", spectral_layers=3, bias=False)
            ws = [0.2 * torch.randn(*p.shape, generator=g) for p in mod.w]
            wout = 0.2 * torch.randn(*mod.wout.shape, generator=g)
            with torch.no_grad():
                for p, w in zip(mod.w, ws):
                    p.copy_(w)
                mod.wout.copy_(wout)
            y_ref = mod(x).detach()
            y_or = sfno_oracle.spectral_attention_s2(x, ws, wout, sht, isht)
            payload = dict(x=x, ws=ws, wout=wout, y=y_ref)
        report["filter_" + ftype] = rel_l2(y_or, y_ref)
        assert report["filter_" + ftype] < 1e-6, report
        torch.save(payload, os.path.join(GOLD, "filter_%s_24x48.pt" % ftype.replace("-", "")))

    # ---- 3. reference FourierNeuralOperatorNet, small ---------------------------------
    keyshapes = {}
    for ftype in ("linear", "non-linear"):
        cfg = small_cfg(ftype, embed=16, num_layers=4, mlp_ratio=2.0, spectral_layers=3)
        net = ref.sfnonet.FourierNeuralOperatorNet("cpu", ref.Attributes(), **cfg).eval()
        sd = sfno_oracle.make_state_dict(**oracle_kwargs(cfg, seed=3))
        keyshapes["net_" + ftype] = load_into(net, sd)
        g = torch.Generator().manual_seed(5)
        x = torch.randn(2, cfg["in_chans"], *cfg["img_size"], generator=g)
        with torch.no_grad():
            y_ref = net(x)
        tr = sfno_oracle.Transforms(cfg["img_size"], cfg["scale_factor"])
        with torch.no_grad():
            y_or = sfno_oracle.sfno_forward(x, sd, tr, ftype, cfg["num_layers"])
        report["net_" + ftype] = rel_l2(y_or, y_ref)
        assert report["net_" + ftype] < 1e-5, report
        torch.save(dict(cfg=cfg, seed=3, sd_checksum=sd_checksum(sd), x=x, y=y_ref),
                   os.path.join(GOLD, "net_%s_small.pt" % ftype.replace("-", "")))

    # ---- 4. reference FourierNeuralOperatorNet_Filmed, small (embed must be 256) -------
    for film_layers in (1, 3):
        cfg = small_cfg("non-linear", embed=256, num_layers=3, mlp_ratio=0.25, spectral_layers=2,
                        img=(12, 24), scale=2, in_chans=4)
        B = 2
        attrs = ref.Attributes(film_gen_type="mae", cls="x", embed_dim=512, mlp_dim=1024, dropout=0.0,
                               film_layers=film_layers, scale_weight=1, repeat_film=False, batch_size=B)
        net = ref.sfnonet.FourierNeuralOperatorNet_Filmed(
            "cpu", attrs, mlp_ratio=cfg["mlp_ratio"], advanced_logging=True, film_layers=film_layers,
            model_depth=6, **{k: v for k, v in cfg.items() if k != "mlp_ratio"}).eval()
        # NOTE: the Filmed ctor passes mlp_ratio only to its rebuilt blocks; super().__init__ sees the
        # default 2.0 -- irrelevant here because super's blocks are discarded (sfnonet.py:710-718).
        sd = sfno_oracle.make_state_dict(**oracle_kwargs(cfg, seed=9, film_layers=film_layers))
        keyshapes["filmed_%d" % film_layers] = load_into(net, sd)
        g = torch.Generator().manual_seed(6)
        x = torch.randn(B, cfg["in_chans"], *cfg["img_size"], generator=g)
        cond = torch.randn(B, 512, generator=g)
        scale = 0.7
        with torch.no_grad():
            y_ref = net(x, cond, scale)
        tr = sfno_oracle.Transforms(cfg["img_size"], cfg["scale_factor"])
        with torch.no_grad():
            fm = sfno_oracle.film_head(cond, sd).reshape(B, 2, film_layers, 256)
            y_or = sfno_oracle.sfno_forward(x, sd, tr, "non-linear", cfg["num_layers"], film_mod=fm,
                                            film_layers=film_layers, scale=scale)
        report["filmed_%d" % film_layers] = rel_l2(y_or, y_ref)
        assert report["filmed_%d" % film_layers] < 1e-5, report
        torch.save(dict(cfg=cfg, seed=9, film_layers=film_layers, scale=scale, sd_checksum=sd_checksum(sd),
                        x=x, cond=cond, y=y_ref, gamma=net.gamma.detach(), beta=net.beta.detach()),
                   os.path.join(GOLD, "filmed_fl%d_small.pt" % film_layers))

    # ---- 5. FiLM module + backward of the reference filter (autograd of unmodified code) --
    g = torch.Generator().manual_seed(21)
    x = torch.randn(2, 256, 6, 12, generator=g)
    gam, bet = torch.randn(2, 256, generator=g), torch.randn(2, 256, generator=g)
    y = ref.sfnonet.FiLM()(x, gam, bet, 0.3)
    assert rel_l2(sfno_oracle.film(x, gam, bet, 0.3), y) < 1e-7
    torch.save(dict(x=x, gamma=gam, beta=bet, scale=0.3, y=y), os.path.join(GOLD, "film_small.pt"))

    for ftype in ("linear", "non-linear"):
        d = torch.load(os.path.join(GOLD, "filter_%s_24x48.pt" % ftype.replace("-", "")))
        C = 8
        sht = th_shim.RealSHT(24, 48, lmax=12, mmax=13, grid="equiangular").float()
        isht = th_shim.InverseRealSHT(24, 48, lmax=12, mmax=13, grid="equiangular").float()
        sht.weights = sht.weights * 1e5
        isht.pct = isht.pct / 1e5
        x = d["x"].clone().requires_grad_(True)
        gy = torch.randn(d["y"].shape, generator=g)
        if ftype == "linear":
            mod = ref.layers.SpectralConvS2(sht, isht, C, use_complex_kernels=True)
            with torch.no_grad():
                mod.w.copy_(d["w"])
            mod(x).backward(gy)
            grads = dict(gy=gy, gx=x.grad.clone(), gw=mod.w.grad.clone())
        else:
            mod = ref.layers.SpectralAttentionS2(sht, isht, C, use_complex_kernels=True, hidden_size_factor=2,
                                                 complex_activation="real", spectral_layers=3, bias=False)
            with torch.no_grad():
                for p, w in zip(mod.w, d["ws"]):
                    p.copy_(w)
                mod.wout.copy_(d["wout"])
            mod(x).backward(gy)
            grads = dict(gy=gy, gx=x.grad.clone(), gws=[p.grad.clone() for p in mod.w], gwout=mod.wout.grad.clone())
        d.update(grads)
        torch.save(d, os.path.join(GOLD, "filter_%s_24x48.pt" % ftype.replace("-", "")))

    with open(os.path.join(GOLD, "state_dict_keys.json"), "w") as f:
        json.dump(keyshapes, f, indent=0, sort_keys=True)
    with open(os.path.join(GOLD, "oracle_vs_reference.json"), "w") as f:
        json.dump(report, f, indent=1, sort_keys=True)
    print(json.dumps(report, indent=1))
    sz = sum(os.path.getsize(os.path.join(GOLD, f)) for f in os.listdir(GOLD))
    print("golden bytes:", sz)


if __name__ == "__main__":
    main()
