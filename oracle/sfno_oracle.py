"""ORACLE (test infrastructure, not product code) -- CPU restatement of the reference's
spectral hot path, written functionally over a plain `state_dict`.

Every function cites the reference lines it follows (paths relative to /root/reference).
It uses only torch CPU library ops (fft / einsum / conv) -- the same ops the reference
itself is made of -- plus `oracle/th_shim.py` for the un-vendored torch_harmonics.

PINNING: `oracle/gen_golden.py` (run in the build container, where /root/reference is
mounted) imports the UNMODIFIED reference modules through `oracle/ref_import.py`, runs
them on seeded inputs and (a) asserts this restatement reproduces them, (b) writes the
reference's outputs to tests/golden/*.pt.  tests/test_oracle_golden.py re-checks this
file against those committed vectors everywhere (no /root/reference needed).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.
"""
import math

import torch
import torch.nn.functional as F

from . import th_shim


# --------------------------------------------------------------------------- transforms
class Transforms:
    """The four transform objects the net builds once and shares across blocks
    (MSFNO/Models/sfno/sfnonet.py:537-555), including the 1e5 rescale."""

    def __init__(self, img_size=(721, 1440), scale_factor=6, hard_thresholding_fraction=1.0,
                 rescale=1e5):
        H, W = img_size
        self.h, self.w = H // scale_factor, W // scale_factor
        modes_lat = int(self.h * hard_thresholding_fraction)
        modes_lon = int((self.w // 2 + 1) * hard_thresholding_fraction)
        self.lmax, self.mmax = modes_lat, modes_lon
        self.trans_down = th_shim.RealSHT(H, W, lmax=modes_lat, mmax=modes_lon, grid="equiangular").float()
        self.itrans_up = th_shim.InverseRealSHT(H, W, lmax=modes_lat, mmax=modes_lon, grid="equiangular").float()
        self.trans = th_shim.RealSHT(self.h, self.w, lmax=modes_lat, mmax=modes_lon, grid="legendre-gauss").float()
        self.itrans = th_shim.InverseRealSHT(self.h, self.w, lmax=modes_lat, mmax=modes_lon, grid="legendre-gauss").float()
        self.trans_down.weights = self.trans_down.weights * rescale
        self.itrans_up.pct = self.itrans_up.pct / rescale
        self.trans.weights = self.trans.weights * rescale
        self.itrans.pct = self.itrans.pct / rescale


# --------------------------------------------------------------------------- spectral ops
def complex_relu_real(z):
    """ComplexReLU mode "real": ReLU on Re, Im untouched (activations.py:42-46)."""
    return torch.complex(torch.relu(z.real), z.imag)


def spectral_conv_s2(x, w, sht, isht, sparsity_threshold=0.0):
    """SpectralConvS2.forward (layers.py:398-427) with compl_contract_fwd_c
    (contractions.py:37-41).  w: [C_out, C_in, n, 2], n = tril_indices(lmax, mmax)."""
    dtype = x.dtype
    c = sht(x.float())  # [B,C,L,M] complex64
    L, M = c.shape[-2], c.shape[-1]
    ii, jj = torch.tril_indices(L, M)
    a = c[:, :, ii, jj]  # [B,C,n]
    out = torch.einsum("bin,kin->bkn", a, torch.view_as_complex(w.float().contiguous()))
    modes = torch.zeros_like(c)
    modes[:, :, ii, jj] = out
    if sparsity_threshold != 0.0:
        modes = torch.view_as_complex(F.softshrink(torch.view_as_real(modes), lambd=sparsity_threshold))
    return isht(modes).to(dtype)


def spectral_attention_mlp(c, ws, wout):
    """SpectralAttentionS2.forward_mlp (layers.py:604-620), bias=False, dropout identity,
    compl_mul2d_fwd_c (contractions.py:132-137).  c: [B,C,L,M] complex."""
    for w in ws:
        c = torch.einsum("bixy,io->boxy", c, torch.view_as_complex(w.float().contiguous()))
        c = complex_relu_real(c)
    return torch.einsum("bixy,io->boxy", c, torch.view_as_complex(wout.float().contiguous()))


def spectral_attention_s2(x, ws, wout, sht, isht):
    """SpectralAttentionS2.forward (layers.py:622-641)."""
    dtype = x.dtype
    c = sht(x.float())
    c = spectral_attention_mlp(c, ws, wout)
    return isht(c).to(dtype)


def film(x, gammas, betas, scale=1.0):
    """FiLM.forward (sfnonet.py:689-697): (1 + gamma*scale) * x + beta*scale."""
    return (1 + gammas[:, :, None, None] * scale) * x + betas[:, :, None, None] * scale


def instance_norm(x, weight, bias, eps=1e-6):
    """nn.InstanceNorm2d(eps=1e-6, affine=True, track_running_stats=False) (sfnonet.py:491-499)."""
    return F.instance_norm(x, weight=weight, bias=bias, eps=eps)


def mlp_1x1(x, sd, prefix):
    """MLP.fwd = Conv2d(1x1) -> GELU(erf) -> Conv2d(1x1) (layers.py:145-178)."""
    w0, b0 = sd[prefix + "fwd.0.weight"], sd.get(prefix + "fwd.0.bias")
    w2, b2 = sd[prefix + "fwd.2.weight"], sd.get(prefix + "fwd.2.bias")
    x = F.conv2d(x, w0, b0)
    x = F.gelu(x)
    return F.conv2d(x, w2, b2)


def filter_forward(x, sd, prefix, filter_type, sht, isht):
    """SpectralFilterLayer.forward dispatch (sfnonet.py:56-133)."""
    if filter_type == "linear":
        return spectral_conv_s2(x, sd[prefix + "filter.w"], sht, isht)
    ws = []
    i = 0
    while prefix + f"filter.w.{i}" in sd:
        ws.append(sd[prefix + f"filter.w.{i}"])
        i += 1
    return spectral_attention_s2(x, ws, sd[prefix + "filter.wout"], sht, isht)


def block_forward(x, sd, i, num_layers, filter_type, tr, gamma=None, beta=None, scale=1.0):
    """FourierNeuralOperatorBlock[_Filmed].forward (sfnonet.py:221-251, 359-393) for block i
    of a net laid out as in sfnonet.py:557-610 (first block trans_down, last block itrans_up,
    inner/outer skips and MLP only as wired there)."""
    p = f"blocks.{i}."
    first, last = i == 0, i == num_layers - 1
    sht = tr.trans_down if first else tr.trans
    isht = tr.itrans_up if last else tr.itrans
    residual = x
    x = instance_norm(x, sd[p + "norm0.weight"], sd[p + "norm0.bias"])
    x = filter_forward(x, sd, p + "filter_layer.", filter_type, sht, isht).contiguous()
    if p + "inner_skip.weight" in sd:
        x = x + F.conv2d(residual, sd[p + "inner_skip.weight"], sd[p + "inner_skip.bias"])
    if filter_type == "linear":
        x = F.gelu(x)
    x = instance_norm(x, sd[p + "norm1.weight"], sd[p + "norm1.bias"])
    if gamma is not None:
        x = film(x, gamma, beta, scale)
    if p + "mlp.fwd.0.weight" in sd:
        x = mlp_1x1(x, sd, p + "mlp.")
    if 0 < i < num_layers - 1:  # outer_skip == "identity"
        x = x + residual
    return x


def sfno_forward(x, sd, tr, filter_type="non-linear", num_layers=12, film_mod=None, film_layers=0,
                 scale=1.0, big_skip=True):
    """FourierNeuralOperatorNet.forward (sfnonet.py:662-686) and, when film_mod
    ([B,2,film_layers,C]) is given, FourierNeuralOperatorNet_Filmed.forward (sfnonet.py:787-860,
    non-checkpointed branch, repeat_film False)."""
    residual = x
    x = mlp_1x1(x, sd, "encoder.")
    x = x + sd["pos_embed"]
    for i in range(num_layers):
        if film_mod is not None and i >= num_layers - film_layers:
            fi = i - (num_layers - film_layers)
            x = block_forward(x, sd, i, num_layers, filter_type, tr, film_mod[:, 0, fi], film_mod[:, 1, fi], scale)
        else:
            x = block_forward(x, sd, i, num_layers, filter_type, tr)
    if big_skip:
        x = torch.cat((x, residual), dim=1)
    return mlp_1x1(x, sd, "decoder.")


def film_head(cond, sd, prefix="film_gen.film_head.net."):
    """FeedForward film head (sfnonet.py:915-928): LayerNorm -> Linear -> GELU -> Linear,
    reshaped to [B, 2, film_layers, 256] by Film_wrapper.forward (sfnonet.py:900-912)."""
    x = F.layer_norm(cond, (cond.shape[-1],), sd[prefix + "0.weight"], sd[prefix + "0.bias"])
    x = F.linear(x, sd[prefix + "1.weight"], sd[prefix + "1.bias"])
    x = F.gelu(x)
    return F.linear(x, sd[prefix + "4.weight"], sd[prefix + "4.bias"])


# --------------------------------------------------------------------------- random init
def l2_sphere(prd, tar, relative=True, squared=False, sine=True):
    """L2Sphere (sine=True) / L2Sphere_noSine, reduction "sum" or "mean" (/root/reference MSFNO/Models/losses.py:80-155)."""
    H = prd.shape[2]
    w = torch.tensor(th_shim.legendre_gauss_weights(H, -1, 1)[1], dtype=prd.dtype)
    if sine:
        w = torch.abs(w * torch.cos(torch.linspace(-math.pi / 2, math.pi / 2, H, dtype=prd.dtype)))
    w = w[None, None, :, None]
    loss = (w * (prd - tar) ** 2).sum(dim=(-1, -2))
    if relative:
        loss = loss / (w * tar ** 2).sum(dim=(-1, -2))
    if not squared:
        loss = torch.sqrt(loss)
    return loss.sum()


def cosine_mse(x, y, reduction="mean", eps=1e-4):
    """CosineMSELoss (/root/reference MSFNO/Models/losses.py:6-37)."""
    H, W = x.shape[2], x.shape[3]
    w = torch.clamp(torch.cos(torch.linspace(-math.pi / 2, math.pi / 2, H, dtype=x.dtype)), min=0.0) + eps
    w = (w / w.sum(dim=-1, keepdim=True))[None, None, :, None]
    loss = (x - y) ** 2 * w
    return loss.mean() if reduction == "mean" else loss.sum() / W


def trunc_normal_(t, std=0.02, gen=None):
    """trunc_normal_(std=0.02) clipped to [-2,2] absolute (layers.py:29-84)."""
    with torch.no_grad():
        t.normal_(0.0, std, generator=gen)
        bad = (t < -2.0) | (t > 2.0)
        while bad.any():
            t[bad] = torch.empty(int(bad.sum())).normal_(0.0, std, generator=gen)
            bad = (t < -2.0) | (t > 2.0)
    return t


def make_state_dict(filter_type="non-linear", img_size=(721, 1440), scale_factor=6, in_chans=73,
                    out_chans=73, embed=256, num_layers=12, mlp_ratio=2.0, spectral_layers=3,
                    seed=0, film_layers=0, film_embed=512, film_mlp=1024, pos_embed=True):
    """Random-init weights with the reference's shapes, keys and init scales (SURVEY.md 8(b);
    sfnonet.py:635-646, layers.py:376-386,575-590).  Seeded by a private generator: the VALUES
    are ours (synthetic), the LAYOUT is the reference's."""
    g = torch.Generator().manual_seed(seed)
    H, W = img_size
    h, w = H // scale_factor, W // scale_factor
    L, M = h, w // 2 + 1
    n = int(torch.tril_indices(L, M).shape[1])
    hid = int(embed * mlp_ratio)
    sd = {}

    def conv(name, cout, cin, bias=True):
        sd[name + ".weight"] = trunc_normal_(torch.empty(cout, cin, 1, 1), gen=g)
        if bias:
            sd[name + ".bias"] = torch.zeros(cout)

    if pos_embed:
        sd["pos_embed"] = trunc_normal_(torch.empty(1, embed, H, W), gen=g)
    conv("encoder.fwd.0", embed, in_chans)
    conv("encoder.fwd.2", embed, embed, bias=False)
    for i in range(num_layers):
        p = f"blocks.{i}."
        for nm in ("norm0", "norm1"):
            sd[p + nm + ".weight"] = torch.ones(embed)
            sd[p + nm + ".bias"] = torch.zeros(embed)
        f = p + "filter_layer.filter."
        if filter_type == "linear":
            sd[f + "w"] = 0.02 * torch.randn(embed, embed, n, 2, generator=g)
        else:
            sd[f + "w.0"] = 0.02 * torch.randn(embed, hid, 2, generator=g)
            for l in range(1, spectral_layers):
                sd[f + f"w.{l}"] = 0.02 * torch.randn(hid, hid, 2, generator=g)
            sd[f + "wout"] = 0.02 * torch.randn(hid, embed, 2, generator=g)
        if 0 < i < num_layers - 1:
            conv(p + "inner_skip", embed, embed)
        if i < num_layers - 1:
            conv(p + "mlp.fwd.0", hid, embed)
            conv(p + "mlp.fwd.2", embed, hid)
    conv("decoder.fwd.0", embed, embed + in_chans)
    conv("decoder.fwd.2", out_chans, embed, bias=False)
    if film_layers:
        q = "film_gen.film_head.net."
        sd[q + "0.weight"] = torch.ones(film_embed)
        sd[q + "0.bias"] = torch.zeros(film_embed)
        s1, s4 = 1.0 / math.sqrt(film_embed), 1.0 / math.sqrt(film_mlp)
        sd[q + "1.weight"] = (torch.rand(film_mlp, film_embed, generator=g) * 2 - 1) * s1
        sd[q + "1.bias"] = (torch.rand(film_mlp, generator=g) * 2 - 1) * s1
        sd[q + "4.weight"] = (torch.rand(2 * film_layers * 256, film_mlp, generator=g) * 2 - 1) * s4
        sd[q + "4.bias"] = (torch.rand(2 * film_layers * 256, generator=g) * 2 - 1) * s4
    return sd
