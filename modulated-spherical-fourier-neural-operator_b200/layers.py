"""Spectral layers of the SFNO block: drop-in replacements for the classes in
/root/reference MSFNO/Models/sfno/layers.py (SpectralConvS2 :336-427, SpectralAttentionS2 :536-641,
MLP :145-178, DropPath :88-118, trunc_normal_ :29-84) and activations.py (ComplexReLU :9-51).

Constructor signatures, parameter names/shapes and state_dict keys are the reference's; the forward
passes call the sm_100a kernels behind include/msfno_b200.h and keep the spectral coefficients in the
packed layouts end to end (no tril gather/scatter, no zeros()+slice-assign, no view_as_real shuffles).
"""
import ctypes
import math

import torch
import torch.nn as nn

from . import _lib
from . import precision as _precision
from ._lib import check, lib, ptr
from .sht import RealSHT, InverseRealSHT, _stream, _require_cuda, relayout


# ------------------------------------------------------------------------------- init helpers
def trunc_normal_(tensor, mean=0.0, std=1.0, a=-2.0, b=2.0):
    """Truncated normal init (values outside [a, b] redrawn), as layers.py:29-84."""
    return nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)


class DropPath(nn.Module):
    """Stochastic depth per sample (layers.py:88-118)."""

    def __init__(self, drop_prob=None):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        if not self.drop_prob or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.dim() - 1)).bernoulli_(keep)
        return x.div(keep) * mask


class MLP(nn.Module):
    """1x1-conv channel MLP (layers.py:145-178): Conv2d -> act -> [drop] -> Conv2d -> [drop]."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, output_bias=True,
                 drop_rate=0.0, checkpointing_mlp=False):
        super().__init__()
        self.checkpointing_mlp = checkpointing_mlp
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        fc1 = nn.Conv2d(in_features, hidden_features, 1, bias=True)
        act = act_layer()
        fc2 = nn.Conv2d(hidden_features, out_features, 1, bias=output_bias)
        if drop_rate > 0.0:
            drop = nn.Dropout(drop_rate)
            self.fwd = nn.Sequential(fc1, act, drop, fc2, drop)
        else:
            self.fwd = nn.Sequential(fc1, act, fc2)

    def forward(self, x):
        if self.checkpointing_mlp:
            from torch.utils.checkpoint import checkpoint
            return checkpoint(self.fwd, x, use_reentrant=False)
        return self.fwd(x)


class ComplexReLU(nn.Module):
    """ComplexReLU (activations.py:9-51).  Only mode "real" (ReLU on Re, Im untouched) is on the hot
    path, where it is fused into the spectral-MLP GEMM epilogue; the module exists for state_dict
    compatibility (buffer `bias`) and for standalone use on complex tensors."""

    def __init__(self, negative_slope=0.0, mode="cartesian", bias_shape=None):
        super().__init__()
        self.mode = mode
        if mode in ("modulus", "halfplane"):
            raise NotImplementedError("ComplexReLU mode %r is not on the MSFNO hot path" % mode)
        self.register_buffer("bias", torch.zeros((1), dtype=torch.float32))
        self.negative_slope = negative_slope

    def forward(self, z):
        if self.mode == "real":
            return torch.complex(nn.functional.leaky_relu(z.real, self.negative_slope), z.imag)
        if self.mode == "cartesian":
            return torch.view_as_complex(nn.functional.leaky_relu(torch.view_as_real(z), self.negative_slope))
        return z


# ------------------------------------------------------------------------------- autograd functions
class _SpecConv(torch.autograd.Function):
    """a_pm [B,P,2Ci], w [Co,Ci,n,2] -> out_pm [B,P,2Co]  (msfno_specconv_fwd / bwd_x / bwd_w)."""

    @staticmethod
    def forward(ctx, a_pm, w, sht):
        plan = sht._get_plan(a_pm.device)
        B, Ci, Co = a_pm.shape[0], w.shape[1], w.shape[0]
        if w.shape[2] != plan.ntril:
            raise RuntimeError("SpectralConvS2: weight has %d modes, transform has %d" % (w.shape[2], plan.ntril))
        out = torch.empty((B, plan.P, 2 * Co), dtype=torch.float32, device=a_pm.device)
        ws = torch.empty(lib.msfno_specconv_ws_floats(plan.h, B, Ci, Co), dtype=torch.float32, device=a_pm.device)
        check(lib.msfno_specconv_fwd(plan.h, ptr(a_pm), ptr(w), ptr(out), ptr(ws), B, Ci, Co, _stream()), "specconv_fwd")
        ctx.sht = sht
        ctx.save_for_backward(a_pm, w)
        return out

    @staticmethod
    def backward(ctx, g):
        a_pm, w = ctx.saved_tensors
        plan = ctx.sht._get_plan(g.device)
        B, Ci, Co = a_pm.shape[0], w.shape[1], w.shape[0]
        g = g.contiguous()
        ga = gw = None
        ws = torch.empty(lib.msfno_specconv_ws_floats(plan.h, B, Ci, Co), dtype=torch.float32, device=g.device)
        if ctx.needs_input_grad[0]:
            ga = torch.empty_like(a_pm)
            check(lib.msfno_specconv_bwd_x(plan.h, ptr(g), ptr(w), ptr(ga), ptr(ws), B, Ci, Co, _stream()), "specconv_bwd_x")
        if ctx.needs_input_grad[1]:
            gw = torch.empty_like(w)
            check(lib.msfno_specconv_bwd_w(plan.h, ptr(a_pm), ptr(g), ptr(gw), ptr(ws), B, Ci, Co, _stream()), "specconv_bwd_w")
        return ga, gw, None


class _SpecAttn(torch.autograd.Function):
    """a_pm [B,P,2C] -> out_cm [B,2C,P] through the complex MLP (msfno_specattn_fwd / bwd)."""

    @staticmethod
    def forward(ctx, a_pm, sht, precision, wout, *ws_layers):
        plan = sht._get_plan(a_pm.device)
        B, C = a_pm.shape[0], a_pm.shape[2] // 2
        hid, nl = wout.shape[0], len(ws_layers)
        out = torch.empty((B, 2 * C, plan.P), dtype=torch.float32, device=a_pm.device)
        wsf = lib.msfno_specattn_ws_floats(plan.h, B, C, hid, nl)
        ws = torch.empty(wsf, dtype=torch.float32, device=a_pm.device)
        warr = (ctypes.c_void_p * nl)(*[w.data_ptr() for w in ws_layers])
        check(lib.msfno_specattn_fwd(plan.h, ptr(a_pm), warr, nl, ptr(wout), ptr(out), ptr(ws), B, C, hid, precision,
                                     _stream()), "specattn_fwd")
        ctx.sht, ctx.dims = sht, (B, C, hid, nl)
        ctx.save_for_backward(a_pm, ws, wout, *ws_layers)
        return out

    @staticmethod
    def backward(ctx, g):
        a_pm, ws, wout, *ws_layers = ctx.saved_tensors
        B, C, hid, nl = ctx.dims
        plan = ctx.sht._get_plan(g.device)
        g = g.contiguous()
        ga = torch.empty_like(a_pm)
        # weight gradients only where autograd wants them (inputs: a_pm, sht, precision, wout, *ws_layers): with a
        # frozen backbone the weight-gradient GEMMs, the largest of this backward, are skipped altogether
        need = ctx.needs_input_grad
        gwout = torch.empty_like(wout) if need[3] else None
        gws = [torch.empty_like(w) if need[4 + i] else None for i, w in enumerate(ws_layers)]
        scratch = torch.empty(lib.msfno_specattn_bwd_scratch_floats(plan.h, B, C, hid, nl), dtype=torch.float32,
                              device=g.device)
        garr = (ctypes.c_void_p * nl)(*[t.data_ptr() if t is not None else None for t in gws])
        check(lib.msfno_specattn_bwd(plan.h, ptr(a_pm), ptr(g), ptr(ws), ptr(ga), garr, ptr(gwout), ptr(scratch), nl, B, C,
                                     hid, _stream()), "specattn_bwd")
        return (ga, None, None, gwout, *gws)


# ------------------------------------------------------------------------------- spectral modules
class SpectralConvS2(nn.Module):
    """Spectral convolution on the sphere (layers.py:336-427): SHT -> per-mode complex channel
    contraction with w [C,C,n,2] over the n = |tril(lmax,mmax)| modes -> inverse SHT."""

    def __init__(self, forward_transform, inverse_transform, hidden_size, sparsity_threshold=0.0,
                 use_complex_kernels=False, compression=None, rank=128, bias=False):
        super().__init__()
        if compression is not None:
            raise NotImplementedError("tensor-train compression is not used by the reference nets (sfnonet.py:429)")
        if bias:
            raise NotImplementedError("SpectralConvS2 bias is dead code in the reference (layers.py:393-396)")
        self.hidden_size = hidden_size
        self.sparsity_threshold = sparsity_threshold
        self.scale = 0.02
        self.forward_transform = forward_transform
        self.inverse_transform = inverse_transform
        self.modes_lat = forward_transform.lmax
        self.modes_lon = forward_transform.mmax
        assert inverse_transform.lmax == self.modes_lat
        assert inverse_transform.mmax == self.modes_lon
        ii, jj = torch.tril_indices(self.modes_lat, self.modes_lon)
        self.register_buffer("ii", ii)
        self.register_buffer("jj", jj)
        self.w = nn.Parameter(self.scale * torch.randn(hidden_size, hidden_size, len(ii), 2))

    def spectral(self, a_pm):
        """PM coefficients -> CM coefficients (the contraction + optional soft-shrink)."""
        w = self.w if self.w.dtype == torch.float32 else self.w.float()
        out = _SpecConv.apply(a_pm, w.contiguous(), self.forward_transform)
        out = relayout(out, self.forward_transform, _lib.LAYOUT_PM, _lib.LAYOUT_CM, a_pm.shape[0], self.w.shape[0])
        if self.sparsity_threshold != 0.0:
            out = nn.functional.softshrink(out, lambd=self.sparsity_threshold)
        return out

    @_lib.on_input_device
    def forward(self, x, in_scale=None, in_shift=None, **epilogue):
        _require_cuda(x, "SpectralConvS2")
        dtype = x.dtype
        a = self.forward_transform.forward_packed(x, in_scale, in_shift)
        c = self.spectral(a)
        y = self.inverse_transform.inverse_packed(c, **epilogue)
        return y.to(dtype)


class SpectralAttentionS2(nn.Module):
    """Spectral 'attention' on the sphere (layers.py:536-641): SHT -> mode-shared complex MLP
    (spectral_layers x [complex linear + ComplexReLU("real")] + wout) -> inverse SHT."""

    def __init__(self, forward_transform, inverse_transform, embed_dim, sparsity_threshold=0.0, hidden_size_factor=2,
                 use_complex_network=True, use_complex_kernels=False, complex_activation="real", bias=False,
                 spectral_layers=1, drop_rate=0.0, precision=None):
        super().__init__()
        if bias:
            raise NotImplementedError("SpectralAttentionS2 bias is never enabled by the reference nets (sfnonet.py:89)")
        if complex_activation != "real":
            raise NotImplementedError("only complex_activation='real' is on the MSFNO hot path")
        if drop_rate > 0.0:
            raise NotImplementedError("spectral dropout is not used by the reference nets")
        self.embed_dim = embed_dim
        self.sparsity_threshold = sparsity_threshold
        self.hidden_size = int(hidden_size_factor * embed_dim)
        self.scale = 0.02
        self.spectral_layers = spectral_layers
        self.modes_lat = forward_transform.lmax
        self.modes_lon = forward_transform.mmax
        assert inverse_transform.lmax == self.modes_lat
        assert inverse_transform.mmax == self.modes_lon
        # the reference keeps only bound .forward handles (layers.py:570-571) so the transforms do not
        # appear as sub-modules / in the state_dict; mirror that with non-module attributes
        object.__setattr__(self, "_sht", forward_transform)
        object.__setattr__(self, "_isht", inverse_transform)
        self.forward_transform = forward_transform.forward
        self.inverse_transform = inverse_transform.forward
        w = [self.scale * torch.randn(embed_dim, self.hidden_size, 2)]
        for _ in range(1, spectral_layers):
            w.append(self.scale * torch.randn(self.hidden_size, self.hidden_size, 2))
        self.w = nn.ParameterList([nn.Parameter(t) for t in w])
        self.wout = nn.Parameter(self.scale * torch.randn(self.hidden_size, embed_dim, 2))
        self.drop = nn.Identity()
        self.activation = ComplexReLU(mode=complex_activation, bias_shape=(self.hidden_size, 1, 1))
        self.precision = precision

    def spectral(self, a_pm):
        tier = self.precision if self.precision is not None else _precision.get_precision()
        prec = _lib.PREC_TF32 if tier == "tf32" else _lib.PREC_FP32
        ws = [w.float().contiguous() for w in self.w]
        wout = self.wout.float().contiguous()
        if not torch.is_grad_enabled():
            return self._spectral_inference(a_pm, prec, wout, ws)
        return _SpecAttn.apply(a_pm, self._sht, prec, wout, *ws)

    def _spectral_inference(self, a_pm, prec, wout, ws):
        """No-autograd path: the workspace (packed real weights + hidden activations) is kept per (batch, device, tier)
        and the packed weights are reused while the parameters are unchanged (data_ptr / _version)."""
        plan = self._sht._get_plan(a_pm.device)
        B, C = a_pm.shape[0], a_pm.shape[2] // 2
        hid, nl = wout.shape[0], len(ws)
        key = (B, str(a_pm.device), prec)
        wkey = (_lib.cache_epoch(),) + tuple((id(w), w.data_ptr(), w._version) for w in (self.wout, *self.w))
        cache = self.__dict__.setdefault("_ws_cache", {})
        ent = cache.get(key)
        if ent is None:
            buf = torch.empty(lib.msfno_specattn_ws_floats(plan.h, B, C, hid, nl), dtype=torch.float32, device=a_pm.device)
            ent = cache[key] = [buf, None]
        flags = prec | (4 if ent[1] == wkey else 0)
        out = torch.empty((B, 2 * C, plan.P), dtype=torch.float32, device=a_pm.device)
        warr = (ctypes.c_void_p * nl)(*[w.data_ptr() for w in ws])
        check(lib.msfno_specattn_fwd(plan.h, ptr(a_pm), warr, nl, ptr(wout), ptr(out), ptr(ent[0]), B, C, hid, flags,
                                     _stream()), "specattn_fwd")
        ent[1] = wkey
        _lib.note_cached(ent[0])
        return out

    @_lib.on_input_device
    def forward_mlp(self, xr):
        """Reference-compatible entry: xr real view [B,C,L,M,2] -> [B,C,L,M,2]."""
        B, C = xr.shape[0], xr.shape[1]
        a = relayout(xr.float(), self._sht, _lib.LAYOUT_STD, _lib.LAYOUT_PM, B, C)
        c = self.spectral(a)
        return relayout(c, self._sht, _lib.LAYOUT_CM, _lib.LAYOUT_STD, B, C)

    @_lib.on_input_device
    def forward(self, x, in_scale=None, in_shift=None, **epilogue):
        _require_cuda(x, "SpectralAttentionS2")
        dtype = x.dtype
        a = self._sht.forward_packed(x, in_scale, in_shift)
        c = self.spectral(a)
        y = self._isht.inverse_packed(c, **epilogue)
        return y.to(dtype)
