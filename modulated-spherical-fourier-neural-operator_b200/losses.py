"""Spherical losses: drop-in replacements for the classes of /root/reference MSFNO/Models/losses.py that the trainer
uses on the [B, 73, 721, 1440] forecast: CosineMSELoss (:6-37), L2Sphere (:80-117),
L2Sphere_noSine (:119-155).  Same constructor arguments and return values.

The reference rebuilds the Gauss-Legendre weights with numpy on the HOST at every call (losses.py:90,129) and runs
six elementwise passes with four full-size temporaries; here the latitude weights are a cached device vector and the
two plane sums (sum w (prd-tar)^2, sum w tar^2) come from ONE pass over the two tensors (msfno_weighted_sq_sums, fp64
accumulation), with a one-pass backward (msfno_weighted_diff).  CUDA only, like the rest of the package."""
import math

import torch
import torch.nn as nn

from . import _lib
from . import quadrature as _quadrature
from ._lib import check, lib, ptr

_WCACHE = {}


def _stream():
    return torch.cuda.current_stream().cuda_stream


def latitude_weights(kind, H, device):
    """[H] fp32 device vector, built once per (kind, H, device).
    "l2sphere": |w_quad cos(lat)| (losses.py:90-94), "nosine": w_quad (:129-132), ("cosine", eps): CosineMSELoss (:16-19)."""
    key = (kind, H, str(device))
    w = _WCACHE.get(key)
    if w is None:
        if kind in ("l2sphere", "nosine"):
            wq = torch.tensor(_quadrature.legendre_gauss_weights(H, -1, 1)[1], dtype=torch.float32)
            if kind == "l2sphere":
                wq = torch.abs(wq * torch.cos(torch.linspace(-math.pi / 2, math.pi / 2, H, dtype=torch.float32)))
            w = wq
        else:
            eps = kind[1]
            w = torch.clamp(torch.cos(torch.linspace(-math.pi / 2, math.pi / 2, H, dtype=torch.float32)), min=0.0) + eps
            w = w / w.sum(dim=-1, keepdim=True)
        w = _WCACHE[key] = w.contiguous().to(device)
    return w


class _WeightedSqSums(torch.autograd.Function):
    """(prd, tar, wlat) -> [B, C, 2] float64: (sum w (prd - tar)^2, sum w tar^2) per plane."""

    @staticmethod
    def forward(ctx, prd, tar, wlat):
        B, C, H, W = prd.shape
        out = torch.empty((B, C, 2), dtype=torch.float64, device=prd.device)
        check(lib.msfno_weighted_sq_sums(ptr(prd), ptr(tar), ptr(wlat), ptr(out), B * C, H, W, _stream()), "weighted_sq_sums")
        ctx.save_for_backward(prd, tar, wlat)
        return out

    @staticmethod
    def backward(ctx, g):
        prd, tar, wlat = ctx.saved_tensors
        B, C, H, W = prd.shape
        gp = gt = None
        # d/dprd sum w (prd - tar)^2 = 2 w (prd - tar); the target of a loss takes no gradient in the reference's use
        if ctx.needs_input_grad[0]:
            coef = (2.0 * g[..., 0]).float().contiguous()
            gp = torch.empty_like(prd)
            check(lib.msfno_weighted_diff(ptr(prd), ptr(tar), ptr(wlat), ptr(coef), ptr(gp), B * C, H, W, _stream()), "weighted_diff")
        if ctx.needs_input_grad[1]:
            raise RuntimeError("msfno_b200 losses: gradients with respect to the target are not implemented")
        return gp, gt, None


@_lib.on_input_device
def weighted_sq_sums(prd, tar, wlat):
    if not prd.is_cuda:
        raise RuntimeError("msfno_b200 losses run on CUDA only (no CPU fallback)")
    assert prd.shape == tar.shape and prd.dim() == 4
    return _WeightedSqSums.apply(prd.contiguous().float(), tar.contiguous().float(), wlat)


class _L2SphereBase(nn.Module):
    _kind = "l2sphere"

    def __init__(self, relative=True, squared=False, reduction="sum", dampening=None):
        super().__init__()
        self.relative = relative
        self.squared = squared
        self.reduction = reduction

    def forward(self, prd, tar):
        B, C, H, W = prd.shape
        wlat = latitude_weights(self._kind, H, prd.device)
        if self.reduction == "none":   # the un-reduced [B, C, H, W] field (losses.py:96-101)
            sw = wlat[None, None, :, None]
            loss = sw * (prd - tar) ** 2
            if self.relative:
                loss = loss / (sw * tar ** 2).sum(dim=(-1, -2))
            return loss
        sums = weighted_sq_sums(prd, tar, wlat)
        loss = sums[..., 0]
        if self.relative:
            loss = loss / sums[..., 1]
        if not self.squared:
            loss = torch.sqrt(loss)
        return loss.sum().to(prd.dtype)   # "mean" and "sum" both return the sum in the reference (:110-115)


class L2Sphere(_L2SphereBase):
    """losses.py:80-117."""
    _kind = "l2sphere"


class L2Sphere_noSine(_L2SphereBase):
    """losses.py:119-155."""
    _kind = "nosine"


class CosineMSELoss(nn.Module):
    """losses.py:6-37: mean / sum of cos(lat)-weighted squared errors."""

    def __init__(self, reduction="mean", eps=1e-4):
        super().__init__()
        self.reduction = reduction
        self.eps = eps

    def forward(self, x, y):
        B, C, H, W = x.shape
        wlat = latitude_weights(("cosine", float(self.eps)), H, x.device)
        if self.reduction == "none":
            return (x - y) ** 2 * wlat[None, None, :, None]
        total = weighted_sq_sums(x, y, wlat)[..., 0].sum()
        if self.reduction == "mean":
            return (total / (B * C * H * W)).to(x.dtype)
        return (total / W).to(x.dtype)
