"""RealSHT / InverseRealSHT: drop-in replacements for the `torch_harmonics` classes the reference
builds at /root/reference MSFNO/Models/sfno/sfnonet.py:537-548 and calls at layers.py:405,421,629,638.

Same constructor signature, attributes (nlat, nlon, lmax, mmax, grid, norm, csphase) and the same
assignable non-persistent buffers (`weights` [mmax,lmax,nlat] for the analysis, `pct` for the synthesis)
that the reference rescales in place after construction (sfnonet.py:551-555).  The arithmetic runs in
hand-written sm_100a kernels behind the C ABI (include/msfno_b200.h): a shared-memory Stockham FFT
along longitude that only produces / consumes the mmax orders kept, and a grouped Legendre GEMM over
the azimuthal order.  There is no CPU path: calling forward on a non-CUDA tensor raises.

Besides the reference-compatible complex [..., lmax, mmax] interface the modules expose the packed
internal layouts (`forward_packed` / `inverse_packed`) that the spectral layers use to avoid the
tril gather/scatter and complex<->real shuffles of the reference (layers.py:406-413).
"""
import ctypes

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from . import legendre as _legendre
from . import precision as _precision
from . import quadrature as _quadrature
from ._lib import check, lib, ptr


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _require_cuda(t, who):
    if not t.is_cuda:
        raise RuntimeError("%s: msfno_b200 runs on CUDA (sm_100a) only; got a %s tensor -- there is no CPU fallback"
                           % (who, t.device.type))


class _Plan:
    """Owns one msfno_plan and tracks which table tensor it was built from."""

    def __init__(self, nlat, nlon, lmax, mmax, device):
        self.device = device
        handle = ctypes.c_void_p()
        with torch.cuda.device(device):
            check(lib.msfno_plan_create(ctypes.byref(handle), nlat, nlon, lmax, mmax), "plan_create")
        self.h = handle
        self.kpad = lib.msfno_plan_query(self.h, _lib.Q_KPAD)
        self.mlim = lib.msfno_plan_query(self.h, _lib.Q_MLIM)
        self.P = lib.msfno_plan_query(self.h, _lib.Q_NPACK)
        self.ntril = lib.msfno_plan_query(self.h, _lib.Q_NTRIL)
        self.table_key = None
        self.table_ref = None
        self.precision = _lib.PREC_FP32

    def set_table(self, table, analysis):
        # identity + version of the tensor object; the plan keeps the tensor alive (self.table_ref) so its address
        # cannot be recycled for a different table while the key is still in use
        key = (id(table), table._version, _lib.cache_epoch())
        if key != self.table_key or self.table_ref is not table:
            t = table.detach()
            if t.dtype != torch.float32 or not t.is_contiguous():
                t = t.float().contiguous()
            with torch.cuda.device(self.device):
                check(lib.msfno_plan_set_table(self.h, ptr(t), 1 if analysis else 0, _stream()), "plan_set_table")
            self.table_key = key
            self.table_ref = table

    def __del__(self):
        try:
            if self.h:
                lib.msfno_plan_destroy(self.h)
                self.h = None
        except Exception:
            pass


# ------------------------------------------------------------------------------- autograd functions
class _SHTForward(torch.autograd.Function):
    """x [B,C,nlat,nlon] -> coef_pm [B,P,2C]  (msfno_sht_fwd / msfno_sht_bwd)."""

    @staticmethod
    def forward(ctx, x, mod, in_scale, in_shift):
        plan = mod._get_plan(x.device)
        B, C = x.shape[0], x.shape[1]
        coef = torch.empty((B, plan.P, 2 * C), dtype=torch.float32, device=x.device)
        ws = torch.empty(lib.msfno_sht_ws_floats(plan.h, B, C), dtype=torch.float32, device=x.device)
        check(lib.msfno_sht_fwd(plan.h, ptr(x), ptr(in_scale), ptr(in_shift), ptr(coef), ptr(ws), B, C, _stream()), "sht_fwd")
        ctx.mod, ctx.shape = mod, x.shape
        ctx.save_for_backward(in_scale)
        return coef

    @staticmethod
    def backward(ctx, g):
        (in_scale,) = ctx.saved_tensors
        plan = ctx.mod._get_plan(g.device)
        B, C = ctx.shape[0], ctx.shape[1]
        g = g.contiguous()
        gx = torch.empty(ctx.shape, dtype=torch.float32, device=g.device)
        ws = torch.empty(lib.msfno_sht_ws_floats(plan.h, B, C), dtype=torch.float32, device=g.device)
        check(lib.msfno_sht_bwd(plan.h, ptr(g), ptr(in_scale), ptr(gx), ptr(ws), B, C, _stream()), "sht_bwd")
        return gx, None, None, None


class _ISHTForward(torch.autograd.Function):
    """coef_cm [B,2C,P] -> y [B,C,nlat,nlon]  (msfno_isht_fwd / msfno_isht_bwd).
    The fused epilogue (skip add, GELU, statistics) is only available without autograd."""

    @staticmethod
    def forward(ctx, coef, mod):
        plan = mod._get_plan(coef.device)
        B, C = coef.shape[0], coef.shape[1] // 2
        y = torch.empty((B, C, mod.nlat, mod.nlon), dtype=torch.float32, device=coef.device)
        ws = torch.empty(lib.msfno_sht_ws_floats(plan.h, B, C), dtype=torch.float32, device=coef.device)
        check(lib.msfno_isht_fwd(plan.h, ptr(coef), ptr(y), ptr(ws), B, C, None, 0, None, _stream()), "isht_fwd")
        ctx.mod, ctx.BC = mod, (B, C)
        return y

    @staticmethod
    def backward(ctx, gy):
        plan = ctx.mod._get_plan(gy.device)
        B, C = ctx.BC
        gy = gy.contiguous()
        g = torch.empty((B, 2 * C, plan.P), dtype=torch.float32, device=gy.device)
        ws = torch.empty(lib.msfno_sht_ws_floats(plan.h, B, C), dtype=torch.float32, device=gy.device)
        check(lib.msfno_isht_bwd(plan.h, ptr(gy), ptr(g), ptr(ws), B, C, _stream()), "isht_bwd")
        return g, None


class _Relayout(torch.autograd.Function):
    """Coefficient layout change (msfno_coef_relayout); its adjoint is the reverse relayout."""

    @staticmethod
    def forward(ctx, src, mod, sl, dl, B, C):
        plan = mod._get_plan(src.device)
        if dl == _lib.LAYOUT_STD:
            dst = torch.empty((B, C, mod.lmax, mod.mmax, 2), dtype=torch.float32, device=src.device)
        elif dl == _lib.LAYOUT_PM:
            dst = torch.empty((B, plan.P, 2 * C), dtype=torch.float32, device=src.device)
        else:
            dst = torch.empty((B, 2 * C, plan.P), dtype=torch.float32, device=src.device)
        check(lib.msfno_coef_relayout(plan.h, ptr(src), sl, ptr(dst), dl, B, C, _stream()), "coef_relayout")
        ctx.args = (mod, sl, dl, B, C)
        return dst

    @staticmethod
    def backward(ctx, g):
        mod, sl, dl, B, C = ctx.args
        return _Relayout.apply(g.contiguous(), mod, dl, sl, B, C), None, None, None, None, None


def relayout(t, mod, sl, dl, B, C):
    return _Relayout.apply(t.contiguous(), mod, sl, dl, B, C)


# ------------------------------------------------------------------------------- modules
class _SHTBase(nn.Module):
    _analysis = True
    _table_name = "weights"

    def __init__(self, nlat, nlon, lmax=None, mmax=None, grid="lobatto", norm="ortho", csphase=True):
        super().__init__()
        self.nlat, self.nlon, self.grid, self.norm, self.csphase = nlat, nlon, grid, norm, csphase
        cost, w, lmax_default = _quadrature.grid_nodes(grid, nlat)
        self.lmax = lmax or lmax_default
        self.mmax = mmax or nlon // 2 + 1
        theta = np.flip(np.arccos(cost))  # colatitude ascending: row 0 = north pole
        tab = _legendre.precompute_legpoly(self.mmax, self.lmax, theta, norm=norm, inverse=not self._analysis,
                                           csphase=csphase)
        if self._analysis:
            tab = tab * w[None, None, :]
        self.register_buffer(self._table_name, torch.from_numpy(np.ascontiguousarray(tab)), persistent=False)
        self._plans = {}

    def extra_repr(self):
        return f"nlat={self.nlat}, nlon={self.nlon},\n lmax={self.lmax}, mmax={self.mmax},\n grid={self.grid}, csphase={self.csphase}"

    def _get_plan(self, device):
        key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
        plan = self._plans.get(key)
        if plan is None:
            plan = _Plan(self.nlat, self.nlon, self.lmax, self.mmax, device)
            self._plans[key] = plan
        table = getattr(self, self._table_name)
        if table.device != device:
            raise RuntimeError("%s.%s lives on %s but the input is on %s" % (type(self).__name__, self._table_name,
                                                                             table.device, device))
        tier = _lib.PREC_TF32 if (_precision.get_precision() == "tf32" and _precision.legendre_on_tensor_cores()) else _lib.PREC_FP32
        if tier != plan.precision:
            check(lib.msfno_plan_set_precision(plan.h, tier), "plan_set_precision")
            plan.precision = tier
            plan.table_key = None   # the re-laid tables are TF32-rounded in the tensor-core tier: rebuild them
        plan.set_table(table, self._analysis)
        return plan

    def __deepcopy__(self, memo):
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = {} if k == "_plans" else copy.deepcopy(v, memo)
        return new

    def __getstate__(self):
        d = dict(self.__dict__)
        d["_plans"] = {}
        return d


class RealSHT(_SHTBase):
    """Forward real SHT: real [..., nlat, nlon] -> complex [..., lmax, mmax]."""
    _analysis = True
    _table_name = "weights"

    @_lib.on_input_device
    def forward_packed(self, x, in_scale=None, in_shift=None):
        """x [B,C,nlat,nlon] fp32 CUDA -> coefficients in the PM layout [B,P,2C].  Optional fused
        per-(b,c) affine x*in_scale + in_shift (shape [B,C] or [B*C])."""
        _require_cuda(x, "RealSHT")
        assert x.dim() == 4 and x.shape[-2] == self.nlat and x.shape[-1] == self.nlon
        x = x.contiguous().float()
        if in_scale is not None:
            in_scale = in_scale.contiguous().float()
            in_shift = in_shift.contiguous().float()
        return _SHTForward.apply(x, self, in_scale, in_shift)

    @_lib.on_input_device
    def forward(self, x):
        _require_cuda(x, "RealSHT")
        assert x.shape[-2] == self.nlat
        assert x.shape[-1] == self.nlon
        lead = x.shape[:-2]
        if x.dim() == 4:
            B, C = x.shape[0], x.shape[1]
        else:
            B, C = 1, int(np.prod(lead)) if len(lead) else 1
        pm = self.forward_packed(x.reshape(B, C, self.nlat, self.nlon))
        std = relayout(pm, self, _lib.LAYOUT_PM, _lib.LAYOUT_STD, B, C)
        return torch.view_as_complex(std).reshape(*lead, self.lmax, self.mmax)


class InverseRealSHT(_SHTBase):
    """Inverse real SHT: complex [..., lmax, mmax] -> real [..., nlat, nlon]."""
    _analysis = False
    _table_name = "pct"

    @_lib.on_input_device
    def inverse_packed(self, coef_cm, skip_add=None, act_gelu=False, stats=None):
        """coefficients in the CM layout [B,2C,P] -> y [B,C,nlat,nlon].  skip_add / act_gelu / stats
        select the fused epilogue (inference only: not differentiable)."""
        _require_cuda(coef_cm, "InverseRealSHT")
        coef_cm = coef_cm.contiguous()
        if skip_add is None and not act_gelu and stats is None:
            return _ISHTForward.apply(coef_cm, self)
        if torch.is_grad_enabled() and (coef_cm.requires_grad or (skip_add is not None and skip_add.requires_grad)):
            raise RuntimeError("InverseRealSHT fused epilogue is inference-only; call under torch.no_grad()")
        plan = self._get_plan(coef_cm.device)
        B, C = coef_cm.shape[0], coef_cm.shape[1] // 2
        y = torch.empty((B, C, self.nlat, self.nlon), dtype=torch.float32, device=coef_cm.device)
        ws = torch.empty(lib.msfno_sht_ws_floats(plan.h, B, C), dtype=torch.float32, device=coef_cm.device)
        if skip_add is not None:
            skip_add = skip_add.contiguous().float()
            assert skip_add.shape == y.shape
        flags = (1 if act_gelu else 0) | (2 if _precision.get_precision() == "tf32" else 0)   # bit 1: TF32-round y
        check(lib.msfno_isht_fwd(plan.h, ptr(coef_cm), ptr(y), ptr(ws), B, C, ptr(skip_add), flags,
                                 ptr(stats), _stream()), "isht_fwd")
        return y

    @_lib.on_input_device
    def forward(self, x):
        _require_cuda(x, "InverseRealSHT")
        assert x.shape[-2] == self.lmax
        assert x.shape[-1] == self.mmax
        lead = x.shape[:-2]
        if x.dim() == 4:
            B, C = x.shape[0], x.shape[1]
        else:
            B, C = 1, int(np.prod(lead)) if len(lead) else 1
        xr = torch.view_as_real(x.to(torch.complex64)).reshape(B, C, self.lmax, self.mmax, 2)
        cm = relayout(xr, self, _lib.LAYOUT_STD, _lib.LAYOUT_CM, B, C)
        y = self.inverse_packed(cm)
        return y.reshape(*lead, self.nlat, self.nlon)
