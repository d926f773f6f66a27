"""1x1-convolution front end of msfno_conv1x1_fwd (inference path of the channel MLPs).

Mirrors what the reference does with nn.Conv2d(.., 1) + bias + nn.GELU + residual adds + torch.cat
(/root/reference MSFNO/Models/sfno/layers.py:161-168, sfnonet.py:232,249,671,682-684) in ONE kernel launch."""
import weakref

import torch

from . import _lib
from . import precision as _precision
from ._lib import check, lib, ptr

_pad_cache = {}   # id(weight tensor) -> (weakref to it, {(version, tier, cols): padded tensor})


def round_tf32(t):
    """Round-to-nearest (ties away) to TF32, as cvt.rna.tf32.f32 does: tensor-core operands are otherwise truncated,
    which biases every dot product towards zero."""
    i = t.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def padded_weight(weight, cols=None):
    """[Cout, Cin, 1, 1] (or [Cout, Cin]) -> contiguous [Cout, ceil4(n)] of the input-channel columns `cols` = (lo, hi)
    (default all), zero padded, TF32-rounded in the tensor-core tier.  Cached per tensor OBJECT (weak reference),
    version and tier -- never per address: freed parameter memory is routinely reused by the next model."""
    tf32 = _precision.get_precision() == "tf32"
    ent = _pad_cache.get(id(weight))
    if ent is None or ent[0]() is not weight:
        wid = id(weight)
        ent = (weakref.ref(weight, lambda _r, k=wid: _pad_cache.pop(k, None)), {})
        _pad_cache[wid] = ent
    key = (weight._version, _lib.cache_epoch(), tf32, cols)
    hit = ent[1].get(key)
    if hit is not None:
        return _lib.note_cached(hit)
    w2 = weight.detach().reshape(weight.shape[0], -1).float()
    if cols is not None:
        w2 = w2[:, cols[0]:cols[1]]
    cin = w2.shape[1]
    ld = (cin + 3) // 4 * 4
    if ld != cin:
        w2 = torch.nn.functional.pad(w2, (0, ld - cin))
    w2 = w2.contiguous()
    if tf32:
        w2 = round_tf32(w2)
    for k in [k for k in ent[1] if k[:2] != key[:2]]:
        del ent[1][k]   # drop entries of older versions of this tensor
    ent[1][key] = w2
    return _lib.note_cached(w2)


@_lib.on_input_device
def conv1x1(x, w, cin, bias=None, act_gelu=False, add=None, x2=None, w2=None, cin2=0, per_sample_w=False,
            per_sample_bias=False, final=False, w_rounded=False, gelu_grad_of=None):
    """y = act(conv1x1(x, w) [+ conv1x1(x2, w2)] + bias) + add   for contiguous NCHW fp32 CUDA tensors.
    w: [Cout, ld] (or [B, Cout, ld] with per_sample_w) zero-padded rows; bias [Cout] (or [B, Cout]);
    add: [B or 1, Cout, H, W].  gelu_grad_of=h [B, Cout, H, W]: y = (conv + bias) * gelu'(h) instead (activation adjoint)."""
    if gelu_grad_of is not None:
        assert add is None and not act_gelu
        add = gelu_grad_of
    B, _, H, W = x.shape
    HW = H * W
    cout = w.shape[-2]
    y = torch.empty((B, cout, H, W), dtype=torch.float32, device=x.device)
    prec = _lib.PREC_FP32
    if _precision.get_precision() == "tf32":
        prec = _lib.PREC_TF32 | (0 if final else 2)   # bit 1: TF32-round the output (it feeds another tensor-core GEMM)
        if per_sample_w and not w_rounded:
            w = round_tf32(w)
    add_bs = 0
    if add is not None:
        add = add.contiguous()
        add_bs = cout * HW if add.shape[0] == B and B > 1 else (0 if add.shape[0] == 1 else cout * HW)
    check(lib.msfno_conv1x1_fwd(ptr(x), x.shape[1] * HW, cin, ptr(w), w.shape[-1], (cout * w.shape[-1]) if per_sample_w else 0,
                                ptr(x2), (x2.shape[1] * HW) if x2 is not None else 0, cin2, ptr(w2),
                                w2.shape[-1] if w2 is not None else 0, ptr(bias), cout if per_sample_bias else 0, ptr(add),
                                add_bs, ptr(y), B, cout, HW, 2 if gelu_grad_of is not None else (1 if act_gelu else 0), prec,
                                torch.cuda.current_stream().cuda_stream), "conv1x1_fwd")
    return y


def mlp1x1_supported(chid, cout, HW):
    """Shapes the fused two-layer kernel (msfno_mlp1x1_fwd, tensor-core tier) takes."""
    ok_hid = chid % 32 == 0 and (chid <= 256 or (chid % 256 == 0 and chid <= 1024))
    return _precision.get_precision() == "tf32" and ok_hid and cout <= 256 and HW % 4 == 0


@_lib.on_input_device
def mlp1x1(x, w1, cin, b1, w2, b2=None, add=None, x2=None, w1b=None, cin2=0, per_sample_w1=False, per_sample_b1=False,
           final=False, w1_rounded=False, stats=None, out=None, per_sample_b2=False):
    """y = conv1x1(gelu(conv1x1(x, w1) [+ conv1x1(x2, w1b)] + b1), w2) + b2 + add in ONE kernel; the hidden activation
    never reaches HBM.  w1: [Chid, ld] (or [B, Chid, ld]), w1b: [Chid, ld2], w2: [Cout, ld3] zero-padded rows.
    stats: optional [B*Cout, 2] float64 tensor that receives the plane sums / sums of squares of y.
    out: optional preallocated [B, Cout, H, W] result.  It may alias x2 (the in-place rollout step): a CTA reads every
    operand block of a pixel tile before it stores that tile, and no other CTA touches those pixels."""
    B, _, H, W = x.shape
    HW = H * W
    chid, cout = w1.shape[-2], w2.shape[-2]
    if out is not None:
        if out.shape != (B, cout, H, W) or out.dtype != torch.float32 or not out.is_contiguous() or out.device != x.device:
            raise RuntimeError("mlp1x1: out must be a contiguous float32 [B, Cout, H, W] tensor on x's device")
        y = out
    else:
        y = torch.empty((B, cout, H, W), dtype=torch.float32, device=x.device)
    if per_sample_w1 and not w1_rounded:
        w1 = round_tf32(w1)
    add_bs = 0
    if add is not None:
        add = add.contiguous()
        add_bs = cout * HW if add.shape[0] == B and B > 1 else 0
    check(lib.msfno_mlp1x1_fwd(ptr(x), x.shape[1] * HW, cin, ptr(w1), w1.shape[-1], (chid * w1.shape[-1]) if per_sample_w1 else 0,
                               ptr(x2), (x2.shape[1] * HW) if x2 is not None else 0, cin2, ptr(w1b),
                               w1b.shape[-1] if w1b is not None else 0, ptr(b1), chid if per_sample_b1 else 0, chid, ptr(w2),
                               w2.shape[-1], ptr(b2), cout if per_sample_b2 else 0, ptr(add), add_bs, ptr(y), ptr(stats), B, cout, HW, 0 if final else 2,
                               torch.cuda.current_stream().cuda_stream), "mlp1x1_fwd")
    return y
