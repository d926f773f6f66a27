"""Precision tiers of the hot path (north_star: <= 1e-5 rel-L2 for "fp32", <= 2e-3 for the tensor-core tier).

"fp32": every contraction is fp32-grade with fp32 accumulation.  Default engine "tc3x": the GEMM-shaped stages (spectral
        complex MLP, Legendre contractions, 1x1 convs, all adjoints) run on the tcgen05 tensor cores as 3xTF32 -- three
        MMAs per k-step on hi / lo operand splits made in shared memory (csrc/gemm_tc.cu); engine "ffma": CUDA-core FFMA
        kernels (set_fp32_engine).  The longitude transforms are the four-step FFT kernels.
"tf32": the GEMM-shaped stages run as plain TF32 MMAs on TF32-rounded operands, the longitude transforms as TF32 DFT GEMMs.
The tier is an explicit setting, never autocast-driven: the reference forces fp32 inside the transforms
(/root/reference MSFNO/Models/sfno/layers.py:403-407,418-422,627-639)."""
import contextlib

import torch

_TIER = "fp32"
# (the default tier is applied to the PyTorch library flags at import, see set_precision)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def set_precision(tier):
    global _TIER
    if tier not in ("fp32", "tf32"):
        raise ValueError("precision tier must be 'fp32' or 'tf32'")
    _TIER = tier
    # the PyTorch library ops either side of the path (cuDNN / cuBLAS 1x1 convolutions of layers that need weight
    # gradients, and every autograd backward of them, which runs outside any forward-time scope) follow the tier:
    # PyTorch allows TF32 in cuDNN convolutions by default, which alone puts ~1e-3 on full-size gradients
    torch.backends.cudnn.allow_tf32 = tier == "tf32"
    torch.backends.cuda.matmul.allow_tf32 = tier == "tf32"


def set_fp32_engine(engine):
    """Engine of the fp32 tier's GEMM-shaped stages: "tc3x" (3xTF32 on the tensor cores, default) or "ffma"."""
    from . import _lib
    codes = {"tc3x": _lib.FP32_ENGINE_TC3X, "ffma": _lib.FP32_ENGINE_FFMA}
    if engine not in codes:
        raise ValueError("fp32 engine must be 'tc3x' or 'ffma'")
    _lib.check(_lib.lib.msfno_set_fp32_engine(codes[engine]), "set_fp32_engine")


def get_fp32_engine():
    from . import _lib
    return "ffma" if _lib.lib.msfno_get_fp32_engine() == _lib.FP32_ENGINE_FFMA else "tc3x"


_LEGENDRE_TC = True


def set_legendre_on_tensor_cores(flag):
    """tf32 tier only: run the Legendre contractions of the forward transforms on the tensor cores (default) or keep
    them on the fp32 CUDA-core engine (more headroom under the 2e-3 tolerance, ~0.8 ms per step slower)."""
    global _LEGENDRE_TC
    _LEGENDRE_TC = bool(flag)


def legendre_on_tensor_cores():
    return _LEGENDRE_TC


def get_precision():
    return _TIER


@contextlib.contextmanager
def library_scope():
    """Make the PyTorch library ops either side of the path (1x1 convs) follow the selected tier."""
    want = _TIER == "tf32"
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = want
    torch.backends.cuda.matmul.allow_tf32 = want
    try:
        yield
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
