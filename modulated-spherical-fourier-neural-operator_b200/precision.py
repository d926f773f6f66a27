"""Precision tiers of the hot path (north_star: <= 1e-5 rel-L2 for "fp32", <= 2e-3 for the tensor-core tier).

"fp32": every contraction accumulates fp32 products of fp32 operands (CUDA-core FFMA kernels; library
        1x1 convolutions with TF32 disabled).
"tf32": the GEMM-shaped stages (spectral complex MLP, 1x1-conv MLPs) run on the tensor cores in TF32.
The tier is an explicit setting, never autocast-driven: the reference forces fp32 inside the transforms
(/root/reference MSFNO/Models/sfno/layers.py:403-407,418-422,627-639)."""
import contextlib

import torch

_TIER = "fp32"


def set_precision(tier):
    global _TIER
    if tier not in ("fp32", "tf32"):
        raise ValueError("precision tier must be 'fp32' or 'tf32'")
    _TIER = tier


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
