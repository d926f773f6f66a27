"""msfno_b200 -- B200-native (sm_100a) implementation of the MSFNO spectral hot path behind the
reference's own torch.nn.Module surface.  The directory is named
`modulated-spherical-fourier-neural-operator_b200`; import it as `msfno_b200` (repo-root shim).

Public surface (mirrors /root/reference MSFNO/Models/sfno/{sfnonet,layers}.py and torch_harmonics):
    RealSHT, InverseRealSHT, quadrature, legendre
    SpectralConvS2, SpectralAttentionS2, ComplexReLU, MLP
    SpectralFilterLayer, FiLM, FourierNeuralOperatorBlock[_Filmed],
    FourierNeuralOperatorNet[_Filmed], Film_wrapper, FeedForward
There is no CPU fallback: importing requires libmsfno_b200.so (see __graft_entry__.build()).
"""
from . import _lib  # noqa: F401  (loads the shared library; raises loudly when it is missing)
from ._lib import invalidate_caches
from . import legendre, quadrature
from .precision import get_fp32_engine, get_precision, set_fp32_engine, set_precision
from .sht import InverseRealSHT, RealSHT
from .layers import MLP, ComplexReLU, DropPath, SpectralAttentionS2, SpectralConvS2, trunc_normal_
from .sfnonet import (FeedForward, FiLM, Film_wrapper, FourierNeuralOperatorBlock, FourierNeuralOperatorBlock_Filmed,
                      FourierNeuralOperatorNet, FourierNeuralOperatorNet_Filmed, SpectralFilterLayer)
from . import harmonics
from . import losses
from .losses import CosineMSELoss, L2Sphere, L2Sphere_noSine
from .pipeline import HostPipeline
from .graph import GraphedForward
from .trainer import Trainer, TrainerConfig, freeze_backbone

__all__ = [
    "RealSHT", "InverseRealSHT", "quadrature", "legendre", "harmonics", "set_precision", "get_precision", "set_fp32_engine", "get_fp32_engine", "invalidate_caches",
    "SpectralConvS2", "SpectralAttentionS2", "ComplexReLU", "MLP", "DropPath", "trunc_normal_",
    "SpectralFilterLayer", "FiLM", "FourierNeuralOperatorBlock", "FourierNeuralOperatorBlock_Filmed",
    "FourierNeuralOperatorNet", "FourierNeuralOperatorNet_Filmed", "Film_wrapper", "FeedForward", "HostPipeline", "GraphedForward", "Trainer", "TrainerConfig", "freeze_backbone", "losses", "L2Sphere", "L2Sphere_noSine", "CosineMSELoss",
]
