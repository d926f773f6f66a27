"""Namespace with the part of the `torch_harmonics` API the reference uses, so that
`sys.modules["torch_harmonics"] = msfno_b200.harmonics` makes the UNMODIFIED reference files
(sfnonet.py:45 `import torch_harmonics as harmonics`, losses.py:3) run on the sm_100a kernels."""
from . import legendre, quadrature  # noqa: F401
from .sht import InverseRealSHT, RealSHT  # noqa: F401
