"""ORACLE helper (build container only): import the UNMODIFIED reference modules from
/root/reference so the restatement in sfno_oracle.py / th_shim.py can be checked against
the real thing and golden vectors can be generated (oracle/gen_golden.py).

/root/reference does not exist on the GPU box: there the byte-identical copy staged by
oracle/build_ref.sh under the git-ignored oracle/_ref/ is imported instead (bench.py --impl reference,
tests/test_gpu_dropin.py).  The reference needs a handful of absent third-party modules purely at
import time; they are stubbed in sys.modules.  torch_harmonics
is replaced by oracle/th_shim.py (un-vendored dependency, conda_environment.yml:62).
"""
import importlib.machinery
import os
import sys
import types

import numpy as np

# the mounted reference tree (build container), else the byte-identical copy staged by oracle/build_ref.sh in the
# git-ignored oracle/_ref/ (which travels to the GPU box with the snapshot)
_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
_MOUNT = os.environ.get("MSFNO_REFERENCE_ROOT", "/root/reference")
REFERENCE_ROOT = _MOUNT if os.path.isdir(os.path.join(_MOUNT, "MSFNO", "Models", "sfno")) else _STAGED


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "MSFNO", "Models", "sfno"))


def use_harmonics(mod):
    """Point the reference's `import torch_harmonics as harmonics` at `mod` (oracle.th_shim on the CPU legs,
    msfno_b200.harmonics for the drop-in tests): the reference looks the classes up at construction time."""
    st = sys.modules["torch_harmonics"]
    st.RealSHT, st.InverseRealSHT, st.quadrature = mod.RealSHT, mod.InverseRealSHT, mod.quadrature
    if hasattr(mod, "legendre"):
        st.legendre = mod.legendre


def _stub(name, **attrs):
    m = types.ModuleType(name)
    # a real spec: importlib.util.find_spec() raises ValueError on a sys.modules entry whose __spec__ is None, and
    # torch._dynamo calls it for "xarray" & co. when it is first imported (e.g. by torch.optim.Adam) later in the process
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_loaded = None


def load():
    """Returns a namespace with the reference's own classes (unmodified code)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference tree not mounted at %s" % REFERENCE_ROOT)
    from . import th_shim

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    if "xarray" not in sys.modules:
        _stub("xarray")  # sfnonet.py:8, unused there
    _stub("numpy.lib.arraypad", pad=np.pad)  # layers.py:9, unused, gone in NumPy 2
    if "climetlab" not in sys.modules:
        c = _stub("climetlab")
        u = _stub("climetlab.utils")
        h = _stub("climetlab.utils.humanize", seconds=lambda s: "%.1fs" % s)
        c.utils, u.humanize = u, h  # MSFNO/utils.py:6
    if "torch_geometric" not in sys.modules:
        g = _stub("torch_geometric")
        gn = _stub("torch_geometric.nn", GCNConv=object)
        gp = _stub("torch_geometric.nn.pool", global_mean_pool=None)
        g.nn, gn.pool = gn, gp  # gcn/gcn.py:5-6
    _stub("torch_harmonics", RealSHT=th_shim.RealSHT, InverseRealSHT=th_shim.InverseRealSHT,
          quadrature=th_shim.quadrature, legendre=th_shim.legendre)

    from MSFNO.Models.sfno import sfnonet, layers, contractions, activations
    from MSFNO.Models import losses
    from MSFNO.utils import Attributes

    _loaded = types.SimpleNamespace(sfnonet=sfnonet, layers=layers, contractions=contractions,
                                    activations=activations, losses=losses, Attributes=Attributes, th_shim=th_shim)
    return _loaded
