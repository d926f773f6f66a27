"""ORACLE helper (build container only): import the UNMODIFIED reference modules from
/root/reference so the restatement in sfno_oracle.py / th_shim.py can be checked against
the real thing and golden vectors can be generated (oracle/gen_golden.py).

/root/reference does not exist on the GPU box: nothing under tests -m gpu, smoke() or
bench.py calls this.  The reference needs a handful of absent third-party modules purely at
import time; they are stubbed in sys.modules.  torch_harmonics
is replaced by oracle/th_shim.py (un-vendored dependency, conda_environment.yml:62).
"""
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("MSFNO_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "MSFNO", "Models", "sfno"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
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
    from MSFNO.utils import Attributes

    _loaded = types.SimpleNamespace(sfnonet=sfnonet, layers=layers, contractions=contractions,
                                    activations=activations, Attributes=Attributes)
    return _loaded
