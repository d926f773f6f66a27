"""Spherical losses against the oracle restatement and, where a copy of the reference is present,
against the reference's own losses.py (/root/reference MSFNO/Models/losses.py:6-37,80-155): values and gradients."""
import pytest
import torch

from conftest import rel_l2
from oracle import ref_import, sfno_oracle

pytestmark = pytest.mark.gpu

import msfno_b200

TOL = 1e-5


def _pair(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g), torch.randn(*shape, generator=g)


@pytest.mark.parametrize("shape", [(2, 5, 24, 48), (1, 3, 91, 181), (1, 73, 721, 1440)])
@pytest.mark.parametrize("cls,kw", [("L2Sphere", dict(relative=True, squared=False)), ("L2Sphere", dict(relative=False, squared=True)),
                                    ("L2Sphere_noSine", dict(relative=True, squared=True)), ("CosineMSELoss", dict(reduction="mean")),
                                    ("CosineMSELoss", dict(reduction="sum"))])
def test_losses_forward_backward(shape, cls, kw):
    a, b = _pair(shape, 3)
    ao = a.clone().requires_grad_(True)
    if cls == "CosineMSELoss":
        want = sfno_oracle.cosine_mse(ao, b, **kw)
    else:
        want = sfno_oracle.l2_sphere(ao, b, sine=(cls == "L2Sphere"), **kw)
    want.backward()
    ag = a.cuda().requires_grad_(True)
    launches = msfno_b200._lib.lib.msfno_launch_count()
    got = getattr(msfno_b200, cls)(**kw)(ag, b.cuda())
    got.backward()
    assert msfno_b200._lib.lib.msfno_launch_count() >= launches + 2
    assert abs(float(got) - float(want)) <= TOL * abs(float(want)), (float(got), float(want))
    assert rel_l2(ag.grad, ao.grad) < TOL
    if ref_import.available() and shape[-1] <= 181:
        ref = ref_import.load()
        ref_import.use_harmonics(ref.th_shim)
        ref_val = getattr(ref.losses, cls)(**kw)(a, b)
        assert abs(float(got) - float(ref_val)) <= TOL * abs(float(ref_val))


def test_l2sphere_reduction_none_matches_reference_shape():
    a, b = _pair((2, 3, 12, 24), 5)
    got = msfno_b200.L2Sphere(relative=False, reduction="none")(a.cuda(), b.cuda())
    assert got.shape == a.shape
