"""Host-side logic of the product package (no GPU): tables, state_dict surface, C-ABI exports,
loud failure on CPU tensors, FFT core emulated on the CPU."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import GOLD, ROOT
from oracle import th_shim

import msfno_b200
from msfno_b200 import _lib


def test_cabi_exports_every_declared_symbol():
    names = _lib.declared_symbols()
    assert len(names) >= 25
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    assert set(names) <= exported, sorted(set(names) - exported)
    info = json.loads(_lib.lib.msfno_build_info().decode())
    assert info["arch"] == "sm_100a"


def test_library_is_sm100a_with_bulk_copy():
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    assert "UBLKCP" in sass  # cp.async.bulk (TMA engine) row loads of the FFT kernel


@pytest.mark.parametrize("grid,nlat,nlon,L,M", [("equiangular", 721, 1440, 120, 121), ("legendre-gauss", 120, 240, 120, 121),
                                                ("equiangular", 24, 48, 12, 13), ("lobatto", 13, 24, 12, 13)])
def test_tables_match_oracle(grid, nlat, nlon, L, M):
    a, b = msfno_b200.RealSHT(nlat, nlon, lmax=L, mmax=M, grid=grid), th_shim.RealSHT(nlat, nlon, lmax=L, mmax=M, grid=grid)
    assert a.weights.shape == b.weights.shape == (M, L, nlat)
    assert (a.weights - b.weights).abs().max() < 1e-12 * max(1.0, float(b.weights.abs().max()))
    a, b = msfno_b200.InverseRealSHT(nlat, nlon, lmax=L, mmax=M, grid=grid), th_shim.InverseRealSHT(nlat, nlon, lmax=L, mmax=M, grid=grid)
    assert (a.pct - b.pct).abs().max() < 1e-12 * float(b.pct.abs().max())
    assert a.lmax == b.lmax and a.mmax == b.mmax and "weights" not in a.state_dict() and "pct" not in a.state_dict()


def test_default_lmax_mmax_follow_library():
    for grid in ("equiangular", "legendre-gauss", "lobatto"):
        a, b = msfno_b200.RealSHT(16, 32, grid=grid), th_shim.RealSHT(16, 32, grid=grid)
        assert (a.lmax, a.mmax) == (b.lmax, b.mmax)
    x, w = msfno_b200.quadrature.legendre_gauss_weights(721, -1, 1)  # losses.py:90 call shape
    assert abs(w.sum() - 2) < 1e-12


def test_state_dict_keys_match_reference():
    ks = json.load(open(os.path.join(GOLD, "state_dict_keys.json")))
    for ftype in ("linear", "non-linear"):
        net = msfno_b200.FourierNeuralOperatorNet("cpu", None, filter_type=ftype, img_size=(24, 48), scale_factor=2,
                                                  in_chans=5, out_chans=5, embed_dim_sfno=16, num_layers=4)
        mine = {k: list(v.shape) for k, v in net.state_dict().items()}
        assert mine == ks["net_" + ftype]

    class Cfg:
        film_gen_type, cls, embed_dim, mlp_dim, dropout, scale_weight, repeat_film, batch_size = "mae", "x", 512, 1024, 0.0, 1, False, 2

    for fl in (1, 3):
        cfg = Cfg()
        cfg.film_layers = fl
        net = msfno_b200.FourierNeuralOperatorNet_Filmed(
            "cpu", cfg, mlp_ratio=0.25, advanced_logging=True, film_layers=fl, model_depth=6, filter_type="non-linear",
            img_size=(12, 24), scale_factor=2, in_chans=4, out_chans=4, embed_dim_sfno=256, num_layers=3, spectral_layers=2)
        mine = {k: list(v.shape) for k, v in net.state_dict().items()}
        assert mine == ks["filmed_%d" % fl]


def test_no_cpu_fallback():
    sht = msfno_b200.RealSHT(12, 24, lmax=12, mmax=13, grid="legendre-gauss").float()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sht(torch.randn(1, 2, 12, 24))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        msfno_b200.FiLM()(torch.randn(1, 2, 4, 4), torch.zeros(1, 2), torch.zeros(1, 2))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "modulated-spherical-fourier-neural-operator_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "th_shim" not in src, f


def test_fft_core_host_emulation(tmp_path):
    """csrc/fft_core.cuh (butterflies, Stockham stages, real split/merge) compiled for the host with the 32
    lanes executed in a loop, against numpy.fft."""
    core = os.path.join(ROOT, "modulated-spherical-fourier-neural-operator_b200", "csrc", "fft_core.cuh")
    hdr = tmp_path / "fft_core_host.h"
    hdr.write_text(open(core).read().replace("#include <cuda_runtime.h>", ""))
    src = open(os.path.join(ROOT, "tests", "host_emul", "fft_emul.cpp")).read()
    src = src.replace('#include "../../modulated-spherical-fourier-neural-operator_b200/csrc/fft_core_host.h"',
                      '#include "%s"' % hdr)
    cpp = tmp_path / "emul.cpp"
    cpp.write_text(src)
    exe = str(tmp_path / "emul")
    subprocess.run(["g++", "-O1", "-o", exe, str(cpp)], check=True)

    def run(mode, N, M, vals):
        inp = "%d %d %d\n" % (mode, N, M) + "\n".join("%.9g %.9g" % (a, b) for a, b in vals)
        out = subprocess.run([exe], input=inp, capture_output=True, text=True, check=True).stdout.split()
        return np.array(out, dtype=np.float64)

    rng = np.random.default_rng(0)
    for N, M in ((1440, 120), (240, 120), (48, 12), (24, 13), (72, 9), (2880, 240)):
        x = rng.standard_normal(N).astype(np.float32)
        got = run(0, N, M, x.reshape(-1, 2)).reshape(-1, 2)
        ref = np.fft.rfft(x.astype(np.float64))[:M]
        assert np.linalg.norm(got[:, 0] + 1j * got[:, 1] - ref) / np.linalg.norm(ref) < 1e-6
        X = (rng.standard_normal(M) + 1j * rng.standard_normal(M)).astype(np.complex64)
        goti = run(1, N, M, np.stack([X.real, X.imag], 1))
        refi = np.fft.irfft(X.astype(np.complex128), n=N, norm="forward")
        assert np.linalg.norm(goti - refi) / np.linalg.norm(refi) < 1e-6


def test_fft2d_core_host_emulation(tmp_path):
    """csrc/fft2d_core.cuh + fft_reg.cuh (four-step FFT with in-register DFTs and compile-time twiddles, including the
    pruned inverse that never reads the all-zero spectrum rows) compiled for the host, against numpy.fft."""
    csrc = os.path.join(ROOT, "modulated-spherical-fourier-neural-operator_b200", "csrc")
    for f in ("fft_core.cuh", "fft_reg.cuh", "fft2d_core.cuh"):
        (tmp_path / f).write_text(open(os.path.join(csrc, f)).read().replace("#include <cuda_runtime.h>", ""))
    src = open(os.path.join(ROOT, "tests", "host_emul", "fft2d_emul.cpp")).read()
    cpp = tmp_path / "emul2d.cpp"
    cpp.write_text(src.replace('"FFT2D_CORE_HOST"', '"%s"' % (tmp_path / "fft2d_core.cuh")))
    exe = str(tmp_path / "emul2d")
    subprocess.run(["g++", "-std=c++17", "-O1", "-o", exe, str(cpp)], check=True)

    def run(mode, N, M, vals):
        inp = "%d %d %d\n" % (mode, N, M) + "\n".join("%.9g %.9g" % (a, b) for a, b in vals)
        out = subprocess.run([exe], input=inp, capture_output=True, text=True, check=True).stdout.split()
        return np.array(out, dtype=np.float64)

    rng = np.random.default_rng(0)
    for N, M in ((1440, 120), (1440, 121), (240, 120), (2880, 240), (48, 25), (72, 9)):
        x = rng.standard_normal(N).astype(np.float32)
        got = run(0, N, M, x.reshape(-1, 2)).reshape(-1, 2)
        ref = np.fft.rfft(x.astype(np.float64))[:M]
        assert np.linalg.norm(got[:, 0] + 1j * got[:, 1] - ref) / np.linalg.norm(ref) < 1e-6
        X = (rng.standard_normal(M) + 1j * rng.standard_normal(M)).astype(np.complex64)
        goti = run(1, N, M, np.stack([X.real, X.imag], 1))
        refi = np.fft.irfft(X.astype(np.complex128), n=N, norm="forward")
        assert np.linalg.norm(goti - refi) / np.linalg.norm(refi) < 1e-6


@pytest.mark.parametrize("nlon,mlim", [(1440, 121), (240, 121), (48, 13), (16, 9)])
def test_parity_split_inverse_dft_matches_irfft(nlon, mlim):
    """The algebra idft_eo_kernel (csrc/dft_tc.cu) runs on the tensor cores, restated in numpy: with the orders split by
    parity, y[j] = E[j] + O[j] and y[j + nlon/2] = E[j] - O[j] for 0 <= j < nlon/2, where E / O are products with the even /
    odd halves of the inverse DFT matrix c_m (cos, -sin)(2 pi m j / nlon) (c_0 = c_Nyquist = 1, else 2; the imaginary parts
    of the DC and Nyquist bins do not contribute).  Must equal irfft(X, n=nlon, norm="forward") of the oracle
    (torch_harmonics.InverseRealSHT) including the ignored Im(DC) / Im(Nyquist)."""
    rng = np.random.default_rng(nlon + mlim)
    rows, half = 5, nlon // 2
    X = rng.standard_normal((rows, mlim)) + 1j * rng.standard_normal((rows, mlim))
    want = torch.fft.irfft(torch.from_numpy(X), n=nlon, dim=-1, norm="forward").numpy()
    j = np.arange(half)
    E, O = np.zeros((rows, half)), np.zeros((rows, half))
    for m in range(mlim):
        cm = 1.0 if (m == 0 or 2 * m == nlon) else 2.0
        ang = 2.0 * np.pi * ((m * j) % nlon) / nlon
        im = 0.0 if (m == 0 or 2 * m == nlon) else 1.0
        term = cm * (np.outer(X[:, m].real, np.cos(ang)) - im * np.outer(X[:, m].imag, np.sin(ang)))
        if m % 2 == 0:
            E += term
        else:
            O += term
    got = np.concatenate([E + O, E - O], axis=1)
    assert np.abs(got - want).max() < 1e-12 * max(1.0, np.abs(want).max())


def test_c_abi_rejects_bad_arguments_without_a_gpu():
    """Error behaviour of the boundary (SURVEY.md section 8(b) "Errors"): every entry point validates its arguments before
    any CUDA call, returns MSFNO_ERR_BAD_SHAPE (1) and leaves a message for msfno_last_error(); the Python wrapper turns
    the code into a RuntimeError.  NULL pointers / zero sizes only -- nothing is launched."""
    import ctypes
    from msfno_b200 import _lib
    lib = _lib.lib
    bad = _lib.ERR_BAD_SHAPE if hasattr(_lib, "ERR_BAD_SHAPE") else 1
    cases = {
        "fold_affine": lambda: lib.msfno_fold_affine(None, None, None, None, None, None, 1, 1, 1, 1, 0, None),
        "fold_norm_affine": lambda: lib.msfno_fold_norm_affine(None, None, None, None, None, None, 1.0, 1e-6, 10, None, None, None,
                                                               1, 1, 1, 1, 0, None),
        "norm_film_coeffs": lambda: lib.msfno_norm_film_coeffs(None, None, None, None, None, 1.0, 1e-6, None, None, 1, 1, 10, None),
        "gelu_bwd_mul": lambda: lib.msfno_gelu_bwd_mul(None, None, None, 0, 0, None),
        "mean_carry": lambda: lib.msfno_mean_carry(None, 0, None, None, None, 0, None, None, None, None, 1, 1, None),
        "lat_segments": lambda: lib.msfno_lat_segments(1, None, None, 0, 0, 0, 0, None, None, None),
        "peer_alloc": lambda: lib.msfno_peer_alloc(0, None, None),
        "peer_open": lambda: lib.msfno_peer_open(None, None),
        "peer_block_copy": lambda: lib.msfno_peer_block_copy(None, 0, None, None),
        "peer_barrier": lambda: lib.msfno_peer_barrier(None, 0, 0, None, None),
        "plane_affine": lambda: lib.msfno_plane_affine(None, None, None, None, 0, 0, None),
    }
    for name, call in cases.items():
        assert call() == bad, name
        assert name in _lib.last_error()
    plan = ctypes.c_void_p()
    assert lib.msfno_plan_create(ctypes.byref(plan), 0, 0, 0, 0) == bad and not plan.value
    assert "plan_create" in _lib.last_error()
    with pytest.raises(RuntimeError, match="fold_affine: bad argument"):
        _lib.check(lib.msfno_fold_affine(None, None, None, None, None, None, 1, 1, 1, 1, 0, None), "fold_affine")


def test_staged_reference_copy_is_unmodified():
    """oracle/_ref/ (git-ignored, staged by oracle/build_ref.sh) must be byte-identical to the reference it was copied
    from: the recorded SHA-256 sums are checked, and against the mounted tree where there is one."""
    import hashlib
    import os
    staged = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")
    sums = os.path.join(staged, "SHA256SUMS")
    if not os.path.exists(sums):
        import pytest
        pytest.skip("oracle/_ref not staged (run oracle/build_ref.sh where /root/reference is mounted)")
    n = 0
    for line in open(sums):
        digest, rel = line.split()
        with open(os.path.join(staged, rel), "rb") as f:
            assert hashlib.sha256(f.read()).hexdigest() == digest, rel
        mounted = os.path.join("/root/reference", rel)
        if os.path.exists(mounted):
            with open(mounted, "rb") as f:
                assert hashlib.sha256(f.read()).hexdigest() == digest, "staged copy differs from " + mounted
        n += 1
    assert n >= 9


def test_launch_shares_reads_every_committed_launch_list():
    """tools/launch_shares.py (the "share of the step" evidence in DESIGN.md) must find the step boundary of every launch
    list committed under profiles/ -- whatever tier or round produced it -- and the shares must add up."""
    import glob
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "tools"))
    import launch_shares
    lists = sorted(glob.glob(os.path.join(root, "profiles", "r0*_launches_bench_*.csv")))
    assert lists, "no launch list committed under profiles/"
    for path in lists:
        res = launch_shares.shares(path)
        assert res["launches_per_step"] >= 100, path          # a 12-block forward, not one block
        assert abs(sum(v["share"] for v in res["by_kernel"].values()) - 1.0) < 1e-9, path


def test_bind_host_to_device_never_raises():
    """pipeline.bind_host_to_device: returns the CPU list it bound to or None (no GPU / no sysfs topology); never raises and
    never leaves the process with an empty affinity mask."""
    import os
    from msfno_b200.pipeline import bind_host_to_device
    before = os.sched_getaffinity(0)
    try:
        got = bind_host_to_device("cuda:0")
        assert got is None or isinstance(got, str)
        assert len(os.sched_getaffinity(0)) >= 1
    finally:
        os.sched_setaffinity(0, before)
