"""ctypes binding of libmsfno_b200.so (the C ABI declared in include/msfno_b200.h).

There is NO fallback: if the shared library is missing or a symbol cannot be resolved, importing
this module raises, and every operator of the package raises on non-CUDA tensors.
"""
import ctypes
import os
import re
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
# MSFNO_B200_LIB: another BUILD of the same library (e.g. the -DMSFNO_TRACE build the tools use), never another implementation
LIB_PATH = os.environ.get("MSFNO_B200_LIB") or os.path.join(_HERE, "libmsfno_b200.so")
HEADER_PATH = os.path.join(_ROOT, "include", "msfno_b200.h")

OK = 0
LAYOUT_STD, LAYOUT_PM, LAYOUT_CM = 0, 1, 2
Q_KPAD, Q_MLIM, Q_NPACK, Q_NTRIL, Q_LJ = 0, 1, 2, 3, 4
PREC_FP32, PREC_TF32 = 0, 1
FP32_ENGINE_TC3X, FP32_ENGINE_FFMA = 0, 1

c_void_p, c_int, c_long, c_float, c_size_t = ctypes.c_void_p, ctypes.c_int, ctypes.c_long, ctypes.c_float, ctypes.c_size_t


def build(verbose=False):
    """Compile the CUDA sources in csrc/ for sm_100a into libmsfno_b200.so (in-tree)."""
    script = os.path.join(_HERE, "csrc", "build.sh")
    res = subprocess.run(["bash", script], capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout[-4000:])
        print(res.stderr[-4000:])
    if res.returncode != 0:
        raise RuntimeError("building libmsfno_b200.so failed")
    return LIB_PATH


def declared_symbols():
    """Every function name declared in include/msfno_b200.h."""
    with open(HEADER_PATH) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(msfno_[a-z0-9_]+)\s*\(", src)))


_P = c_void_p  # device / host pointers travel as integers
_SIGS = {
    "msfno_last_error": (ctypes.c_char_p, []),
    "msfno_build_info": (ctypes.c_char_p, []),
    "msfno_launch_count": (ctypes.c_ulonglong, []),
    "msfno_set_fp32_engine": (c_int, [c_int]),
    "msfno_get_fp32_engine": (c_int, []),
    "msfno_plan_create": (c_int, [ctypes.POINTER(c_void_p), c_int, c_int, c_int, c_int]),
    "msfno_plan_destroy": (c_int, [_P]),
    "msfno_plan_query": (c_long, [_P, c_int]),
    "msfno_plan_set_table": (c_int, [_P, _P, c_int, _P]),
    "msfno_plan_set_precision": (c_int, [_P, c_int]),
    "msfno_plan_get_maps": (c_int, [_P, _P, _P]),
    "msfno_sht_ws_floats": (c_size_t, [_P, c_int, c_int]),
    "msfno_sht_fwd": (c_int, [_P, _P, _P, _P, _P, _P, c_int, c_int, _P]),
    "msfno_sht_bwd": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, _P]),
    "msfno_isht_fwd": (c_int, [_P, _P, _P, _P, c_int, c_int, _P, c_int, _P, _P]),
    "msfno_isht_bwd": (c_int, [_P, _P, _P, _P, c_int, c_int, _P]),
    "msfno_fft_stage": (c_int, [_P, c_int, c_int, _P, _P, c_int, c_int, _P]),
    "msfno_legendre_stage": (c_int, [_P, c_int, _P, _P, c_int, c_int, c_int, c_int, _P]),
    "msfno_lat_segments": (c_int, [c_int, _P, _P, c_long, c_int, c_int, c_int, ctypes.POINTER(c_int), ctypes.POINTER(c_int), _P]),
    "msfno_fft_stage_peer": (c_int, [_P, c_int, _P, _P, _P, c_int, _P]),
    "msfno_peer_alloc": (c_int, [c_size_t, ctypes.POINTER(c_void_p), _P]),
    "msfno_peer_free": (c_int, [_P]),
    "msfno_peer_open": (c_int, [_P, ctypes.POINTER(c_void_p)]),
    "msfno_peer_close": (c_int, [_P]),
    "msfno_peer_block_copy": (c_int, [_P, c_int, _P, _P]),
    "msfno_peer_barrier": (c_int, [_P, c_int, c_int, _P, _P]),
    "msfno_coef_relayout": (c_int, [_P, _P, c_int, _P, c_int, c_int, c_int, _P]),
    "msfno_specconv_ws_floats": (ctypes.c_size_t, [_P, c_int, c_int, c_int]),
    "msfno_specconv_fwd": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, _P]),
    "msfno_specconv_bwd_x": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, _P]),
    "msfno_specconv_bwd_w": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, _P]),
    "msfno_specattn_ws_floats": (c_size_t, [_P, c_int, c_int, c_int, c_int]),
    "msfno_specattn_fwd": (c_int, [_P, _P, ctypes.POINTER(c_void_p), c_int, _P, _P, _P, c_int, c_int, c_int, c_int, _P]),
    "msfno_specattn_bwd_scratch_floats": (c_size_t, [_P, c_int, c_int, c_int, c_int]),
    "msfno_specattn_bwd": (c_int, [_P, _P, _P, _P, _P, ctypes.POINTER(c_void_p), _P, _P, c_int, c_int, c_int, c_int, _P]),
    "msfno_film_affine_fwd": (c_int, [_P, _P, _P, c_float, _P, c_int, c_int, c_long, _P]),
    "msfno_film_affine_bwd": (c_int, [_P, _P, _P, c_float, _P, _P, _P, c_int, c_int, c_long, _P]),
    "msfno_plane_stats": (c_int, [_P, _P, c_int, c_long, _P]),
    "msfno_norm_film_coeffs": (c_int, [_P, _P, _P, _P, _P, c_float, c_float, _P, _P, c_int, c_int, c_long, _P]),
    "msfno_mean_carry": (c_int, [_P, c_long, _P, _P, _P, c_long, _P, _P, _P, _P, c_int, c_int, _P]),
    "msfno_fold_affine": (c_int, [_P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P]),
    "msfno_fold_norm_affine": (c_int, [_P, _P, _P, _P, _P, _P, c_float, c_float, c_long, _P, _P, _P, c_int, c_int, c_int, c_int,
                                       c_int, _P]),
    "msfno_gelu_bwd_mul": (c_int, [_P, _P, _P, ctypes.c_longlong, c_int, _P]),
    "msfno_plane_affine": (c_int, [_P, _P, _P, _P, c_int, c_long, _P]),
    "msfno_conv1x1_fwd": (c_int, [_P, c_long, c_int, _P, c_long, c_long, _P, c_long, c_int, _P, c_long, _P, c_long, _P, c_long, _P,
                          c_int, c_int, c_long, c_int, c_int, _P]),
    "msfno_mlp1x1_fwd": (c_int, [_P, c_long, c_int, _P, c_long, c_long, _P, c_long, c_int, _P, c_long, _P, c_long, c_int, _P, c_long,
                                 _P, c_long, _P, c_long, _P, _P, c_int, c_int, c_long, c_int, _P]),
    "msfno_weighted_sq_sums": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, _P]),
    "msfno_weighted_diff": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, _P]),
    "msfno_gemm_nt": (c_int, [_P, c_long, _P, c_long, _P, c_long, c_int, c_int, c_int, c_int, c_int, _P]),
    "msfno_gemm_ex": (c_int, [_P, c_long, c_int, _P, c_long, c_int, _P, c_long, c_int, c_int, c_int, c_int, _P, c_long, c_int, c_int, _P]),
}


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libmsfno_b200.so not found at %s -- run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU or PyTorch fallback for the MSFNO hot path)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)  # AttributeError (loud) if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def last_error():
    return lib.msfno_last_error().decode()


def check(rc, what=""):
    if rc != OK:
        raise RuntimeError("msfno_b200 %s failed (code %d): %s" % (what, rc, last_error()))


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def on_input_device(fn):
    """Decorator of the public entry points (module forwards, conv1x1 / mlp1x1): run with the first tensor argument's
    device as the CURRENT device.  The C side launches on the current device (stream handles, tensor-map encoding,
    per-device constants), and PyTorch's own operators guard the device themselves -- so a module living on cuda:1
    must work while cuda:0 is current, exactly like the reference's modules do."""
    import functools

    import torch

    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        dev = None
        for a in args:
            if torch.is_tensor(a):
                dev = a.device if a.is_cuda else None
                break
        if dev is None or dev.index is None or dev.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)

    return wrapped


# ---- derived-weight caches (padded / TF32-packed copies of parameters) ------------------------------------------------
# They are keyed on the parameter object and its autograd version counter, which in-place writes through `param.data`
# (p.data.copy_(...), EMA utilities) do NOT bump.  `invalidate_caches()` bumps a global epoch that is part of every
# key; load_state_dict / .to() / .float() on the nets call it, and callers that write through `.data` must too.
_CACHE_EPOCH = [0]
_capture_refs = None   # list while a GraphedForward is warming up / capturing: every cached tensor a launch used


def cache_epoch():
    return _CACHE_EPOCH[0]


def invalidate_caches():
    """Forget every derived copy of a parameter (padded / TF32-rounded / packed weights, re-laid Legendre tables).
    Call after writing parameters through `.data` (the version counter does not see such writes).  A CUDA graph
    captured earlier (GraphedForward) keeps replaying the OLD derived copies: re-capture it."""
    _CACHE_EPOCH[0] += 1


def note_cached(t):
    """Called by the caches with every tensor they hand to a launch: a graph being captured keeps it alive."""
    if _capture_refs is not None and t is not None:
        _capture_refs.append(t)
    return t
