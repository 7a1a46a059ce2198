"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of oracle/msda_oracle.c.

Tensors in / tensors out (CPU, float32 or float64), same layouts as the reference op
(models/ops/functions/ms_deform_attn_func.py:24-42).
"""
import ctypes
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libmsda_oracle.so")
_lib = None


def build(force=False):
    """Compile the C oracle with gcc (seconds)."""
    if force or not os.path.exists(_LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        i64 = ctypes.c_int64
        vp = ctypes.c_void_p
        for sfx in ("f32", "f64"):
            f = getattr(_lib, "msda_oracle_forward_" + sfx)
            f.argtypes = [vp] * 5 + [i64] * 7 + [vp]
            f.restype = None
            b = getattr(_lib, "msda_oracle_backward_" + sfx)
            b.argtypes = [vp] * 6 + [i64] * 7 + [vp] * 3
            b.restype = None
    return _lib


def _prep(value, shapes, lsi, loc, attn):
    assert value.dtype in (torch.float32, torch.float64)
    dt = value.dtype
    value = value.detach().cpu().contiguous()
    loc = loc.detach().cpu().to(dt).contiguous()
    attn = attn.detach().cpu().to(dt).contiguous()
    shapes = shapes.detach().cpu().to(torch.int64).contiguous()
    lsi = lsi.detach().cpu().to(torch.int64).contiguous()
    N, S, M, D = value.shape
    _, Lq, _, L, P, _ = loc.shape
    sfx = "f32" if dt == torch.float32 else "f64"
    return value, shapes, lsi, loc, attn, (N, S, M, D, L, Lq, P), sfx


def forward(value, shapes, lsi, loc, attn):
    """-> (N, Lq, M*D) on CPU."""
    lib = _load()
    value, shapes, lsi, loc, attn, dims, sfx = _prep(value, shapes, lsi, loc, attn)
    N, S, M, D, L, Lq, P = dims
    out = torch.empty(N, Lq, M * D, dtype=value.dtype)
    getattr(lib, "msda_oracle_forward_" + sfx)(
        value.data_ptr(), shapes.data_ptr(), lsi.data_ptr(), loc.data_ptr(), attn.data_ptr(),
        N, S, M, D, L, Lq, P, out.data_ptr())
    return out


def backward(value, shapes, lsi, loc, attn, grad_out):
    """-> (grad_value, grad_loc, grad_attn) on CPU."""
    lib = _load()
    value, shapes, lsi, loc, attn, dims, sfx = _prep(value, shapes, lsi, loc, attn)
    N, S, M, D, L, Lq, P = dims
    grad_out = grad_out.detach().cpu().to(value.dtype).contiguous()
    gv = torch.empty_like(value)
    gl = torch.empty_like(loc)
    ga = torch.empty_like(attn)
    getattr(lib, "msda_oracle_backward_" + sfx)(
        value.data_ptr(), shapes.data_ptr(), lsi.data_ptr(), loc.data_ptr(), attn.data_ptr(),
        grad_out.data_ptr(), N, S, M, D, L, Lq, P, gv.data_ptr(), gl.data_ptr(), ga.data_ptr())
    return gv, gl, ga
