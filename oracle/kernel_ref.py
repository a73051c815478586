"""ctypes binding of oracle/kernel_ref.c.  TEST INFRASTRUCTURE ONLY.

numpy/torch-CPU in, numpy/torch-CPU out; float32 or float64.  Builds the
library with gcc on first use (``make -C oracle``).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libkernel_ref.so")
_lib = None


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("kernel_ref.c", "kernel_ref_body.inc")]
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs
    )
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "_build/libkernel_ref.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _np(t, dtype):
    if torch.is_tensor(t):
        t = t.detach().cpu().numpy()
    return np.ascontiguousarray(t, dtype=dtype)


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _prep(value, shapes, level_start, loc):
    dt = np.float64 if _np(value, None).dtype == np.float64 else np.float32
    value = _np(value, dt)
    B, S, H, D = value.shape
    shapes = _np(shapes, np.int64)
    level_start = _np(level_start, np.int64)
    loc = _np(loc, dt)
    L = shapes.shape[0]
    Nq, P = loc.shape[1], loc.shape[4]
    suf = "_f64" if dt == np.float64 else "_f32"
    return dt, suf, value, shapes, level_start, loc, (B, S, H, D, L, Nq, P)


def _dims(d):
    return [ctypes.c_int(int(x)) for x in d]


def box_attn_forward(value, shapes, level_start, loc, attn):
    dt, suf, value, shapes, level_start, loc, d = _prep(value, shapes, level_start, loc)
    attn = _np(attn, dt)
    B, S, H, D, L, Nq, P = d
    out = np.empty((B, Nq, H * D), dt)
    getattr(_load(), "bxo_box_attn_fwd" + suf)(
        _ptr(value), _ptr(shapes), _ptr(level_start), _ptr(loc), _ptr(attn), *_dims(d), _ptr(out))
    return torch.from_numpy(out)


def box_attn_backward(value, shapes, level_start, loc, attn, grad_out):
    dt, suf, value, shapes, level_start, loc, d = _prep(value, shapes, level_start, loc)
    attn_np = _np(attn, dt)
    grad_out = _np(grad_out, dt)
    gv, gl, ga = np.empty_like(value), np.empty_like(loc), np.empty_like(attn_np)
    getattr(_load(), "bxo_box_attn_bwd" + suf)(
        _ptr(value), _ptr(shapes), _ptr(level_start), _ptr(loc), _ptr(attn_np), _ptr(grad_out),
        *_dims(d), _ptr(gv), _ptr(gl), _ptr(ga))
    return torch.from_numpy(gv), torch.from_numpy(gl), torch.from_numpy(ga)


def instance_attn_forward(value, shapes, level_start, loc, spatial_w, level_w):
    dt, suf, value, shapes, level_start, loc, d = _prep(value, shapes, level_start, loc)
    sw, lw = _np(spatial_w, dt), _np(level_w, dt)
    B, S, H, D, L, Nq, P = d
    out = np.empty((B, Nq, H * D), dt)
    mask = np.empty((B, Nq, P, H * D), dt)
    getattr(_load(), "bxo_instance_attn_fwd" + suf)(
        _ptr(value), _ptr(shapes), _ptr(level_start), _ptr(loc), _ptr(sw), _ptr(lw), *_dims(d),
        _ptr(out), _ptr(mask))
    return torch.from_numpy(out), torch.from_numpy(mask)


def instance_attn_backward(value, shapes, level_start, loc, spatial_w, level_w, grad_out, grad_mask):
    dt, suf, value, shapes, level_start, loc, d = _prep(value, shapes, level_start, loc)
    sw, lw = _np(spatial_w, dt), _np(level_w, dt)
    grad_out, grad_mask = _np(grad_out, dt), _np(grad_mask, dt)
    gv, gl, gs, gw = np.empty_like(value), np.empty_like(loc), np.empty_like(sw), np.empty_like(lw)
    getattr(_load(), "bxo_instance_attn_bwd" + suf)(
        _ptr(value), _ptr(shapes), _ptr(level_start), _ptr(loc), _ptr(sw), _ptr(lw),
        _ptr(grad_out), _ptr(grad_mask), *_dims(d), _ptr(gv), _ptr(gl), _ptr(gs), _ptr(gw))
    return tuple(torch.from_numpy(x) for x in (gv, gl, gs, gw))
