"""ctypes front-end for the C restatement (oracle/msda_oracle.c).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes
import os

import numpy as np

from . import build_oracle

_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = build_oracle.OUT
        if not os.path.exists(path):
            path = build_oracle.build()
        _LIB = ctypes.CDLL(path)
        _LIB.msda_oracle_threads.restype = ctypes.c_int
    return _LIB


def threads() -> int:
    return int(lib().msda_oracle_threads())


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _prep(value, spatial_shapes, level_start, loc, attn, dtype):
    dtype = np.dtype(dtype)
    assert dtype in (np.dtype(np.float32), np.dtype(np.float64))
    value = np.ascontiguousarray(value, dtype=dtype)
    loc = np.ascontiguousarray(loc, dtype=dtype)
    attn = np.ascontiguousarray(attn, dtype=dtype)
    shapes = np.ascontiguousarray(np.asarray(spatial_shapes, dtype=np.int64).reshape(-1, 2))
    starts = np.ascontiguousarray(np.asarray(level_start, dtype=np.int64).reshape(-1))
    N, S, M, D = value.shape
    _, Lq, _, L, P, _ = loc.shape
    assert attn.shape == (N, Lq, M, L, P) and shapes.shape[0] == L and starts.shape[0] == L
    dims = [ctypes.c_int(int(v)) for v in (N, S, M, D, Lq, L, P)]
    suffix = "f32" if dtype == np.dtype(np.float32) else "f64"
    return value, shapes, starts, loc, attn, dims, suffix, (N, S, M, D, Lq, L, P)


def msda_forward(value, spatial_shapes, level_start, loc, attn, dtype=np.float32):
    """C forward (deformable_transformer.py:115-141).  Returns (N, Lq, M*D)."""
    value, shapes, starts, loc, attn, dims, suffix, (N, S, M, D, Lq, L, P) = _prep(
        value, spatial_shapes, level_start, loc, attn, dtype)
    out = np.empty((N, Lq, M * D), dtype=value.dtype)
    fn = getattr(lib(), "msda_oracle_forward_" + suffix)
    fn.restype = ctypes.c_int
    rc = fn(_ptr(value), _ptr(shapes), _ptr(starts), _ptr(loc), _ptr(attn), _ptr(out), *dims)
    if rc != 0:
        raise RuntimeError(f"msda_oracle_forward_{suffix} failed: {rc}")
    return out


def msda_backward(grad_output, value, spatial_shapes, level_start, loc, attn, dtype=np.float32):
    """C backward.  Returns (grad_value, grad_loc, grad_attn)."""
    value, shapes, starts, loc, attn, dims, suffix, (N, S, M, D, Lq, L, P) = _prep(
        value, spatial_shapes, level_start, loc, attn, dtype)
    gout = np.ascontiguousarray(grad_output, dtype=value.dtype).reshape(N, Lq, M * D)
    gvalue = np.empty_like(value)
    gloc = np.empty_like(loc)
    gattn = np.empty_like(attn)
    fn = getattr(lib(), "msda_oracle_backward_" + suffix)
    fn.restype = ctypes.c_int
    rc = fn(_ptr(gout), _ptr(value), _ptr(shapes), _ptr(starts), _ptr(loc), _ptr(attn),
            _ptr(gvalue), _ptr(gloc), _ptr(gattn), *dims)
    if rc != 0:
        raise RuntimeError(f"msda_oracle_backward_{suffix} failed: {rc}")
    return gvalue, gloc, gattn
