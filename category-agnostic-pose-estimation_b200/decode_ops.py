"""Python face of the decode-step kernels (``csrc/decode_step.cu``; C ABI in include/cape_msda.h): single-token attention
over a K/V cache, skinny linear layers with fused epilogues, tiny heads with the reference-point refinement.

Inference-only helpers used by :class:`cape_b200.transformer.AutoregressiveGenerator` (inside its CUDA graph); plain
functions over fp32 CUDA tensors, no autograd.  CUDA only, no fallback.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import _lib
from .ops import _ptr, _stream


def _p(t: Optional[torch.Tensor]):
    return None if t is None else _ptr(t)


def _rows2d(t: torch.Tensor, name: str) -> torch.Tensor:
    if t.dtype != torch.float32 or not t.is_cuda or t.dim() != 2 or t.stride(1) != 1:
        raise ValueError(f"{name} must be a CUDA fp32 (rows, cols) tensor with unit column stride, got "
                         f"{t.dtype} {tuple(t.shape)} strides {t.stride()}")
    return t


def decode_attention(q, k_cache, v_cache, k_new=None, v_new=None, pos=None, key_bias=None, n_heads: int = 8):
    """Attention of one new token per sequence (head dim 32).  q (B, C); caches (B, T, C).  Self-attention: ``k_new`` /
    ``v_new`` (B, C) and the device position ``pos`` (int64[1]) — the rows are appended at ``pos`` and 0..pos attended.
    Cross-attention: leave them None; ``key_bias`` (B, T) is added to the scores.  Returns (B, C)."""
    lib = _lib.load()
    q = _rows2d(q, "q")
    b, c = q.shape
    t = k_cache.shape[1]
    if k_cache.shape != (b, t, c) or v_cache.shape != (b, t, c) or not k_cache.is_contiguous() or not v_cache.is_contiguous():
        raise ValueError("k_cache / v_cache must be contiguous (B, T, C)")
    new_stride = 0
    if k_new is not None:
        k_new, v_new = _rows2d(k_new, "k_new"), _rows2d(v_new, "v_new")
        if k_new.stride(0) != v_new.stride(0):
            raise ValueError("k_new and v_new must share a row stride")
        new_stride = k_new.stride(0)
    out = torch.empty(b, c, dtype=torch.float32, device=q.device)
    with torch.cuda.device(q.device):
        rc = lib.cape_decode_attention(_ptr(q), q.stride(0), _p(k_new), _p(v_new), new_stride, _ptr(k_cache), _ptr(v_cache),
                                       _p(pos), _p(key_bias), _ptr(out), b, t, n_heads, c // n_heads, _stream(q.device))
    _lib.check(rc, "cape_decode_attention")
    return out


def skinny_linear(x, wt, bias=None, *, x2=None, residual=None, gamma=None, beta=None, eps: float = 1e-5, relu=False,
                  sine_dim_t=None):
    """y = epilogue(x @ wt + bias) with ``wt`` the TRANSPOSED weight (K, N).  ``gamma`` / ``beta`` select the
    LayerNorm(residual + .) epilogue, ``relu`` the ReLU one; ``x2`` is added to ``x`` first; with ``sine_dim_t`` the input
    is the sine embedding of the (rows, 2) points in ``x``."""
    lib = _lib.load()
    x = _rows2d(x, "x")
    rows = x.shape[0]
    k, n = wt.shape
    if not wt.is_contiguous() or wt.dtype != torch.float32:
        raise ValueError("wt must be a contiguous fp32 (K, N) tensor")
    if sine_dim_t is None and x.shape[1] != k:
        raise ValueError(f"x has {x.shape[1]} columns, wt expects {k}")
    epilogue = 2 if gamma is not None else (1 if relu else 0)
    if x2 is not None:
        x2 = _rows2d(x2, "x2")
    if residual is not None:
        residual = _rows2d(residual, "residual")
    y = torch.empty(rows, n, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        rc = lib.cape_skinny_linear(_ptr(x), x.stride(0), _p(x2), 0 if x2 is None else x2.stride(0), _ptr(wt), _p(bias),
                                    _p(residual), 0 if residual is None else residual.stride(0), _p(gamma), _p(beta),
                                    eps, _p(sine_dim_t), _ptr(y), n, rows, k, n, epilogue, _stream(x.device))
    _lib.check(rc, "cape_skinny_linear")
    return y


def skinny_linear_split(x, wt, bias, split: int, *, x2=None):
    """(x [+ x2]) @ wt + bias with the output columns split over two tensors: returns (y[:, :split], y[:, split:]) as two
    contiguous tensors (MSDeformAttn's offsets | logits projections in one launch)."""
    lib = _lib.load()
    x = _rows2d(x, "x")
    rows = x.shape[0]
    k, n = wt.shape
    if x2 is not None:
        x2 = _rows2d(x2, "x2")
    y = torch.empty(rows, split, dtype=torch.float32, device=x.device)
    y2 = torch.empty(rows, n - split, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        rc = lib.cape_skinny_linear_split(_ptr(x), x.stride(0), _p(x2), 0 if x2 is None else x2.stride(0), _ptr(wt),
                                          _p(bias), _ptr(y), split, _ptr(y2), n - split, split, rows, k, n,
                                          _stream(x.device))
    _lib.check(rc, "cape_skinny_linear_split")
    return y, y2


def msda_output_proj(value_cache, spatial_shapes, level_start_index, reference_points, offsets, logits, wt, bias, *,
                     residual, gamma, beta, eps: float = 1e-5):
    """MSDeformAttn sampling on the projected-value cache + output projection + residual + LayerNorm in one launch
    (``cape_msda_output_proj``): ``LayerNorm(residual + sample(...) @ wt + bias)``.  ``value_cache`` (B, S, M, 32) fp32,
    ``reference_points`` (B, Lq, L, 2), ``offsets`` (B, Lq, M, L, P, 2), ``logits`` (B, Lq, M, L * P), ``wt`` (M * 32, N)."""
    lib = _lib.load()
    b, s_len, m, d = value_cache.shape
    lq, lv, pts = offsets.shape[1], offsets.shape[3], offsets.shape[4]
    k, n = wt.shape
    if value_cache.dtype != torch.float32 or not value_cache.is_contiguous():
        raise ValueError("value_cache must be a contiguous fp32 (B, S, M, 32) tensor")
    if k != m * d or not wt.is_contiguous() or wt.dtype != torch.float32:
        raise ValueError(f"wt must be a contiguous fp32 ({m * d}, N) tensor")
    ref = reference_points.contiguous().float()
    off = offsets.contiguous().float()
    lg = logits.contiguous().float()
    residual = _rows2d(residual, "residual")
    if residual.shape[0] != b * lq:
        raise ValueError(f"residual has {residual.shape[0]} rows, expected {b * lq}")
    dims = _lib.Dims(b, s_len, m, d, lq, lv, pts)
    y = torch.empty(b * lq, n, dtype=torch.float32, device=value_cache.device)
    with torch.cuda.device(value_cache.device):
        rc = lib.cape_msda_output_proj(_ptr(value_cache), _ptr(spatial_shapes), _ptr(level_start_index), _ptr(ref), _ptr(off),
                                       _ptr(lg), ctypes.byref(dims), _ptr(wt), _p(bias), _ptr(residual), residual.stride(0),
                                       _ptr(gamma), _ptr(beta), eps, _ptr(y), n, n, _stream(value_cache.device))
    _lib.check(rc, "cape_msda_output_proj")
    return y


def coord_head_refine(x, wt, bias, w3, b3, ref, valid_ratios):
    """Last two layers of the coordinate MLP + refinement: returns (ref' (rows, 2), ref' * valid_ratios (rows, L, 2)) for
    h = relu(x @ wt + bias), ref' = sigmoid(h @ w3.T + b3 + inverse_sigmoid(ref))."""
    lib = _lib.load()
    x = _rows2d(x, "x")
    rows = x.shape[0]
    k, n = wt.shape
    levels = valid_ratios.shape[1]
    if tuple(ref.shape) != (rows, 2) or not ref.is_contiguous() or tuple(valid_ratios.shape) != (rows, levels, 2) \
            or not valid_ratios.is_contiguous() or tuple(w3.shape) != (2, n) or not w3.is_contiguous():
        raise ValueError("coord_head_refine: ref (rows, 2), valid_ratios (rows, L, 2), w3 (2, N), all contiguous")
    ref_out = torch.empty(rows, 2, dtype=torch.float32, device=x.device)
    ref_levels = torch.empty(rows, levels, 2, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        rc = lib.cape_coord_head_refine(_ptr(x), x.stride(0), _ptr(wt), _p(bias), _ptr(w3), _ptr(b3), _ptr(ref),
                                        _ptr(valid_ratios), _ptr(ref_out), _ptr(ref_levels), rows, k, n, levels,
                                        _stream(x.device))
    _lib.check(rc, "cape_coord_head_refine")
    return ref_out, ref_levels


def tiny_linear(x, w, bias=None, refine_ref=None):
    """y (rows, N <= 8) = x @ w.T + bias with ``w`` (N, K) as nn.Linear stores it; ``refine_ref`` (rows, N) turns the result
    into sigmoid(y + inverse_sigmoid(refine_ref))."""
    lib = _lib.load()
    x = _rows2d(x, "x")
    n, k = w.shape
    if refine_ref is not None and (tuple(refine_ref.shape) != (x.shape[0], n) or not refine_ref.is_contiguous()):
        raise ValueError("refine_ref must be contiguous (rows, N)")
    y = torch.empty(x.shape[0], n, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        rc = lib.cape_tiny_linear(_ptr(x), x.stride(0), _ptr(w.contiguous()), _p(bias), _p(refine_ref), _ptr(y), x.shape[0],
                                  k, n, _stream(x.device))
    _lib.check(rc, "cape_tiny_linear")
    return y
