"""Closed-form numpy oracle for multi-scale deformable attention.  TEST INFRASTRUCTURE ONLY.

Restates, without ``grid_sample``, what the reference computes in
``ms_deform_attn_core_pytorch`` (``/root/reference/models/deformable_transformer.py:115-141``):

* ``:129``  ``sampling_grids = 2 * sampling_locations - 1``
* ``:136``  ``F.grid_sample(..., mode='bilinear', padding_mode='zeros', align_corners=False)``
  i.e. pixel coordinate ``x = ((g + 1) * W - 1) / 2`` (= ``loc_x * W - 0.5``), four corners
  ``(y0,x0) (y0,x0+1) (y0+1,x0) (y0+1,x0+1)`` with weights ``(1-lx)(1-ly), lx(1-ly), (1-lx)ly, lx*ly``,
  every corner outside ``[0,W-1] x [0,H-1]`` contributing zero;
* ``:139-141``  weighted sum over the ``L*P`` samples, output laid out ``(N, Lq, M*D)``.

The backward is the analytic derivative of that expression (what autograd of ``grid_sample``
produces): see :func:`msda_backward`.

Layouts (all C-contiguous):
    value               (N, S, M, D)     S = sum_l H_l * W_l, level l occupies rows start_l .. start_l + H_l*W_l
    spatial_shapes      (L, 2) int64     rows are (H_l, W_l)
    level_start_index   (L,)   int64
    sampling_locations  (N, Lq, M, L, P, 2)   last dim (x, y), normalised to [0, 1], unclamped
    attention_weights   (N, Lq, M, L, P)
    output              (N, Lq, M*D)
"""
from __future__ import annotations

import numpy as np

__all__ = ["level_start_index", "msda_forward", "msda_backward", "msda_decode", "query_pool_forward", "points_sample"]


def level_start_index(spatial_shapes) -> np.ndarray:
    """Row offset of each level in the flattened value tensor.

    Follows ``/root/reference/models/deformable_transformer_v2.py:204``
    (``cat((zeros(1), shapes.prod(1).cumsum(0)[:-1]))``).
    """
    shapes = np.asarray(spatial_shapes, dtype=np.int64).reshape(-1, 2)
    sizes = shapes[:, 0] * shapes[:, 1]
    return np.concatenate([np.zeros(1, np.int64), np.cumsum(sizes)[:-1]])


def _corner_terms(loc_l, H, W, dtype):
    """Per-level corner indices, weights, validity.

    loc_l: (N, Lq, M, P, 2).  Mirrors grid_sample's unnormalise for align_corners=False
    (the reference call at deformable_transformer.py:136-137 after the ``2*loc-1`` at :129).
    Returns x0, y0 (int64), lx, ly (dtype).
    """
    g = dtype(2) * loc_l - dtype(1)                      # :129
    x = (g[..., 0] + dtype(1)) * dtype(W / 2.0) - dtype(0.5)
    y = (g[..., 1] + dtype(1)) * dtype(H / 2.0) - dtype(0.5)
    x0f = np.floor(x)
    y0f = np.floor(y)
    lx = (x - x0f).astype(dtype)
    ly = (y - y0f).astype(dtype)
    # clip before the int cast so absurd locations cannot overflow; they are invalid anyway
    x0 = np.clip(x0f, -2, W + 1).astype(np.int64)
    y0 = np.clip(y0f, -2, H + 1).astype(np.int64)
    return x0, y0, lx, ly


_CORNERS = ((0, 0), (0, 1), (1, 0), (1, 1))  # (dy, dx)


def msda_forward(value, spatial_shapes, level_start, sampling_locations, attention_weights,
                 dtype=np.float64) -> np.ndarray:
    """Forward of deformable_transformer.py:115-141 in closed form.  Returns (N, Lq, M*D)."""
    dtype = np.dtype(dtype).type
    value = np.asarray(value, dtype=dtype)
    loc = np.asarray(sampling_locations, dtype=dtype)
    attn = np.asarray(attention_weights, dtype=dtype)
    shapes = np.asarray(spatial_shapes, dtype=np.int64).reshape(-1, 2)
    starts = np.asarray(level_start, dtype=np.int64).reshape(-1)
    N, S, M, D = value.shape
    _, Lq, _, L, P, _ = loc.shape
    out = np.zeros((N, Lq, M, D), dtype=dtype)
    if N == 0 or Lq == 0:
        return out.reshape(N, Lq, M * D)
    n_idx = np.arange(N).reshape(N, 1, 1, 1)
    m_idx = np.arange(M).reshape(1, 1, M, 1)
    for l in range(L):
        H, W = int(shapes[l, 0]), int(shapes[l, 1])
        x0, y0, lx, ly = _corner_terms(loc[:, :, :, l], H, W, dtype)
        a = attn[:, :, :, l]                                     # (N, Lq, M, P)
        for dy, dx in _CORNERS:
            xi, yi = x0 + dx, y0 + dy
            valid = (xi >= 0) & (xi < W) & (yi >= 0) & (yi < H)
            wx = lx if dx else (dtype(1) - lx)
            wy = ly if dy else (dtype(1) - ly)
            w = np.where(valid, a * wx * wy, dtype(0))
            rows = starts[l] + np.clip(yi, 0, H - 1) * W + np.clip(xi, 0, W - 1)
            v = value[n_idx, rows, m_idx]                        # (N, Lq, M, P, D)
            out += (w[..., None] * v).sum(axis=3)
    return out.reshape(N, Lq, M * D)


def msda_backward(grad_output, value, spatial_shapes, level_start, sampling_locations,
                  attention_weights, dtype=np.float64):
    """Analytic backward of :func:`msda_forward`.

    With ``G = grad_output[n,q,m,:]`` and, for one sample, corner values ``v_c`` (zero when the
    corner is out of bounds), weights ``w_c = wx_c * wy_c``:

        grad_attn      = sum_c w_c <G, v_c>
        grad_loc_x     = A * W_l * sum_c (+wy_c if dx else -wy_c) <G, v_c>
        grad_loc_y     = A * H_l * sum_c (+wx_c if dy else -wx_c) <G, v_c>
        grad_value[c] += A * w_c * G                      (only for in-bounds corners)

    which is what autograd produces through ``grid_sampler_2d_backward`` for the reference
    expression (deformable_transformer.py:129-141).  Returns (grad_value, grad_loc, grad_attn).
    """
    dtype = np.dtype(dtype).type
    value = np.asarray(value, dtype=dtype)
    loc = np.asarray(sampling_locations, dtype=dtype)
    attn = np.asarray(attention_weights, dtype=dtype)
    shapes = np.asarray(spatial_shapes, dtype=np.int64).reshape(-1, 2)
    starts = np.asarray(level_start, dtype=np.int64).reshape(-1)
    N, S, M, D = value.shape
    _, Lq, _, L, P, _ = loc.shape
    G = np.asarray(grad_output, dtype=dtype).reshape(N, Lq, M, D)
    gvalue = np.zeros((N * S * M, D), dtype=dtype)
    gloc = np.zeros_like(loc)
    gattn = np.zeros_like(attn)
    if N == 0 or Lq == 0:
        return gvalue.reshape(N, S, M, D), gloc, gattn
    n_idx = np.arange(N).reshape(N, 1, 1, 1)
    m_idx = np.arange(M).reshape(1, 1, M, 1)
    for l in range(L):
        H, W = int(shapes[l, 0]), int(shapes[l, 1])
        x0, y0, lx, ly = _corner_terms(loc[:, :, :, l], H, W, dtype)
        a = attn[:, :, :, l]
        ga = np.zeros_like(a)
        gx = np.zeros_like(a)
        gy = np.zeros_like(a)
        for dy, dx in _CORNERS:
            xi, yi = x0 + dx, y0 + dy
            valid = (xi >= 0) & (xi < W) & (yi >= 0) & (yi < H)
            wx = lx if dx else (dtype(1) - lx)
            wy = ly if dy else (dtype(1) - ly)
            rows = starts[l] + np.clip(yi, 0, H - 1) * W + np.clip(xi, 0, W - 1)
            v = value[n_idx, rows, m_idx]                         # (N, Lq, M, P, D)
            dot = np.where(valid, np.einsum("nqmpd,nqmd->nqmp", v, G), dtype(0))
            ga += wx * wy * dot
            gx += (wy if dx else -wy) * dot
            gy += (wx if dy else -wx) * dot
            coef = np.where(valid, a * wx * wy, dtype(0))          # (N, Lq, M, P)
            contrib = coef[..., None] * G[:, :, :, None, :]        # (N, Lq, M, P, D)
            flat = ((n_idx * S + rows) * M + m_idx).reshape(-1)
            np.add.at(gvalue, flat, contrib.reshape(-1, D))
        gattn[:, :, :, l] = ga
        gloc[:, :, :, l, :, 0] = a * dtype(W) * gx
        gloc[:, :, :, l, :, 1] = a * dtype(H) * gy
    return gvalue.reshape(N, S, M, D), gloc, gattn


def msda_decode(value_cache, spatial_shapes, level_start, reference_points, sampling_offsets,
                attention_logits, dtype=np.float64) -> np.ndarray:
    """Incremental-decode variant: the prologue of ``MSDeformAttn.forward`` fused with the core.

    Follows /root/reference/models/deformable_transformer.py:99-105,112:
        attn = softmax(logits over L*P)                               (:100-101)
        loc  = ref[:, :, None, :, None, :] + offsets / (W_l, H_l)     (:102-105, 2-d reference points)
        out  = core(value_cache, shapes, loc, attn)                   (:112)
    ``value_cache`` is the projected value ``(B, S, M, D)`` that the reference's dead ``VCache``
    (models/kv_cache.py:37-70) was meant to hold.

    reference_points (B, k, L, 2); sampling_offsets (B, k, M, L, P, 2); attention_logits (B, k, M, L*P).
    """
    dtype = np.dtype(dtype).type
    shapes = np.asarray(spatial_shapes, dtype=np.int64).reshape(-1, 2)
    off = np.asarray(sampling_offsets, dtype=dtype)
    B, k, M, L, P, _ = off.shape
    logits = np.asarray(attention_logits, dtype=dtype).reshape(B, k, M, L * P)
    logits = logits - logits.max(axis=-1, keepdims=True)
    e = np.exp(logits)
    attn = (e / e.sum(axis=-1, keepdims=True)).reshape(B, k, M, L, P)
    ref = np.asarray(reference_points, dtype=dtype)
    normalizer = np.stack([shapes[:, 1], shapes[:, 0]], -1).astype(dtype)  # (L, 2) = (W, H)
    loc = ref[:, :, None, :, None, :] + off / normalizer[None, None, None, :, None, :]
    return msda_forward(value_cache, shapes, level_start, loc, attn, dtype=dtype)


def query_pool_forward(value, spatial_shapes, level_start, sampling_locations, attention_weights,
                       dtype=np.float64) -> np.ndarray:
    """Closed form of decoder V4's inline sampler, /root/reference/models/deformable_transformer_v2.py:670-687:
    the per-sample bilinear values of :func:`msda_forward`, but weighted and summed over the QUERIES (``:686``), one
    output row per (level, point).  Returns (N, L*P, M*D)."""
    dtype = np.dtype(dtype).type
    value = np.asarray(value, dtype=dtype)
    loc = np.asarray(sampling_locations, dtype=dtype)
    attn = np.asarray(attention_weights, dtype=dtype)
    shapes = np.asarray(spatial_shapes, dtype=np.int64).reshape(-1, 2)
    starts = np.asarray(level_start, dtype=np.int64).reshape(-1)
    N, S, M, D = value.shape
    _, Lq, _, L, P, _ = loc.shape
    out = np.zeros((N, L, P, M, D), dtype=dtype)
    n_idx = np.arange(N).reshape(N, 1, 1, 1)
    m_idx = np.arange(M).reshape(1, 1, M, 1)
    for l in range(L):
        H, W = int(shapes[l, 0]), int(shapes[l, 1])
        x0, y0, lx, ly = _corner_terms(loc[:, :, :, l], H, W, dtype)
        a = attn[:, :, :, l]                                      # (N, Lq, M, P)
        for dy, dx in _CORNERS:
            xi, yi = x0 + dx, y0 + dy
            valid = (xi >= 0) & (xi < W) & (yi >= 0) & (yi < H)
            w = np.where(valid, a * (lx if dx else dtype(1) - lx) * (ly if dy else dtype(1) - ly), dtype(0))
            rows = starts[l] + np.clip(yi, 0, H - 1) * W + np.clip(xi, 0, W - 1)
            v = value[n_idx, rows, m_idx]                         # (N, Lq, M, P, D)
            out[:, l] += np.einsum("nqmp,nqmpd->npmd", w, v)
    return out.reshape(N, L * P, M * D)


def points_sample(x, pos, n_heads, height, width, dtype=np.float64) -> np.ndarray:
    """Closed form of MSDeformablePoints' resampling, /root/reference/models/deformable_points.py:124-128:
    ``grid_sample(bilinear, zeros padding, align_corners=True)`` — pixel coordinate ``(g + 1) / 2 * (size - 1)`` — of the
    contiguous ``(B, H*W, C)`` block viewed as ``(B*G, c, H, W)`` at ``pos`` (B*G, Hk, Wk, 2) given as (y, x).
    Returns (B, Hk*Wk, C)."""
    dtype = np.dtype(dtype).type
    x = np.ascontiguousarray(np.asarray(x, dtype=dtype))
    pos = np.asarray(pos, dtype=dtype)
    B, _, C = x.shape
    c = C // n_heads
    img = x.reshape(B * n_heads, c, height, width)
    BG, Hk, Wk, _ = pos.shape
    py = (pos[..., 0] + dtype(1)) * dtype(0.5) * dtype(height - 1)
    px = (pos[..., 1] + dtype(1)) * dtype(0.5) * dtype(width - 1)
    x0f, y0f = np.floor(px), np.floor(py)
    lx, ly = px - x0f, py - y0f
    x0 = np.clip(x0f, -2, width + 1).astype(np.int64)
    y0 = np.clip(y0f, -2, height + 1).astype(np.int64)
    out = np.zeros((BG, c, Hk, Wk), dtype=dtype)
    bg = np.arange(BG).reshape(BG, 1, 1)
    for dy, dx in _CORNERS:
        xi, yi = x0 + dx, y0 + dy
        valid = (xi >= 0) & (xi < width) & (yi >= 0) & (yi < height)
        w = np.where(valid, (lx if dx else dtype(1) - lx) * (ly if dy else dtype(1) - ly), dtype(0))
        v = img[bg, :, np.clip(yi, 0, height - 1), np.clip(xi, 0, width - 1)]      # (BG, Hk, Wk, c)
        out += (w[..., None] * v).transpose(0, 3, 1, 2)
    return out.reshape(B, n_heads, c, Hk * Wk).transpose(0, 3, 1, 2).reshape(B, Hk * Wk, C)
