"""``grid_sample`` restatement of the MSDeformAttn core.  TEST INFRASTRUCTURE ONLY.

The reference's arithmetic for this path lives in a third-party dependency, PyTorch ATen's
``grid_sampler_2d`` / ``grid_sampler_2d_backward`` (pinned only as ``torch>=2.0.0``,
``/root/reference/requirements_cape.txt:5``; this image has torch 2.11.0).  This module restates the
reference's use of it (``/root/reference/models/deformable_transformer.py:115-141``) so that the
CPU baseline runs on the same ATen kernels the reference would, with the same per-level
materialisations (transposed value copy, one ``grid_sample`` per level, stacked samples, multiply,
sum) — it is what ``bench.py``'s ``cpu_baseline`` and ``--impl reference`` legs time.

It is never imported by the product package.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

__all__ = ["msda_core", "msda_core_fwd_bwd", "msda_module_forward", "msda_decode"]


def _as_hw_list(spatial_shapes):
    if isinstance(spatial_shapes, torch.Tensor):
        spatial_shapes = spatial_shapes.tolist()
    return [(int(h), int(w)) for h, w in spatial_shapes]


def msda_core(value: torch.Tensor, spatial_shapes, sampling_locations: torch.Tensor,
              attention_weights: torch.Tensor) -> torch.Tensor:
    """Same contract as ``ms_deform_attn_core_pytorch`` (deformable_transformer.py:115-141).

    value (N,S,M,D); sampling_locations (N,Lq,M,L,P,2) in (x,y); attention_weights (N,Lq,M,L,P).
    Returns (N, Lq, M*D) contiguous.
    """
    n, _, m, d = value.shape
    lq, n_levels, n_points = sampling_locations.shape[1], sampling_locations.shape[3], sampling_locations.shape[4]
    hw = _as_hw_list(spatial_shapes)
    per_level = torch.split(value, [h * w for h, w in hw], dim=1)            # :128
    grid_all = sampling_locations * 2 - 1                                     # :129  [0,1] -> [-1,1]
    sampled = []
    for lvl in range(n_levels):                                               # :131
        h, w = hw[lvl]
        # (N, HW, M, D) -> (N*M, D, H, W): channel-first image per (sample, head)        :134
        img = per_level[lvl].permute(0, 2, 3, 1).reshape(n * m, d, h, w)
        # (N, Lq, M, P, 2) -> (N*M, Lq, P, 2)                                            :135
        grid = grid_all[:, :, :, lvl].permute(0, 2, 1, 3, 4).reshape(n * m, lq, n_points, 2)
        sampled.append(F.grid_sample(img, grid, mode="bilinear", padding_mode="zeros",
                                     align_corners=False))                   # :136-137 -> (N*M, D, Lq, P)
    weights = attention_weights.permute(0, 2, 1, 3, 4).reshape(n * m, 1, lq, n_levels * n_points)   # :139
    stacked = torch.stack(sampled, dim=3).reshape(n * m, d, lq, n_levels * n_points)               # :140
    out = (stacked * weights).sum(dim=3)                                      # :140
    return out.reshape(n, m * d, lq).permute(0, 2, 1).contiguous()            # :140-141


def msda_core_fwd_bwd(value, spatial_shapes, sampling_locations, attention_weights, grad_output):
    """Forward + autograd backward of :func:`msda_core`.  Returns (out, gvalue, gloc, gattn)."""
    v = value.detach().clone().requires_grad_(True)
    loc = sampling_locations.detach().clone().requires_grad_(True)
    a = attention_weights.detach().clone().requires_grad_(True)
    out = msda_core(v, spatial_shapes, loc, a)
    gv, gl, ga = torch.autograd.grad(out, (v, loc, a), grad_output)
    return out.detach(), gv, gl, ga


def msda_module_forward(weights: dict, query, reference_points, input_flatten, spatial_shapes,
                        padding_mask=None, n_heads=8, n_levels=4, n_points=4):
    """``MSDeformAttn.forward`` (deformable_transformer.py:76-114) from a plain weight dict.

    ``weights`` uses the reference's parameter names
    (``{sampling_offsets,attention_weights,value_proj,output_proj}.{weight,bias}``).
    """
    n, lq, c = query.shape
    s = input_flatten.shape[1]
    hw = _as_hw_list(spatial_shapes)
    assert sum(h * w for h, w in hw) == s                                      # :94
    value = F.linear(input_flatten, weights["value_proj.weight"], weights["value_proj.bias"])     # :95
    if padding_mask is not None:
        value = value.masked_fill(padding_mask[..., None], 0.0)               # :96-97
    value = value.view(n, s, n_heads, c // n_heads)                           # :98
    off = F.linear(query, weights["sampling_offsets.weight"], weights["sampling_offsets.bias"])
    off = off.view(n, lq, n_heads, n_levels, n_points, 2)                     # :99
    logits = F.linear(query, weights["attention_weights.weight"], weights["attention_weights.bias"])
    attn = F.softmax(logits.view(n, lq, n_heads, n_levels * n_points), -1)
    attn = attn.view(n, lq, n_heads, n_levels, n_points)                      # :100-101
    if reference_points.shape[-1] == 2:                                       # :102-105
        wh = torch.tensor([[w, h] for h, w in hw], dtype=query.dtype, device=query.device)
        loc = reference_points[:, :, None, :, None, :] + off / wh[None, None, None, :, None, :]
    elif reference_points.shape[-1] == 4:                                     # :106-108
        loc = reference_points[:, :, None, :, None, :2] \
            + off / n_points * reference_points[:, :, None, :, None, 2:] * 0.5
    else:
        raise ValueError("Last dim of reference_points must be 2 or 4")       # :109-111
    out = msda_core(value, hw, loc, attn)                                     # :112
    return F.linear(out, weights["output_proj.weight"], weights["output_proj.bias"])   # :113


def msda_decode(value_cache, spatial_shapes, reference_points, sampling_offsets, attention_logits):
    """Decode variant on a cached projected value (see oracle.msda_numpy.msda_decode)."""
    b, k, m, n_levels, n_points, _ = sampling_offsets.shape
    hw = _as_hw_list(spatial_shapes)
    attn = F.softmax(attention_logits.reshape(b, k, m, n_levels * n_points), -1)
    attn = attn.view(b, k, m, n_levels, n_points)
    wh = torch.tensor([[w, h] for h, w in hw], dtype=sampling_offsets.dtype, device=sampling_offsets.device)
    loc = reference_points[:, :, None, :, None, :] + sampling_offsets / wh[None, None, None, :, None, :]
    return msda_core(value_cache, hw, loc, attn)
