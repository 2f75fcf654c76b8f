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

__all__ = ["msda_core", "msda_core_fwd_bwd", "msda_module_forward", "msda_decode", "query_pool_core", "points_sample"]


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


def query_pool_core(value, spatial_shapes, sampling_locations, attention_weights):
    """The sampling of TransformerDecoderLayerV4._sample_reference_points
    (/root/reference/models/deformable_transformer_v2.py:670-687): per-level grid_sample as in :115-141, then the weighted
    sum over the QUERIES (dim -2 of the stacked samples).  Returns (N, L*P, M*D)."""
    n, _, m, d = value.shape
    lq, n_levels, n_points = sampling_locations.shape[1], sampling_locations.shape[3], sampling_locations.shape[4]
    hw = _as_hw_list(spatial_shapes)
    per_level = torch.split(value, [h * w for h, w in hw], dim=1)
    grid_all = sampling_locations * 2 - 1                                                          # :673
    sampled = []
    for lvl in range(n_levels):
        h, w = hw[lvl]
        img = per_level[lvl].permute(0, 2, 3, 1).reshape(n * m, d, h, w)                           # :677
        grid = grid_all[:, :, :, lvl].permute(0, 2, 1, 3, 4).reshape(n * m, lq, n_points, 2)        # :679
        sampled.append(F.grid_sample(img, grid, mode="bilinear", padding_mode="zeros", align_corners=False))
    weights = attention_weights.permute(0, 2, 1, 3, 4).reshape(n * m, 1, lq, n_levels * n_points)   # :685
    stacked = torch.stack(sampled, dim=3).reshape(n * m, d, lq, n_levels * n_points)
    out = (stacked * weights).sum(dim=2)                                                            # :686 sum over queries
    return out.reshape(n, m * d, n_levels * n_points).permute(0, 2, 1).contiguous()


def points_sample(x, pos, n_heads, height, width):
    """MSDeformablePoints' resampling (/root/reference/models/deformable_points.py:124-128): the contiguous (B, H*W, C)
    block is viewed as (B*G, c, H, W), sampled with align_corners=True at pos given as (y, x), and laid out
    (B, Hk*Wk, G*c)."""
    b, _, c_total = x.shape
    c = c_total // n_heads
    img = x.contiguous().reshape(b * n_heads, c, height, width)
    s = F.grid_sample(img, pos[..., (1, 0)], mode="bilinear", align_corners=True)                   # (B*G, c, Hk, Wk)
    hk, wk = s.shape[2], s.shape[3]
    return s.reshape(b, n_heads, c, hk * wk).permute(0, 3, 1, 2).reshape(b, hk * wk, c_total)


def v4_sample_reference_points(w, query, src, spatial_shapes, n_heads, n_levels, n_points):
    """TransformerDecoderLayerV4._sample_reference_points (/root/reference/models/deformable_transformer_v2.py:661-687)
    with the three Linear layers given as a dict ``w`` of ``{sampling_offsets,attention_weights,source_proj}.{weight,bias}``."""
    n, lq, _ = query.shape
    hw = _as_hw_list(spatial_shapes)
    off = F.linear(query, w["sampling_offsets.weight"], w["sampling_offsets.bias"]).view(
        n, lq, n_heads, n_levels, n_points, 2)                                                      # :663
    wh = torch.tensor([[w_, h_] for h_, w_ in hw], dtype=query.dtype)                               # :664
    loc = off / wh[None, None, None, :, None, :]                                                    # :665
    attn = F.linear(query, w["attention_weights.weight"], w["attention_weights.bias"]).view(
        n, lq, n_heads, n_levels * n_points)
    attn = F.softmax(attn, 1).view(n, lq, n_heads, n_levels, n_points)                              # :667 (over queries)
    value = F.linear(src, w["source_proj.weight"], w["source_proj.bias"]).view(n, src.shape[1], n_heads, -1)   # :669
    return query_pool_core(value, hw, loc, attn)


def deformable_points_forward(w, x, spatial_shapes, n_heads, offset_range_factor):
    """MSDeformablePoints.forward (/root/reference/models/deformable_points.py:91-130) with the parameters given as a dict
    keyed like the reference's state_dict (``conv_offset.{i}.{0,1.norm,3}.*``, ``proj_q.{i}.*``)."""
    b, _, c_total = x.shape
    c = c_total // n_heads
    hw = _as_hw_list(spatial_shapes)
    n_levels = len(hw)
    out = []
    for i, cur in enumerate(x.split([h * w_ for h, w_ in hw], dim=1)):
        h, w_ = hw[i]
        ksz, stride = (n_levels - 1 - i) * 2 + 1, 2 ** (n_levels - i)                               # :54-55
        q = F.conv2d(cur.permute(0, 2, 1).reshape(b, c_total, h, w_), w[f"proj_q.{i}.weight"], w[f"proj_q.{i}.bias"])
        q_off = q.reshape(b * n_heads, c, h, w_)                                                    # :113
        t = F.conv2d(q_off, w[f"conv_offset.{i}.0.weight"], w[f"conv_offset.{i}.0.bias"], stride, ksz // 2, 1, n_heads)
        t = F.layer_norm(t.permute(0, 2, 3, 1), (c,), w[f"conv_offset.{i}.1.norm.weight"],
                         w[f"conv_offset.{i}.1.norm.bias"]).permute(0, 3, 1, 2)
        offset = F.conv2d(F.gelu(t), w[f"conv_offset.{i}.3.weight"])                                # (B*G, 2, Hk, Wk)
        hk, wk = offset.shape[2], offset.shape[3]
        if offset_range_factor >= 0:                                                                # :117-119
            rng = torch.tensor([1.0 / hk, 1.0 / wk]).reshape(1, 2, 1, 1)
            offset = offset.tanh().mul(rng).mul(offset_range_factor)
        offset = offset.permute(0, 2, 3, 1)
        ys = torch.linspace(0.5, hk - 0.5, hk, dtype=x.dtype) / hk * 2 - 1                          # :77-87
        xs = torch.linspace(0.5, wk - 0.5, wk, dtype=x.dtype) / wk * 2 - 1
        gy, gx = torch.meshgrid(ys, xs, indexing="ij")
        pos = offset + torch.stack((gy, gx), -1)[None]
        if offset_range_factor < 0:
            pos = pos.clamp(-1.0, 1.0)                                                              # :122
        out.append(points_sample(cur, pos, n_heads, h, w_))
    return torch.cat(out, dim=1)
