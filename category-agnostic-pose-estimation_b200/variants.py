"""Variant samplers of the reference that share the kernel family.

* ``ms_deform_attn_query_pool`` / ``sample_reference_points`` — decoder V4's inline sampler
  (``/root/reference/models/deformable_transformer_v2.py:661-687``): the sampling of ``ms_deform_attn_core_pytorch`` with the
  attention weights soft-maxed over the QUERIES and the weighted sum taken over the queries, one output row per
  (level, point).
* ``points_sample`` / ``MSDeformablePoints`` — ``/root/reference/models/deformable_points.py``: per level, offsets predicted
  by a small conv stack on a coarse grid, then ``grid_sample(align_corners=True)`` of the level (viewed channel-first the way
  the reference's ``reshape`` at ``:125`` views it) at the clamped positions.

Both are non-default paths (decoder ``v4`` / ``v41`` with ``dec_attn_concat_src``; Appendix A.5 of the survey) with small
problem sizes; the kernels are plain fp32 ones (``csrc/msda_variants.cu``).  CUDA only, no fallback.
"""
from __future__ import annotations

import ctypes
from typing import Tuple

import torch
import torch.nn.functional as F
from torch import nn

from . import _lib
from .ops import _check_shapes, _dims, _meta, _ptr, _stream


# ---- query-pooled sampling ------------------------------------------------------------------------------------------
@torch.library.custom_op("cape::ms_deform_attn_query_pool", mutates_args=(), device_types="cuda")
def ms_deform_attn_query_pool(value: torch.Tensor, spatial_shapes: torch.Tensor, level_start_index: torch.Tensor,
                              sampling_locations: torch.Tensor, attention_weights: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    _check_shapes(value, spatial_shapes, level_start_index, sampling_locations, attention_weights)
    value, loc, attn = value.float().contiguous(), sampling_locations.float().contiguous(), \
        attention_weights.float().contiguous()
    shapes, starts = _meta(spatial_shapes, value.device), _meta(level_start_index, value.device)
    dims = _dims(value, loc)
    out = torch.empty((dims.N, dims.L * dims.P, dims.M * dims.D), dtype=torch.float32, device=value.device)
    with torch.cuda.device(value.device):
        rc = lib.cape_msda_query_pool_forward(_ptr(value), _ptr(shapes), _ptr(starts), _ptr(loc), _ptr(attn), _ptr(out),
                                              ctypes.byref(dims), _stream(value.device))
    _lib.check(rc, "cape_msda_query_pool_forward")
    return out


@ms_deform_attn_query_pool.register_fake
def _(value, spatial_shapes, level_start_index, sampling_locations, attention_weights):
    n, _, m, d = value.shape
    return value.new_empty((n, sampling_locations.shape[3] * sampling_locations.shape[4], m * d), dtype=torch.float32)


@torch.library.custom_op("cape::ms_deform_attn_query_pool_backward", mutates_args=(), device_types="cuda")
def ms_deform_attn_query_pool_backward(grad_out: torch.Tensor, value: torch.Tensor, spatial_shapes: torch.Tensor,
                                       level_start_index: torch.Tensor, sampling_locations: torch.Tensor,
                                       attention_weights: torch.Tensor
                                       ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    lib = _lib.load()
    v, loc, attn = value.float().contiguous(), sampling_locations.float().contiguous(), \
        attention_weights.float().contiguous()
    shapes, starts = _meta(spatial_shapes, v.device), _meta(level_start_index, v.device)
    dims = _dims(v, loc)
    g = grad_out.float().contiguous()
    gv, gl, ga = torch.empty_like(v), torch.empty_like(loc), torch.empty_like(attn)
    with torch.cuda.device(v.device):
        rc = lib.cape_msda_query_pool_backward(_ptr(g), _ptr(v), _ptr(shapes), _ptr(starts), _ptr(loc), _ptr(attn),
                                               _ptr(gv), _ptr(gl), _ptr(ga), ctypes.byref(dims), 1, _stream(v.device))
    _lib.check(rc, "cape_msda_query_pool_backward")
    return gv.to(value.dtype), gl.to(sampling_locations.dtype), ga.to(attention_weights.dtype)


@ms_deform_attn_query_pool_backward.register_fake
def _(grad_out, value, spatial_shapes, level_start_index, sampling_locations, attention_weights):
    return torch.empty_like(value), torch.empty_like(sampling_locations), torch.empty_like(attention_weights)


def _qp_setup(ctx, inputs, output):
    ctx.save_for_backward(*inputs)


def _qp_backward(ctx, grad_out):
    value, shapes, starts, loc, attn = ctx.saved_tensors
    gv, gl, ga = torch.ops.cape.ms_deform_attn_query_pool_backward(grad_out, value, shapes, starts, loc, attn)
    return gv, None, None, gl, ga


ms_deform_attn_query_pool.register_autograd(_qp_backward, setup_context=_qp_setup)


def sample_reference_points(query, src, src_spatial_shapes, level_start_index, sampling_offsets: nn.Linear,
                            attention_weights: nn.Linear, source_proj: nn.Linear, n_heads: int, n_levels: int,
                            n_points: int):
    """``TransformerDecoderLayerV4._sample_reference_points`` (deformable_transformer_v2.py:661-687) with the layer's three
    Linear modules passed in.  Returns (N, L*P, C)."""
    n, lq, _ = query.shape
    off = sampling_offsets(query).view(n, lq, n_heads, n_levels, n_points, 2)                       # :663
    normalizer = torch.stack([src_spatial_shapes[..., 1], src_spatial_shapes[..., 0]], -1)          # :664
    loc = off / normalizer[None, None, None, :, None, :]                                            # :665 (no reference point)
    attn = F.softmax(attention_weights(query).view(n, lq, n_heads, n_levels * n_points), 1)        # :666-667 over queries
    attn = attn.view(n, lq, n_heads, n_levels, n_points)
    value = source_proj(src).view(src.size(0), src.size(1), n_heads, -1)                           # :669
    return torch.ops.cape.ms_deform_attn_query_pool(value, src_spatial_shapes, level_start_index, loc.contiguous(), attn)


# ---- planar point sampling ------------------------------------------------------------------------------------------
@torch.library.custom_op("cape::points_sample", mutates_args=(), device_types="cuda")
def points_sample(x: torch.Tensor, pos: torch.Tensor, n_heads: int, height: int, width: int) -> torch.Tensor:
    """x (B, H*W, C) contiguous — read as (B*G, c, H, W) like the reference's reshape; pos (B*G, Hk, Wk, 2) as (y, x)."""
    lib = _lib.load()
    b, hw, c_total = x.shape
    if hw != height * width or c_total % n_heads != 0 or pos.dim() != 4 or pos.shape[0] != b * n_heads \
            or pos.shape[-1] != 2:
        raise ValueError(f"points_sample: x {tuple(x.shape)}, pos {tuple(pos.shape)}, G={n_heads}, H={height}, W={width}")
    x, pos = x.float().contiguous(), pos.float().contiguous()
    hk, wk = pos.shape[1], pos.shape[2]
    out = torch.empty((b, hk * wk, c_total), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        rc = lib.cape_points_sample_forward(_ptr(x), _ptr(pos), _ptr(out), b, n_heads, c_total // n_heads, height, width,
                                            hk, wk, _stream(x.device))
    _lib.check(rc, "cape_points_sample_forward")
    return out


@points_sample.register_fake
def _(x, pos, n_heads, height, width):
    return x.new_empty((x.shape[0], pos.shape[1] * pos.shape[2], x.shape[2]), dtype=torch.float32)


@torch.library.custom_op("cape::points_sample_backward", mutates_args=(), device_types="cuda")
def points_sample_backward(grad_out: torch.Tensor, x: torch.Tensor, pos: torch.Tensor, n_heads: int, height: int,
                           width: int) -> Tuple[torch.Tensor, torch.Tensor]:
    lib = _lib.load()
    xf, pf, g = x.float().contiguous(), pos.float().contiguous(), grad_out.float().contiguous()
    b, _, c_total = xf.shape
    gx, gp = torch.empty_like(xf), torch.empty_like(pf)
    with torch.cuda.device(xf.device):
        rc = lib.cape_points_sample_backward(_ptr(g), _ptr(xf), _ptr(pf), _ptr(gx), _ptr(gp), b, n_heads,
                                             c_total // n_heads, height, width, pf.shape[1], pf.shape[2], 1,
                                             _stream(xf.device))
    _lib.check(rc, "cape_points_sample_backward")
    return gx.to(x.dtype), gp.to(pos.dtype)


@points_sample_backward.register_fake
def _(grad_out, x, pos, n_heads, height, width):
    return torch.empty_like(x), torch.empty_like(pos)


def _ps_setup(ctx, inputs, output):
    x, pos, n_heads, height, width = inputs
    ctx.save_for_backward(x, pos)
    ctx.meta = (n_heads, height, width)


def _ps_backward(ctx, grad_out):
    x, pos = ctx.saved_tensors
    gx, gp = torch.ops.cape.points_sample_backward(grad_out, x, pos, *ctx.meta)
    return gx, gp, None, None, None


points_sample.register_autograd(_ps_backward, setup_context=_ps_setup)


class LayerNormProxy(nn.Module):
    """LayerNorm over the channel dimension of a (b, c, h, w) tensor (deformable_points.py:5-30)."""

    def __init__(self, dim):
        super().__init__()
        self.norm = nn.LayerNorm(dim)

    def forward(self, x):
        return self.norm(x.permute(0, 2, 3, 1)).permute(0, 3, 1, 2)


class MSDeformablePoints(nn.Module):
    """Mirror of ``MSDeformablePoints`` (deformable_points.py:31-130), same parameter names.  The conv stack stays in
    PyTorch (cuDNN); the resampling is ``cape::points_sample``."""

    def __init__(self, embed_dim, n_levels, n_heads, offset_range_factor=-1):
        super().__init__()
        self.n_head_channels = embed_dim // n_heads
        self.scale = self.n_head_channels ** -0.5
        self.n_heads = n_heads
        self.nc = self.n_head_channels * n_heads
        self.offset_range_factor = offset_range_factor
        self.kernel_sizes = [(n_levels - 1 - i) * 2 + 1 for i in range(n_levels)]
        self.strides = [2 ** (n_levels - i) for i in range(n_levels)]
        c = self.n_head_channels
        self.conv_offset = nn.ModuleList([nn.Sequential(
            nn.Conv2d(c, c, self.kernel_sizes[i], self.strides[i], self.kernel_sizes[i] // 2, groups=self.n_heads),
            LayerNormProxy(c), nn.GELU(), nn.Conv2d(c, 2, 1, 1, 0, bias=False)) for i in range(n_levels)])
        self.proj_q = nn.ModuleList([nn.Conv2d(self.nc, self.nc, kernel_size=1, stride=1, padding=0)
                                     for _ in range(n_levels)])

    @torch.no_grad()
    def _get_ref_points(self, h_key, w_key, b, dtype, device):
        ys = torch.linspace(0.5, h_key - 0.5, h_key, dtype=dtype, device=device) / h_key * 2.0 - 1.0
        xs = torch.linspace(0.5, w_key - 0.5, w_key, dtype=dtype, device=device) / w_key * 2.0 - 1.0
        gy, gx = torch.meshgrid(ys, xs, indexing="ij")
        return torch.stack((gy, gx), -1)[None].expand(b * self.n_heads, -1, -1, -1)                # (y, x), :77-87

    def forward(self, x, spatial_shapes, level_start_index):
        from .layers import _shapes_as_list
        b = x.size(0)
        hw = _shapes_as_list(spatial_shapes)
        out = []
        for i, cur in enumerate(x.split([h * w for h, w in hw], dim=1)):
            h, w = hw[i]
            cur = cur.contiguous()
            q = self.proj_q[i](cur.permute(0, 2, 1).reshape(b, self.nc, h, w))                      # :112
            q_off = q.reshape(b * self.n_heads, self.n_head_channels, h, w)                         # :113
            offset = self.conv_offset[i](q_off)                                                     # :114
            hk, wk = offset.size(2), offset.size(3)
            if self.offset_range_factor >= 0:                                                       # :117-119
                rng = torch.tensor([1.0 / hk, 1.0 / wk], device=x.device).reshape(1, 2, 1, 1)
                offset = offset.tanh().mul(rng).mul(self.offset_range_factor)
            offset = offset.permute(0, 2, 3, 1)
            pos = offset + self._get_ref_points(hk, wk, b, x.dtype, x.device)
            if self.offset_range_factor < 0:
                pos = pos.clamp(-1.0, 1.0)                                                          # :122
            out.append(torch.ops.cape.points_sample(cur, pos.contiguous(), self.n_heads, h, w))     # :124-128
        return torch.cat(out, dim=1)
