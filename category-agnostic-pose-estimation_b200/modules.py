"""``MSDeformAttn`` — host-side mirror of the reference module (``/root/reference/models/deformable_transformer.py:39-114``).

Same constructor arguments, parameter names / shapes / initialisation (so ``state_dict()`` is interchangeable with the
reference's: 8 tensors, no extra parameters or persistent buffers), same ``forward`` signature and error behaviour.
Differences are confined to *how* the result is computed:

* the sampling core is ``cape::ms_deform_attn`` (hand-written sm_100a kernels) instead of the ``grid_sample`` loop;
* ``use_cache`` — accepted and ignored by the reference (:76) — is honoured: when the caller passes
  ``use_cache=True`` (the decoder does for every token after the first, ``deformable_transformer_v2.py:360-363``)
  the projected value of the encoder memory is reused instead of re-running ``value_proj`` over all S tokens.  It is
  held in the ``VCache`` the reference attaches as ``self.cache`` (``deformable_transformer_v2.py:259``,
  ``kv_cache.py:37-70``), or in a :class:`ValueCache` the caller attaches the same way.  The cross-attention memory is fixed during
  decoding, so this is exact;
* with 2-d reference points the softmax / location arithmetic (:100-105) is fused into the sampling kernels, forward
  and backward (``cape::ms_deform_attn_decode`` and its autograd), so ``sampling_locations`` / ``attention_weights``
  never round-trip HBM; ``module.fuse_prologue = False`` restores the reference's materialised form.
"""
from __future__ import annotations

import math
import warnings

import torch
import torch.nn.functional as F
from torch import nn

from . import functional as CF
from .gemm import linear as _linear
from .ops import masked_fill_rows_


def _is_power_of_2(n) -> bool:
    if not isinstance(n, int) or n < 0:
        raise ValueError(f"invalid input for _is_power_of_2: {n} (type: {type(n)})")
    return n != 0 and (n & (n - 1)) == 0


class ValueCache:
    """Holder for a projected value ``(B, S, M, D)`` with the ``update`` / ``get`` protocol of the reference's
    ``VCache`` (models/kv_cache.py:37-70) but without registering a buffer, so it never leaks into checkpoints
    (SURVEY.md Appendix A.2).  Attach as ``msda.cache = ValueCache()`` to enable ``use_cache``."""

    def __init__(self):
        self.v_cache = None

    def update(self, v_val):
        self.v_cache = v_val

    def get(self):
        return self.v_cache


class MSDeformAttn(nn.Module):
    def __init__(self, d_model=256, n_levels=4, n_heads=8, n_points=4):
        super().__init__()
        if d_model % n_heads != 0:                                                         # :45-46
            raise ValueError("d_model must be divisible by n_heads, but got {} and {}".format(d_model, n_heads))
        if not _is_power_of_2(d_model // n_heads):
            warnings.warn("MSDeformAttn: a per-head dimension that is a power of 2 (32 in CAPE) takes the fast kernel; "
                          "other sizes run the generic one.")
        self.im2col_step = 64          # kept for attribute compatibility (:51); unused
        self.d_model = d_model
        self.n_levels = n_levels
        self.n_heads = n_heads
        self.n_points = n_points
        self.sampling_offsets = nn.Linear(d_model, n_heads * n_levels * n_points * 2)
        self.attention_weights = nn.Linear(d_model, n_heads * n_levels * n_points)
        self.value_proj = nn.Linear(d_model, d_model)
        self.output_proj = nn.Linear(d_model, d_model)
        self._value_cache = None       # identity of the cache holder we last filled (plain attribute, not a buffer)
        self.fuse_prologue = True      # False: materialise sampling_locations / attention_weights like the reference
        self._reset_parameters()

    def _reset_parameters(self):
        """Initialisation of :61-75: zero offset weights with a per-head directional bias scaled by (point index + 1),
        zero attention-weight projection (uniform 1/(L*P) weights), Xavier value/output projections."""
        with torch.no_grad():
            self.sampling_offsets.weight.zero_()
            ang = torch.arange(self.n_heads, dtype=torch.float32) * (2.0 * math.pi / self.n_heads)
            dirs = torch.stack([ang.cos(), ang.sin()], dim=-1)
            dirs = dirs / dirs.abs().max(dim=-1, keepdim=True)[0]
            bias = dirs.view(self.n_heads, 1, 1, 2).repeat(1, self.n_levels, self.n_points, 1)
            bias = bias * torch.arange(1, self.n_points + 1, dtype=torch.float32).view(1, 1, self.n_points, 1)
            self.sampling_offsets.bias.copy_(bias.reshape(-1))
            self.attention_weights.weight.zero_()
            self.attention_weights.bias.zero_()
            nn.init.xavier_uniform_(self.value_proj.weight)
            self.value_proj.bias.zero_()
            nn.init.xavier_uniform_(self.output_proj.weight)
            self.output_proj.bias.zero_()

    # -- projected-value cache (the intended role of the reference's VCache) ------------------------------------
    # A holder is whatever the caller attached as ``self.cache`` with ``update(v)`` / ``get()`` — the reference's
    # ``_setup_caches`` attaches a ``VCache`` to every decoder cross-attention (deformable_transformer_v2.py:256-259)
    # and to nothing else, so encoder self-attention never retains its value.  A holder's content is trusted only
    # after this module itself stored into that very holder object (a fresh VCache starts as zeros).
    def _cache_store(self, value):
        holder = getattr(self, "cache", None)
        if holder is not None and hasattr(holder, "update"):
            holder.update(value)
            self._value_cache = holder          # remember which holder object we filled

    def _cache_load(self, n, s):
        holder = getattr(self, "cache", None)
        if holder is None or self._value_cache is not holder or not hasattr(holder, "get"):
            return None
        value = holder.get()
        if value is None or value.dim() != 4 or value.shape[0] != n or value.shape[1] != s:
            return None
        return value

    def forward(self, query, reference_points, input_flatten, input_spatial_shapes, input_level_start_index,
                input_padding_mask=None, use_cache=False):
        """Same contract as the reference (:76-114).

        query (N, Lq, C); reference_points (N, Lq, L, 2|4); input_flatten (N, S, C); input_spatial_shapes (L, 2);
        input_level_start_index (L,); input_padding_mask (N, S) bool, True = padding.  Returns (N, Lq, C).
        """
        N, Len_q, _ = query.shape
        N, Len_in, _ = input_flatten.shape
        if isinstance(input_spatial_shapes, torch.Tensor) and not input_spatial_shapes.is_cuda:
            # host-resident shapes: the reference's consistency assert (:94) costs nothing here
            assert int((input_spatial_shapes[:, 0] * input_spatial_shapes[:, 1]).sum()) == Len_in
        value = None
        if bool(use_cache) and not torch.is_grad_enabled():
            value = self._cache_load(N, Len_in)
        if value is None:
            value = _linear(self.value_proj, input_flatten)                                 # :95
            if input_padding_mask is not None:                                              # :96-97
                # in place on the fresh projection, decided per 256-row block on the device: the all-False mask the
                # reference always passes costs one pass over the mask bytes, no copy of `value`, no host sync
                value = masked_fill_rows_(value, input_padding_mask)
            value = value.view(N, Len_in, self.n_heads, self.d_model // self.n_heads)       # :98
            if not torch.is_grad_enabled():
                self._cache_store(value)
        sampling_offsets = _linear(self.sampling_offsets, query).view(
            N, Len_q, self.n_heads, self.n_levels, self.n_points, 2)                        # :99
        attention_logits = _linear(self.attention_weights, query).view(
            N, Len_q, self.n_heads, self.n_levels * self.n_points)                          # :100
        if reference_points.shape[-1] not in (2, 4):                                        # :109-111
            raise ValueError("Last dim of reference_points must be 2 or 4, but get {} instead.".format(
                reference_points.shape[-1]))
        if reference_points.shape[-1] == 2 and value.is_cuda and self.fuse_prologue:
            # softmax (:100-101) and ref + off / (W_l, H_l) (:102-105) run inside the sampling kernel, forward and backward
            output = CF.ms_deform_attn_fused(value, input_spatial_shapes, input_level_start_index, reference_points,
                                             sampling_offsets, attention_logits)
            return _linear(self.output_proj, output)
        attention_weights = F.softmax(attention_logits, -1).view(
            N, Len_q, self.n_heads, self.n_levels, self.n_points)                           # :101
        if reference_points.shape[-1] == 2:                                                 # :102-105
            offset_normalizer = torch.stack([input_spatial_shapes[..., 1], input_spatial_shapes[..., 0]], -1)
            sampling_locations = reference_points[:, :, None, :, None, :] \
                + sampling_offsets / offset_normalizer[None, None, None, :, None, :]
        else:                                                                               # :106-108
            sampling_locations = reference_points[:, :, None, :, None, :2] \
                + sampling_offsets / self.n_points * reference_points[:, :, None, :, None, 2:] * 0.5
        output = CF.ms_deform_attn(value, input_spatial_shapes, input_level_start_index, sampling_locations,
                                   attention_weights)                                       # :112
        return _linear(self.output_proj, output)                                                     # :113
