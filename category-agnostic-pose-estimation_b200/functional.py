"""Functional entry points with the reference's names and argument meaning.

``ms_deform_attn_core_pytorch`` keeps the reference's name and 4-argument signature
(``/root/reference/models/deformable_transformer.py:115``) so that rebinding the module-level name — which
``MSDeformAttn.forward`` resolves at call time (``:112``) — is the whole integration (see ``patch.py``).
``ms_deform_attn`` / ``MSDeformAttnFunction.apply`` take upstream Deformable-DETR's 5 (+``im2col_step``) arguments.
"""
from __future__ import annotations

import torch

from . import ops  # noqa: F401  (registers torch.ops.cape.*)

__all__ = ["ms_deform_attn", "ms_deform_attn_core_pytorch", "ms_deform_attn_decode", "ms_deform_attn_fused",
           "MSDeformAttnFunction",
           "level_start_index_from_shapes"]


def level_start_index_from_shapes(spatial_shapes: torch.Tensor) -> torch.Tensor:
    """``cat((0, cumsum(H*W)[:-1]))`` on the tensor's own device — no host sync
    (same formula as ``models/deformable_transformer_v2.py:204``)."""
    sizes = spatial_shapes[:, 0] * spatial_shapes[:, 1]
    return torch.cat((sizes.new_zeros((1,)), sizes.cumsum(0)[:-1]))


def _shapes_tensor(spatial_shapes, device) -> torch.Tensor:
    if isinstance(spatial_shapes, torch.Tensor):
        return spatial_shapes
    return torch.as_tensor([[int(h), int(w)] for h, w in spatial_shapes], dtype=torch.int64, device=device)


def ms_deform_attn(value, spatial_shapes, level_start_index, sampling_locations, attention_weights):
    """(value, spatial_shapes, level_start_index, sampling_locations, attention_weights) -> (N, Lq, M*D).

    Differentiable w.r.t. value, sampling_locations and attention_weights.
    """
    spatial_shapes = _shapes_tensor(spatial_shapes, value.device)
    if not isinstance(level_start_index, torch.Tensor):
        level_start_index = torch.as_tensor(list(level_start_index), dtype=torch.int64, device=value.device)
    return torch.ops.cape.ms_deform_attn(value, spatial_shapes, level_start_index, sampling_locations,
                                         attention_weights)


def ms_deform_attn_core_pytorch(value, value_spatial_shapes, sampling_locations, attention_weights):
    """Drop-in for the reference function of the same name (deformable_transformer.py:115-141)."""
    shapes = _shapes_tensor(value_spatial_shapes, value.device)
    return torch.ops.cape.ms_deform_attn(value, shapes, level_start_index_from_shapes(shapes), sampling_locations,
                                         attention_weights)


def ms_deform_attn_decode(value_cache, spatial_shapes, level_start_index, reference_points, sampling_offsets,
                          attention_logits):
    """The module-level fused op: softmax over the logits and ``ref + off/(W,H)`` (:100-105) inside the sampling kernel.
    Used for the decode step (cached projected value, one new token) and — it is differentiable w.r.t. value,
    reference_points, sampling_offsets and attention_logits — for training."""
    spatial_shapes = _shapes_tensor(spatial_shapes, value_cache.device)
    if level_start_index is None:
        level_start_index = level_start_index_from_shapes(spatial_shapes)
    return torch.ops.cape.ms_deform_attn_decode(value_cache, spatial_shapes, level_start_index, reference_points,
                                                sampling_offsets, attention_logits)


ms_deform_attn_fused = ms_deform_attn_decode   # same op, named for its use as the module's training forward


class MSDeformAttnFunction:
    """Call-compatible with upstream ``MSDeformAttnFunction.apply(value, shapes, starts, loc, attn, im2col_step)``.
    ``im2col_step`` (the reference keeps ``self.im2col_step = 64``, :51) is accepted and ignored."""

    @staticmethod
    def apply(value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights,
              im2col_step=None):
        return ms_deform_attn(value, value_spatial_shapes, value_level_start_index, sampling_locations,
                              attention_weights)
