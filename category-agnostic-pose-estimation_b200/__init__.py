"""B200-native multi-scale deformable attention for CAPE (drop-in for the reference's MSDeformAttn hot path).

Importable as ``cape_b200`` (thin alias package at the repo root; this directory's own name is not a Python
identifier).  Public surface — names and argument meaning follow the reference
(``/root/reference/models/deformable_transformer.py``):

    ms_deform_attn_core_pytorch(value, value_spatial_shapes, sampling_locations, attention_weights)
    ms_deform_attn(value, spatial_shapes, level_start_index, sampling_locations, attention_weights)
    ms_deform_attn_decode(value_cache, spatial_shapes, level_start_index, reference_points, offsets, logits)
    MSDeformAttnFunction.apply(...)            MSDeformAttn(d_model, n_levels, n_heads, n_points)
    patch_reference(module)                    torch.ops.cape.{ms_deform_attn, ms_deform_attn_backward, ms_deform_attn_decode}

The compute path is libcape_msda.so (hand-written sm_100a CUDA behind the C ABI of include/cape_msda.h).  There is no
CPU, eager or Triton fallback: calling an op without the built library, or with CPU tensors, raises.
"""
from . import _lib
from ._lib import CapeLibraryError, available as library_available, launch_count
from .functional import (MSDeformAttnFunction, level_start_index_from_shapes, ms_deform_attn,
                         ms_deform_attn_core_pytorch, ms_deform_attn_decode, ms_deform_attn_fused)
from .modules import MSDeformAttn, ValueCache
from .patch import patch_reference, unpatch_reference
from .layers import (DeformableTransformerEncoder, DeformableTransformerEncoderLayer, IncrementalDecoder, KVCache,
                     TransformerDecoderLayer)
from .variants import MSDeformablePoints, ms_deform_attn_query_pool, points_sample, sample_reference_points
from .sequence import TokenState, TokenizerSpec, seq_embed
from .transformer import (MLP, AutoregressiveGenerator, DeformableTransformer, TransformerDecoder, build_prediction_heads,
                          generate_eager, load_reference_checkpoint, to_cape_predictions)
from .gemm import clear_caches as clear_linear_caches, linear_mode, linear_tf32x3, set_linear_mode
from . import synthetic

__all__ = ["MSDeformAttn", "ValueCache", "MSDeformAttnFunction", "ms_deform_attn", "ms_deform_attn_core_pytorch",
           "ms_deform_attn_decode", "ms_deform_attn_fused", "level_start_index_from_shapes", "patch_reference", "unpatch_reference",
           "library_available", "launch_count", "CapeLibraryError", "synthetic", "DeformableTransformerEncoder",
           "DeformableTransformerEncoderLayer", "TransformerDecoderLayer", "KVCache", "IncrementalDecoder", "MSDeformablePoints",
           "ms_deform_attn_query_pool", "points_sample", "sample_reference_points", "seq_embed", "TokenizerSpec", "TokenState",
           "TransformerDecoder", "DeformableTransformer", "MLP", "build_prediction_heads", "AutoregressiveGenerator",
           "generate_eager", "load_reference_checkpoint", "to_cape_predictions", "set_linear_mode", "linear_mode", "linear_tf32x3", "clear_linear_caches"]
