"""Host-side mirrors of the callers either side of the op: the deformable encoder
layer / encoder, the CAPE decoder layer (v1) and its KV cache.  Plain PyTorch host code around ``cape::ms_deform_attn``;
parameter names, shapes and forward signatures follow the reference so ``state_dict()`` is interchangeable.

    DeformableTransformerEncoderLayer   /root/reference/models/deformable_transformer.py:155-231
    DeformableTransformerEncoder        /root/reference/models/deformable_transformer.py:232-291
    TransformerDecoderLayer (v1)        /root/reference/models/deformable_transformer_v2.py:262-370
    KVCache                             /root/reference/models/kv_cache.py:3-36

Differences from the reference are host-side only: reference points of the encoder are built from a host copy of the
pyramid (one ``tolist()`` per distinct shapes tensor instead of a device sync per level per call), and the decoder
layer accepts ``input_pos`` as a Python int so that a decode step needs no device->host read (the reference evaluates
``input_pos[0] != 0`` on a CUDA tensor for every layer and step, deformable_transformer_v2.py:363).
"""
from __future__ import annotations

import copy

import torch
import torch.nn.functional as F
from torch import nn

from .gemm import linear as _linear
from .modules import MSDeformAttn, ValueCache
from .ops import masked_fill_rows_


def _activation(name: str):
    if name == "relu":
        return F.relu
    if name == "gelu":
        return F.gelu
    if name == "glu":
        return F.glu
    raise RuntimeError(f"activation should be relu/gelu, not {name}.")


def _clones(module: nn.Module, n: int) -> nn.ModuleList:
    return nn.ModuleList([copy.deepcopy(module) for _ in range(n)])


class KVCache(nn.Module):
    """Self-attention K/V store for incremental decoding (kv_cache.py:3-36): ``update`` writes the new token at
    ``input_pos`` and returns the prefix ``[:, :pos+1]``.  Buffers are non-persistent so they never enter checkpoints
    (the reference's leak into ``state_dict()``)."""

    def __init__(self, max_batch_size, max_seq_length, model_dim, dtype=torch.float32):
        super().__init__()
        shape = (max_batch_size, max_seq_length, model_dim)
        self.register_buffer("k_cache", torch.zeros(shape, dtype=dtype), persistent=False)
        self.register_buffer("v_cache", torch.zeros(shape, dtype=dtype), persistent=False)

    def update(self, input_pos, k_val, v_val):
        pos = int(input_pos[0]) if not isinstance(input_pos, int) else input_pos
        self.k_cache[:, pos:pos + k_val.shape[1]] = k_val
        self.v_cache[:, pos:pos + v_val.shape[1]] = v_val
        end = pos + k_val.shape[1]
        return self.k_cache[:, :end], self.v_cache[:, :end]


class DeformableTransformerEncoderLayer(nn.Module):
    """MSDeformAttn self-attention (query = src + pos, value = src) -> add & norm -> FFN -> add & norm
    (deformable_transformer.py:212-231)."""

    def __init__(self, d_model=256, d_ffn=1024, dropout=0.1, activation="relu", n_levels=4, n_heads=8, n_points=4):
        super().__init__()
        self.self_attn = MSDeformAttn(d_model, n_levels, n_heads, n_points)
        self.dropout1 = nn.Dropout(dropout)
        self.norm1 = nn.LayerNorm(d_model)
        self.linear1 = nn.Linear(d_model, d_ffn)
        self.activation = _activation(activation)
        self.dropout2 = nn.Dropout(dropout)
        self.linear2 = nn.Linear(d_ffn, d_model)
        self.dropout3 = nn.Dropout(dropout)
        self.norm2 = nn.LayerNorm(d_model)

    @staticmethod
    def with_pos_embed(tensor, pos):
        return tensor if pos is None else tensor + pos

    def forward_ffn(self, src):
        if self.activation is F.relu:
            hidden = self.dropout2(_linear(self.linear1, src, relu=True))     # ReLU in the GEMM epilogue when routed
        else:
            hidden = self.dropout2(self.activation(self.linear1(src)))
        return self.norm2(src + self.dropout3(_linear(self.linear2, hidden)))

    def forward(self, src, pos, reference_points, spatial_shapes, level_start_index, padding_mask=None):
        attn = self.self_attn(self.with_pos_embed(src, pos), reference_points, src, spatial_shapes, level_start_index,
                              padding_mask)
        src = self.norm1(src + self.dropout1(attn))
        return self.forward_ffn(src)


_SHAPES_CACHE: dict = {}


def _shapes_as_list(spatial_shapes):
    """Host copy of the pyramid.  Tensors are read once per (object, version): in CAPE the same shapes tensor is passed
    to every layer of a forward, so this is one device->host read per model forward instead of one per level per layer."""
    if not isinstance(spatial_shapes, torch.Tensor):
        return [(int(h), int(w)) for h, w in spatial_shapes]
    key = id(spatial_shapes)
    hit = _SHAPES_CACHE.get(key)
    if hit is not None and hit[0]() is spatial_shapes and hit[1] == spatial_shapes._version:
        return hit[2]
    import weakref
    as_list = [(int(h), int(w)) for h, w in spatial_shapes.tolist()]
    if len(_SHAPES_CACHE) > 64:
        _SHAPES_CACHE.clear()
    _SHAPES_CACHE[key] = (weakref.ref(spatial_shapes), spatial_shapes._version, as_list)
    return as_list


class DeformableTransformerEncoder(nn.Module):
    def __init__(self, encoder_layer, num_layers):
        super().__init__()
        self.layers = _clones(encoder_layer, num_layers)
        self.num_layers = num_layers

    @staticmethod
    def get_reference_points(spatial_shapes, valid_ratios, device):
        """Pixel centres of every level, normalised by the valid extent, then scaled to every level's valid ratio
        (deformable_transformer.py:248-271).  Returns (N, S, L, 2)."""
        per_level = []
        for lvl, (h, w) in enumerate(_shapes_as_list(spatial_shapes)):
            ys = torch.linspace(0.5, h - 0.5, h, dtype=torch.float32, device=device)
            xs = torch.linspace(0.5, w - 0.5, w, dtype=torch.float32, device=device)
            gy, gx = torch.meshgrid(ys, xs, indexing="ij")
            gy = gy.reshape(-1)[None] / (valid_ratios[:, None, lvl, 1] * h)
            gx = gx.reshape(-1)[None] / (valid_ratios[:, None, lvl, 0] * w)
            per_level.append(torch.stack((gx, gy), -1))
        points = torch.cat(per_level, 1)
        return points[:, :, None] * valid_ratios[:, None]

    def forward(self, src, spatial_shapes, level_start_index, valid_ratios, pos=None, padding_mask=None):
        reference_points = self.get_reference_points(spatial_shapes, valid_ratios, device=src.device)
        out = src
        for layer in self.layers:
            out = layer(out, pos, reference_points, spatial_shapes, level_start_index, padding_mask)
        return out


class TransformerDecoderLayer(nn.Module):
    """The only decoder layer CAPE can run (v1; deformable_transformer_v2.py:262-370): q/k/v projections, causal
    self-attention (with KVCache when decoding), support cross-attention over the <=100 support keypoints,
    MSDeformAttn cross-attention onto the encoder memory, FFN.  Returns ``(tgt, None)`` like the reference."""

    def __init__(self, d_model=256, d_ffn=1024, dropout=0.1, activation="relu", n_levels=4, n_heads=8, n_points=4,
                 use_qkv_proj=True):
        super().__init__()
        self.d_model = d_model
        if use_qkv_proj:
            self.attn_q = nn.Linear(d_model, d_model, bias=False)
            self.attn_k = nn.Linear(d_model, d_model, bias=False)
            self.attn_v = nn.Linear(d_model, d_model, bias=False)
        else:
            self.attn_q = self.attn_k = self.attn_v = nn.Identity()
        self.self_attn = nn.MultiheadAttention(d_model, n_heads, dropout=dropout)
        self.dropout2 = nn.Dropout(dropout)
        self.norm2 = nn.LayerNorm(d_model)
        self.support_attn = nn.MultiheadAttention(d_model, n_heads, dropout=dropout, batch_first=True)
        self.dropout_support = nn.Dropout(dropout)
        self.norm_support = nn.LayerNorm(d_model)
        self.cross_attn = MSDeformAttn(d_model, n_levels, n_heads, n_points)
        self.dropout1 = nn.Dropout(dropout)
        self.norm1 = nn.LayerNorm(d_model)
        self.linear1 = nn.Linear(d_model, d_ffn)
        self.activation = _activation(activation)
        self.dropout3 = nn.Dropout(dropout)
        self.linear2 = nn.Linear(d_ffn, d_model)
        self.dropout4 = nn.Dropout(dropout)
        self.norm3 = nn.LayerNorm(d_model)
        self.kv_cache = None

    def setup_caches(self, max_batch_size, max_seq_length, dtype=torch.float32, device=None):
        """What ``DeformableTransformer._setup_caches`` does per layer (deformable_transformer_v2.py:256-259), with a
        buffer-free value holder instead of ``VCache``."""
        self.kv_cache = KVCache(max_batch_size, max_seq_length, self.d_model, dtype).to(device)
        self.cross_attn.cache = ValueCache()

    @staticmethod
    def with_pos_embed(tensor, pos):
        return tensor if pos is None else tensor + pos[:, :tensor.size(1)]

    def forward_ffn(self, tgt):
        if self.activation is F.relu:
            hidden = self.dropout3(_linear(self.linear1, tgt, relu=True))
        else:
            hidden = self.dropout3(self.activation(self.linear1(tgt)))
        return self.norm3(tgt + self.dropout4(_linear(self.linear2, hidden)))

    def forward(self, tgt, query_pos, reference_points, src, src_spatial_shapes, level_start_index,
                src_padding_mask=None, tgt_masks=None, attn_concat_src=False, input_pos=None, support_features=None,
                support_mask=None):
        q = self.with_pos_embed(self.attn_q(tgt), query_pos)
        k = self.attn_k(tgt)
        v = self.attn_v(tgt)
        first_pos = None
        if input_pos is not None:
            first_pos = input_pos if isinstance(input_pos, int) else int(input_pos[0])
            if self.kv_cache is not None:
                k, v = self.kv_cache.update(first_pos, k, v)                                  # :325-328
        if attn_concat_src:                                                                   # :333-337
            k = torch.cat([src, k], dim=1)
            v = torch.cat([src, v], dim=1)
            tgt_masks = torch.cat([torch.zeros(q.size(1), src.size(1), device=q.device), tgt_masks],
                                  dim=1).to(dtype=torch.float32)
        attn = self.self_attn(q.transpose(0, 1), k.transpose(0, 1), v.transpose(0, 1), attn_mask=tgt_masks)[0]
        tgt = self.norm2(tgt + self.dropout2(attn.transpose(0, 1)))                           # :339-341
        if support_features is not None:                                                      # :350-357
            sup = self.support_attn(tgt, support_features, support_features, key_padding_mask=support_mask)[0]
            tgt = self.norm_support(tgt + self.dropout_support(sup))
        cross = self.cross_attn(self.with_pos_embed(tgt, query_pos), reference_points, src, src_spatial_shapes,
                                level_start_index, src_padding_mask,
                                use_cache=(first_pos is not None and first_pos != 0))         # :360-363
        tgt = self.norm1(tgt + self.dropout1(cross))
        return self.forward_ffn(tgt), None


class IncrementalDecoder:
    """Token-by-token driver for a stack of :class:`TransformerDecoderLayer` with STATIC shapes, so that the whole
    multi-layer decode step is one CUDA graph.

    What the reference does per generated token (``RoomFormerV2.forward_inference`` loop, roomformer_v2.py:481-598, through
    ``TransformerDecoderLayer.forward``, deformable_transformer_v2.py:320-370): grow the K/V prefix by slicing, run
    value_proj over all S memory tokens in every layer, and read ``input_pos[0] != 0`` back to the host.  Here, per layer:
    K/V are written into fixed ``(B, max_len, C)`` buffers at a device-resident position with ``index_copy_`` and the
    self-attention runs over the full buffer under an additive mask derived from that position; the projected value of
    the encoder memory is computed once in :meth:`reset`; the MSDeformAttn cross-attention is one fused
    ``cape::ms_deform_attn_decode`` launch.  K/V are cached already in-projected, the support keys / values are projected
    once per batch, and the small projections of a step are issued as concatenated GEMMs (call ``invalidate()`` after
    changing layer weights).  Same arithmetic as the eager layer (checked against the reference's own
    incremental outputs in tests/test_msda_gpu.py); no host synchronisation inside a step.
    """

    def __init__(self, layers, max_batch_size: int, max_len: int, device, dtype=torch.float32):
        self.layers = list(layers)
        self.max_len = max_len
        self.batch = max_batch_size
        d_model = self.layers[0].d_model
        self.device = torch.device(device)
        z = lambda *shape: torch.zeros(*shape, device=self.device, dtype=dtype)
        self.k_cache = [z(max_batch_size, max_len, d_model) for _ in self.layers]
        self.v_cache = [z(max_batch_size, max_len, d_model) for _ in self.layers]
        self.pos = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.positions = torch.arange(max_len, device=self.device)
        self.tgt = z(max_batch_size, 1, d_model)
        self.query_pos = z(max_batch_size, 1, d_model)
        self.reference_points = None
        self.values = None
        self.graph = None
        self.out = None
        self._ctx = None

    @torch.no_grad()
    def reset(self, memory, spatial_shapes, level_start_index, support_features=None, support_mask=None,
              padding_mask=None):
        """Start a new batch of sequences: project the encoder memory once per layer (the role of the reference's dead
        ``VCache``).  Everything a step reads lives in buffers owned by this object, so the captured graph survives
        across batches of the same shape (only a change of batch size, memory length or support length re-captures)."""
        n, s, _ = memory.shape
        n_levels = self.layers[0].cross_attn.n_levels
        sig = (n, s, None if support_features is None else tuple(support_features.shape), support_mask is not None)
        if sig != getattr(self, "_sig", None):
            ca0 = self.layers[0].cross_attn
            self.values = [torch.empty(n, s, ca0.n_heads, ca0.d_model // ca0.n_heads, device=self.device,
                                       dtype=memory.dtype) for _ in self.layers]
            self.reference_points = torch.zeros(n, 1, n_levels, 2, device=self.device)
            self._shapes = torch.empty(n_levels, 2, dtype=torch.int64, device=self.device)
            self._starts = torch.empty(n_levels, dtype=torch.int64, device=self.device)
            self._sup = None if support_features is None else torch.empty_like(support_features)
            self._sup_mask = None if support_mask is None else torch.empty_like(support_mask)
            self._sig = sig
            self.graph = None
        for layer, dst in zip(self.layers, self.values):
            ca = layer.cross_attn
            v = _linear(ca.value_proj, memory)
            if padding_mask is not None:
                v = masked_fill_rows_(v, padding_mask)
            dst.copy_(v.view(dst.shape))
        self._shapes.copy_(spatial_shapes)
        self._starts.copy_(level_start_index)
        if self._sup is not None:
            self._sup.copy_(support_features)
        if self._sup_mask is not None:
            self._sup_mask.copy_(support_mask)
        self._ctx = (self._shapes, self._starts, self._sup, self._sup_mask)
        # derived weights are re-derived for every batch (a generator reused across optimizer steps must not decode
        # with stale copies); they are written INTO the existing buffers so the captured graph's pointers survive
        fresh = [self._prepare_layer(layer) for layer in self.layers]
        if getattr(self, "_prep", None) is None:
            self._prep = fresh
        elif not self._refresh(self._prep, fresh):
            self.graph = None
        if self._sup is not None:   # support keys / values are constant while decoding: project them once per batch
            heads = self.layers[0].support_attn.num_heads
            if getattr(self, "_sup_k", None) is None or self._sup_k[0].shape[0] != n \
                    or self._sup_k[0].shape[2] != self._sup.shape[1]:
                shape = (n, heads, self._sup.shape[1], self._sup.shape[2] // heads)
                self._sup_k = [torch.empty(shape, device=self.device, dtype=self._sup.dtype) for _ in self.layers]
                self._sup_v = [torch.empty(shape, device=self.device, dtype=self._sup.dtype) for _ in self.layers]
                self._sup_bias = torch.zeros(n, 1, 1, self._sup.shape[1], device=self.device, dtype=self._sup.dtype)
                self.graph = None
            for prep, k_dst, v_dst in zip(self._prep, self._sup_k, self._sup_v):
                k_dst.copy_(self._split_heads(F.linear(self._sup, prep["wk_s"], prep["bk_s"]), heads))
                v_dst.copy_(self._split_heads(F.linear(self._sup, prep["wv_s"], prep["bv_s"]), heads))
            self._sup_bias.zero_()
            if self._sup_mask is not None:
                self._sup_bias.masked_fill_(self._sup_mask[:, None, None, :], float("-inf"))
        self.pos.zero_()

    def invalidate(self):
        """Forget derived weights and the captured graph.  Not needed after ordinary weight updates (``reset`` refreshes
        the derived tensors in place); use it after replacing layers or changing their shapes."""
        self._prep = None
        self.graph = None

    @classmethod
    def _refresh(cls, old, new) -> bool:
        """Bring the derived-weight structure ``old`` up to date with the freshly derived ``new`` without changing any
        tensor address a captured graph may hold: a tensor that is a live view of a parameter (same address) is left
        alone, a derived copy is overwritten in place.  Returns False when an entry had to be REPLACED (different
        shape / dtype / a re-allocated parameter), i.e. when the captured graph is no longer valid."""
        stable = True
        keys = old.keys() if isinstance(old, dict) else range(len(old))
        for k in keys:
            a, b = old[k], new[k]
            if isinstance(a, torch.Tensor) and isinstance(b, torch.Tensor):
                if a.data_ptr() == b.data_ptr() and a.shape == b.shape:
                    continue
                if a.shape == b.shape and a.dtype == b.dtype and a.device == b.device and a.is_contiguous() \
                        and not isinstance(a, nn.Parameter) and a._base is None:
                    a.copy_(b)
                else:
                    old[k] = b
                    stable = False
            elif isinstance(a, (dict, list)) and type(a) is type(b):
                stable &= cls._refresh(a, b)
            elif isinstance(a, tuple) and isinstance(b, tuple):      # (weight, bias, eps) of a LayerNorm: live parameters
                same = len(a) == len(b) and all(
                    (x.data_ptr() == y.data_ptr()) if isinstance(x, torch.Tensor) and isinstance(y, torch.Tensor) else x == y
                    for x, y in zip(a, b))
                if not same:
                    old[k] = b
                    stable = False
            elif a is not b and a != b:
                old[k] = b
                stable = False
        return stable

    def _prepare_layer(self, layer):
        """Per-layer constants of a decode step, derived once per ``reset`` from the layer's own parameters:
        concatenated / pre-multiplied projection weights so a step issues 9 small GEMMs instead of 14, with the cached
        K/V stored already in-projected (the reference re-projects the whole prefix every step inside
        ``nn.MultiheadAttention``; a Linear acts per token, so caching its output is exact)."""
        c = layer.d_model
        sa, xa, ca = layer.self_attn, layer.support_attn, layer.cross_attn
        wq_in, wk_in, wv_in = sa.in_proj_weight.chunk(3)
        bq_in, bk_in, bv_in = sa.in_proj_bias.chunk(3)
        ident = torch.eye(c, device=self.device, dtype=wq_in.dtype)
        wq = layer.attn_q.weight if isinstance(layer.attn_q, nn.Linear) else ident
        wk = layer.attn_k.weight if isinstance(layer.attn_k, nn.Linear) else ident
        wv = layer.attn_v.weight if isinstance(layer.attn_v, nn.Linear) else ident
        prep = {
            # tgt -> [attn_q(tgt) | in_proj_k(attn_k(tgt)) | in_proj_v(attn_v(tgt))]
            "w_qkv": torch.cat([wq, wk_in @ wk, wv_in @ wv], 0).contiguous(),
            "b_qkv": torch.cat([torch.zeros_like(bq_in), bk_in, bv_in], 0).contiguous(),
            "wq_in": wq_in.contiguous(), "bq_in": bq_in.contiguous(),
            # query -> [sampling_offsets | attention logits]
            "w_ol": torch.cat([ca.sampling_offsets.weight, ca.attention_weights.weight], 0).contiguous(),
            "b_ol": torch.cat([ca.sampling_offsets.bias, ca.attention_weights.bias], 0).contiguous(),
            "n_off": ca.sampling_offsets.weight.shape[0],
        }
        wq_s, wk_s, wv_s = xa.in_proj_weight.chunk(3)
        bq_s, bk_s, bv_s = xa.in_proj_bias.chunk(3)
        prep.update(wq_s=wq_s.contiguous(), bq_s=bq_s.contiguous(), wk_s=wk_s, bk_s=bk_s, wv_s=wv_s, bv_s=bv_s)
        return prep

    def _split_heads(self, x, heads):
        n, t, c = x.shape
        return x.view(n, t, heads, c // heads).transpose(1, 2)            # (n, H, t, d)

    def _layer_step(self, i, layer, tgt, query_pos, mask, reference_points=None):
        shapes, starts, sup, sup_mask = self._ctx
        if reference_points is None:
            reference_points = self.reference_points
        prep = self._prep[i]
        n, c = tgt.shape[0], layer.d_model
        heads = layer.self_attn.num_heads
        # causal self-attention over the static K/V buffers (deformable_transformer_v2.py:322-341)
        q0, k_in, v_in = F.linear(tgt, prep["w_qkv"], prep["b_qkv"]).split(c, dim=-1)
        self.k_cache[i][:n].index_copy_(1, self.pos, k_in)
        self.v_cache[i][:n].index_copy_(1, self.pos, v_in)
        q_in = F.linear(q0 + query_pos, prep["wq_in"], prep["bq_in"])
        attn = F.scaled_dot_product_attention(self._split_heads(q_in, heads),
                                              self._split_heads(self.k_cache[i][:n], heads),
                                              self._split_heads(self.v_cache[i][:n], heads), attn_mask=mask)
        attn = layer.self_attn.out_proj(attn.transpose(1, 2).reshape(n, 1, c))
        tgt = layer.norm2(tgt + attn)
        if sup is not None:                                                # support cross-attention (:350-357)
            q_s = F.linear(tgt, prep["wq_s"], prep["bq_s"])
            xs = F.scaled_dot_product_attention(self._split_heads(q_s, heads), self._sup_k[i], self._sup_v[i],
                                                attn_mask=self._sup_bias)
            xs = layer.support_attn.out_proj(xs.transpose(1, 2).reshape(n, 1, c))
            tgt = layer.norm_support(tgt + xs)
        ca = layer.cross_attn                                              # MSDeformAttn on the cached value (:360-363)
        ol = F.linear(tgt + query_pos, prep["w_ol"], prep["b_ol"])
        offsets = ol[..., :prep["n_off"]].reshape(n, 1, ca.n_heads, ca.n_levels, ca.n_points, 2)
        logits = ol[..., prep["n_off"]:].reshape(n, 1, ca.n_heads, ca.n_levels * ca.n_points)
        sampled = torch.ops.cape.ms_deform_attn_decode(self.values[i], shapes, starts, reference_points, offsets, logits)
        tgt = layer.norm1(tgt + ca.output_proj(sampled))
        return layer.forward_ffn(tgt)

    def _run(self):
        n = self.reference_points.shape[0]
        mask = torch.zeros(1, 1, 1, self.max_len, device=self.device).masked_fill_(
            (self.positions > self.pos)[None, None, None], float("-inf"))
        x = self.tgt[:n]
        for i, layer in enumerate(self.layers):
            x = self._layer_step(i, layer, x, self.query_pos[:n], mask)
        return x

    @torch.no_grad()
    def step(self, pos: int, tgt, query_pos, reference_points, use_graph: bool = True):
        """Decode token ``pos``: tgt / query_pos (B, 1, C), reference_points (B, 1, L, 2).  Returns (B, 1, C); the
        returned tensor is overwritten by the next step when ``use_graph`` is on."""
        n = tgt.shape[0]
        self.tgt[:n].copy_(tgt)
        self.query_pos[:n].copy_(query_pos)
        self.reference_points.copy_(reference_points)
        self.pos.fill_(pos)
        if not use_graph:
            return self._run()
        if self.graph is None:
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                self._run()                       # warm-up outside capture (cuBLAS handles, autotuning)
            torch.cuda.current_stream(self.device).wait_stream(side)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.out = self._run()
        self.graph.replay()
        return self.out
