"""Callers of the op one level up (§8f rank 2): the token decoder, the deformable
transformer that owns encoder + decoder, and the autoregressive generation loop — host-side mirrors with the
reference's parameter names, so a reference ``transformer.state_dict()`` loads unchanged.

    TransformerDecoder        /root/reference/models/deformable_transformer_v2.py:951-1131
    DeformableTransformer     /root/reference/models/deformable_transformer_v2.py:56-259
    MLP, prediction heads     /root/reference/models/roomformer_v2.py:178-249, 956-968
    generation loop           /root/reference/models/roomformer_v2.py:381-676 (RoomFormerV2.forward_inference, decoder part)

Only decoder layer ``v1`` exists here: it is the only one CAPE can run (the other variants do not accept the support
keyword arguments the decoder passes).  ``inject_cls_embed`` (a floor-plan feature) is not
mirrored.

What differs from the reference is confined to where the work happens:

* ``_seq_embed`` is one kernel (``cape::seq_embed``) instead of four embedding lookups and eleven element-wise ops;
* :class:`AutoregressiveGenerator` keeps the whole loop on the device.  The reference round-trips every generated token
  through numpy and Python lists (``.item()`` per sample per step, :548-597), rebuilds nine input tensors from lists
  each step (:483-491), re-projects the encoder memory in every layer for every token and reads ``input_pos[0] != 0``
  back on the host per layer.  Here one step — token embedding, all decoder layers on static K/V buffers with the fused
  MSDeformAttn decode kernel, reference-point refinement, class / coordinate heads, token bookkeeping — is ONE CUDA
  graph replayed per token, and the host only polls the "unfinished" flags every few steps.
"""
from __future__ import annotations

import math
from typing import Optional

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from .layers import (DeformableTransformerEncoder, DeformableTransformerEncoderLayer, IncrementalDecoder,
                     TransformerDecoderLayer, _clones)
from .modules import MSDeformAttn
from .sequence import TokenState, TokenizerSpec


def inverse_sigmoid(x, eps=1e-5):
    """util/misc.py:436-440."""
    x = x.clamp(min=0, max=1)
    return torch.log(x.clamp(min=eps) / (1 - x).clamp(min=eps))


def sincos_position_table(embed_dim: int, seq_len: int) -> np.ndarray:
    """(seq_len, embed_dim) 1-D sin/cos table (deformable_transformer_v2.py:33-53): [sin | cos] halves, float64 omega."""
    assert embed_dim % 2 == 0
    omega = 1.0 / 10000 ** (np.arange(embed_dim // 2, dtype=np.float64) / (embed_dim / 2.0))
    out = np.einsum("m,d->md", np.arange(seq_len, dtype=np.float32).reshape(-1), omega)
    return np.concatenate([np.sin(out), np.cos(out)], axis=1)


class MLP(nn.Module):
    """roomformer_v2.py:956-968."""

    def __init__(self, input_dim, hidden_dim, output_dim, num_layers):
        super().__init__()
        self.num_layers = num_layers
        h = [hidden_dim] * (num_layers - 1)
        self.layers = nn.ModuleList(nn.Linear(n, k) for n, k in zip([input_dim] + h, h + [output_dim]))

    def forward(self, x):
        for i, layer in enumerate(self.layers):
            x = F.relu(layer(x)) if i < self.num_layers - 1 else layer(x)
        return x


def build_prediction_heads(d_model: int, num_classes: int, num_pred: int, with_poly_refine: bool = True):
    """The class / coordinate heads exactly as ``RoomFormerV2.__init__`` makes and initialises them
    (roomformer_v2.py:178-179, 219-237).  Returns (class_embed, coords_embed) ModuleLists of length ``num_pred``."""
    class_embed = nn.Linear(d_model, num_classes)
    coords_embed = MLP(d_model, d_model, 2, 3)
    prior_prob = 0.01
    class_embed.bias.data = torch.ones(num_classes) * (-math.log((1 - prior_prob) / prior_prob))
    nn.init.constant_(coords_embed.layers[-1].weight.data, 0)
    nn.init.constant_(coords_embed.layers[-1].bias.data, 0)
    if with_poly_refine:
        return _clones(class_embed, num_pred), _clones(coords_embed, num_pred)
    return (nn.ModuleList([class_embed for _ in range(num_pred)]),
            nn.ModuleList([coords_embed for _ in range(num_pred)]))


class TransformerDecoder(nn.Module):
    def __init__(self, decoder_layer, num_layers, poly_refine=True, return_intermediate=False, aux_loss=False,
                 query_pos_type="none", vocab_size=None, pad_idx=None, use_anchor=None):
        super().__init__()
        self.layers = _clones(decoder_layer, num_layers)
        self.num_layers = num_layers
        self.poly_refine = poly_refine
        self.return_intermediate = return_intermediate
        self.aux_loss = aux_loss
        self.query_pos_type = query_pos_type
        self.use_anchor = use_anchor
        self.coords_embed = None          # attached by the owner of the heads (roomformer_v2.py:245-246)
        self.class_embed = None
        self.pos_trans = None
        self.pos_trans_norm = None
        d_model = self.layers[0].d_model
        self.token_embed = nn.Embedding(vocab_size, d_model, padding_idx=pad_idx)               # :23-30
        nn.init.normal_(self.token_embed.weight, mean=0, std=d_model ** -0.5)
        if pad_idx is not None:
            nn.init.constant_(self.token_embed.weight[pad_idx], 0)

    def _seq_embed(self, seq11, seq12, seq21, seq22, delta_x1, delta_x2, delta_y1, delta_y2):
        pad = self.token_embed.padding_idx
        return torch.ops.cape.seq_embed(self.token_embed.weight, seq11, seq12, seq21, seq22, delta_x1, delta_x2,
                                        delta_y1, delta_y2, -1 if pad is None else int(pad))

    @staticmethod
    def get_query_pos_embed(ref_points):
        """Sine embedding of (N, T, 2) reference points -> (N, T, 256) (:1005-1018)."""
        num_pos_feats, temperature = 128, 10000
        dim_t = torch.arange(num_pos_feats, dtype=torch.float32, device=ref_points.device)
        dim_t = temperature ** (2 * (dim_t // 2) / num_pos_feats)
        pos = (ref_points * (2 * math.pi))[:, :, :, None] / dim_t
        return torch.stack((pos[:, :, :, 0::2].sin(), pos[:, :, :, 1::2].cos()), dim=4).flatten(2)

    @staticmethod
    def with_pos_embed(tensor, pos):
        return tensor if pos is None else tensor + pos[:, :tensor.size(1)]

    def query_pos_from(self, reference_points):
        return self.pos_trans_norm(self.pos_trans(self.get_query_pos_embed(reference_points)))

    def refine(self, lid, output, reference_points):
        """Iterative refinement of the reference points by layer ``lid``'s coordinate head (:1096-1102)."""
        return (self.coords_embed[lid](output) + inverse_sigmoid(reference_points)).sigmoid()

    def forward(self, tgt, reference_points, src, src_flatten, src_spatial_shapes, src_level_start_index,
                src_valid_ratios, query_pos=None, src_padding_mask=None, tgt_masks=None, seq_kwargs=None,
                force_simple_returns=False, pre_decoder_pos_embed=False, attn_concat_src=False, decode_token_pos=None,
                support_features=None, support_mask=None):
        if support_features is None:                                                            # :1036-1039
            support_features = getattr(self, "support_features", None)
        if support_mask is None:
            support_mask = getattr(self, "support_mask", None)
        output = self._seq_embed(*[seq_kwargs[k] for k in ("seq11", "seq12", "seq21", "seq22", "delta_x1", "delta_x2",
                                                            "delta_y1", "delta_y2")])
        if decode_token_pos is not None:                                                        # :1046-1050
            pos = decode_token_pos if isinstance(decode_token_pos, int) else int(decode_token_pos.reshape(-1)[0])
            if query_pos is not None:
                query_pos = query_pos[:, pos:pos + 1]
            if reference_points is not None:
                reference_points = reference_points[:, pos:pos + 1]
            decode_token_pos = pos
        if reference_points is None:
            reference_points = torch.zeros(output.shape[0], output.shape[1], 2, device=output.device)
        refining = self.poly_refine or self.use_anchor
        if pre_decoder_pos_embed:                                                               # :1056-1061
            if refining and self.query_pos_type == "sine":
                query_pos = self.query_pos_from(reference_points)
            output = self.with_pos_embed(output, query_pos)
            query_pos = None
        intermediate, intermediate_refs, intermediate_classes = [], [], []
        point_classes = torch.zeros(output.shape[0], output.shape[1], self.class_embed[0].out_features,
                                    device=output.device)
        last = len(self.layers) - 1
        for lid, layer in enumerate(self.layers):
            if refining:
                assert reference_points.shape[-1] == 2
                reference_points_input = reference_points[:, :, None] * src_valid_ratios[:, None]
                if not pre_decoder_pos_embed:
                    if self.query_pos_type == "sine":
                        query_pos = self.query_pos_from(reference_points)
                    elif self.query_pos_type == "none":
                        query_pos = None
            else:
                reference_points_input = None
            output, src_tmp = layer(output, query_pos, reference_points_input, src, src_spatial_shapes,
                                    src_level_start_index, src_padding_mask, tgt_masks,
                                    attn_concat_src=attn_concat_src, input_pos=decode_token_pos,
                                    support_features=support_features, support_mask=support_mask)
            if src_tmp is not None:
                src = src_tmp
            if self.poly_refine:
                reference_points = self.refine(lid, output, reference_points)
            elif lid == last:
                if self.use_anchor:
                    reference_points = self.refine(-1, output, reference_points)
                else:
                    reference_points = self.coords_embed[-1](output).sigmoid()
            if self.aux_loss:
                point_classes = self.class_embed[lid](output)
            elif lid == last:
                point_classes = self.class_embed[-1](output)
            if self.return_intermediate:
                intermediate.append(output)
                intermediate_refs.append(reference_points)
                intermediate_classes.append(point_classes)
        if self.return_intermediate and not force_simple_returns:
            return torch.stack(intermediate), torch.stack(intermediate_refs), torch.stack(intermediate_classes)
        return output, reference_points, point_classes


class DeformableTransformer(nn.Module):
    def __init__(self, d_model=256, nhead=8, num_encoder_layers=6, num_decoder_layers=6, dim_feedforward=1024,
                 dropout=0.1, activation="relu", poly_refine=True, return_intermediate_dec=False, aux_loss=False,
                 num_feature_levels=4, dec_n_points=4, enc_n_points=4, query_pos_type="none", vocab_size=None,
                 seq_len=1024, pre_decoder_pos_embed=False, learnable_dec_pe=False, dec_attn_concat_src=False,
                 dec_qkv_proj=True, dec_layer_type="v1", pad_idx=None, use_anchor=False, inject_cls_embed=False):
        super().__init__()
        if dec_layer_type != "v1":
            raise ValueError(f"dec_layer_type={dec_layer_type!r}: only decoder layer 'v1' accepts the support keyword "
                             "arguments CAPE passes, the other variants raise TypeError in the reference too")
        if inject_cls_embed:
            raise NotImplementedError("inject_cls_embed (floor-plan room classes) is outside the CAPE path")
        self.d_model = d_model
        self.nhead = nhead
        self.poly_refine = poly_refine
        self.use_anchor = use_anchor
        self.inject_cls_embed = inject_cls_embed
        self.encoder = DeformableTransformerEncoder(
            DeformableTransformerEncoderLayer(d_model, dim_feedforward, dropout, activation, num_feature_levels, nhead,
                                              enc_n_points), num_encoder_layers)
        decoder_layer = TransformerDecoderLayer(d_model, dim_feedforward, dropout, activation, num_feature_levels, nhead,
                                                dec_n_points, use_qkv_proj=(dec_qkv_proj and not dec_attn_concat_src))
        self.decoder = TransformerDecoder(decoder_layer, num_decoder_layers, poly_refine, return_intermediate_dec,
                                          aux_loss, query_pos_type, vocab_size, pad_idx, use_anchor=use_anchor)
        self.level_embed = nn.Parameter(torch.Tensor(num_feature_levels, d_model))
        if query_pos_type == "sine" and (poly_refine or use_anchor):                              # :126-128
            self.decoder.pos_trans = nn.Linear(d_model, d_model)
            self.decoder.pos_trans_norm = nn.LayerNorm(d_model)
        self.pre_decoder_pos_embed = pre_decoder_pos_embed
        self.pos_embed = nn.Parameter(torch.zeros(1, seq_len, d_model), requires_grad=learnable_dec_pe)
        self.pos_embed.data.copy_(torch.from_numpy(sincos_position_table(d_model, seq_len)).float().unsqueeze(0))
        self.dec_attn_concat_src = dec_attn_concat_src
        self._reset_parameters()

    def _reset_parameters(self):
        """:148-155 — Xavier on every matrix (this includes ``pos_embed`` and the token table), then MSDeformAttn's own
        initialisation, then N(0, 1) level embeddings."""
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)
        for m in self.modules():
            if isinstance(m, MSDeformAttn):
                m._reset_parameters()
        nn.init.normal_(self.level_embed)

    def attach_heads(self, class_embed: nn.ModuleList, coords_embed: nn.ModuleList):
        """What ``RoomFormerV2.__init__`` does at :245-246: the decoder uses the heads of the model that owns it."""
        self.decoder.class_embed = class_embed
        self.decoder.coords_embed = coords_embed
        return self

    @staticmethod
    def get_valid_ratio(mask):
        _, h, w = mask.shape
        valid_h = torch.sum(~mask[:, :, 0], 1)
        valid_w = torch.sum(~mask[:, 0, :], 1)
        return torch.stack([valid_w.float() / w, valid_h.float() / h], -1)

    @staticmethod
    def _create_causal_attention_mask(seq_len):
        return torch.triu(torch.full((seq_len, seq_len), float("-inf")), diagonal=1)

    def encode(self, srcs, masks, pos_embeds):
        """Flatten the pyramid and run the encoder (:181-217).  Returns the ``enc_cache`` dict of the reference."""
        src_flatten, mask_flatten, pos_flatten, shapes = [], [], [], []
        for lvl, (src, mask, pos_embed) in enumerate(zip(srcs, masks, pos_embeds)):
            shapes.append(tuple(src.shape[-2:]))
            src_flatten.append(src.flatten(2).transpose(1, 2))
            mask_flatten.append(mask.flatten(1))
            pos_flatten.append(pos_embed.flatten(2).transpose(1, 2) + self.level_embed[lvl].view(1, 1, -1))
        src_flatten = torch.cat(src_flatten, 1)
        mask_flatten = torch.cat(mask_flatten, 1)
        pos_flatten = torch.cat(pos_flatten, 1)
        spatial_shapes = torch.as_tensor(shapes, dtype=torch.long, device=src_flatten.device)
        level_start_index = torch.cat((spatial_shapes.new_zeros((1,)), spatial_shapes.prod(1).cumsum(0)[:-1]))
        valid_ratios = torch.stack([self.get_valid_ratio(m) for m in masks], 1)
        memory = self.encoder(src_flatten, spatial_shapes, level_start_index, valid_ratios, pos_flatten, mask_flatten)
        return {"memory": memory, "spatial_shapes": spatial_shapes, "level_start_index": level_start_index,
                "valid_ratios": valid_ratios, "mask_flatten": mask_flatten, "src_flatten": src_flatten}

    def forward(self, srcs, masks, pos_embeds, query_embed=None, tgt=None, tgt_masks=None, seq_kwargs=None,
                force_simple_returns=False, return_enc_cache=False, enc_cache=None, decode_token_pos=None,
                support_features=None, support_mask=None):
        if enc_cache is None:
            enc_cache = self.encode(srcs, masks, pos_embeds)
        memory = enc_cache["memory"]
        bs = memory.shape[0]
        assert not (self.use_anchor and self.poly_refine), "use_anchor and poly_refine cannot be used together"
        if self.poly_refine or self.use_anchor:                                                   # :226-230
            reference_points = query_embed.unsqueeze(0).expand(bs, -1, -1).sigmoid()
            query_pos = None
        else:
            reference_points = None
            query_pos = self.pos_embed
        init_reference_out = reference_points
        if tgt_masks is None:                                                                     # :236-241
            if decode_token_pos is not None:
                pos = decode_token_pos if isinstance(decode_token_pos, int) else int(decode_token_pos.max())
                tgt_masks = torch.zeros(1, pos + 1, dtype=torch.float, device=memory.device)
            else:
                tgt_masks = self._create_causal_attention_mask(seq_kwargs["seq11"].shape[1]).to(memory.device)
        hs, inter_references, inter_classes = self.decoder(
            tgt, reference_points, memory, enc_cache["src_flatten"], enc_cache["spatial_shapes"],
            enc_cache["level_start_index"], enc_cache["valid_ratios"], query_pos, enc_cache["mask_flatten"], tgt_masks,
            seq_kwargs, force_simple_returns=force_simple_returns, pre_decoder_pos_embed=self.pre_decoder_pos_embed,
            attn_concat_src=self.dec_attn_concat_src, decode_token_pos=decode_token_pos,
            support_features=support_features, support_mask=support_mask)
        if return_enc_cache:
            return hs, init_reference_out, inter_references, inter_classes, enc_cache
        return hs, init_reference_out, inter_references, inter_classes

    def _setup_caches(self, max_batch_size, max_seq_length, max_vision_length=None, model_dim=None, nhead=None,
                      dtype=torch.float32, device=None):
        """:256-259, with the projected-value holder that ``use_cache`` really serves."""
        for layer in self.decoder.layers:
            layer.setup_caches(max_batch_size, max_seq_length, dtype, device)


class AutoregressiveGenerator(IncrementalDecoder):
    """``RoomFormerV2.forward_inference`` from the encoder onwards, device-resident.

    ``generate`` takes what the reference hands its transformer (projected feature maps, masks, positional encodings,
    the reference-point embedding, support features / mask) and returns the reference's result dict: ``pred_logits``
    (B, steps, n_classes), ``pred_coords`` (B, steps, 2) and ``gen_out`` (per sample a list of ``[x, y]`` / ``2`` / ``-1``
    entries), plus ``sequences`` = argmax of the logits (what ``CAPEModel.forward_inference`` adds, cape_model.py:200-209).

    Supports the configuration CAPE trains (iterative refinement, ``query_pos_type`` 'sine' or 'none', no
    ``pre_decoder_pos_embed``, no ``dec_attn_concat_src``); other configurations are served by the mirror's own
    ``forward(..., decode_token_pos=i)`` path (:func:`generate_eager`).
    """

    def __init__(self, transformer: DeformableTransformer, spec: TokenizerSpec, max_batch_size: int, device,
                 fused: Optional[bool] = None):
        dec = transformer.decoder
        if not transformer.poly_refine or transformer.pre_decoder_pos_embed or transformer.dec_attn_concat_src \
                or dec.query_pos_type not in ("sine", "none") or dec.class_embed is None or dec.coords_embed is None:
            raise NotImplementedError("AutoregressiveGenerator covers the CAPE configuration (poly_refine, query positions "
                                      "'sine' or 'none', attached heads); use generate_eager() for the others")
        super().__init__(dec.layers, max_batch_size, spec.seq_len, device)
        self.transformer = transformer
        self.spec = spec
        self.n_classes = dec.class_embed[-1].out_features
        self.state = TokenState(max_batch_size, spec, self.n_classes, device)
        self.pos = self.state.step                      # the K/V write position IS the generation step counter
        self.ref_table = torch.zeros(spec.seq_len, 2, device=self.device)
        # last-layer hidden state of every step (what the reference collects in output_hs_list, roomformer_v2.py:519-524,
        # for heads applied after the loop such as room_class_embed, :647-654)
        self.hidden = torch.zeros(max_batch_size, spec.seq_len, transformer.d_model, device=self.device)
        self.valid_ratios = None
        layer0 = dec.layers[0]
        can_fuse = (transformer.d_model == 256 and transformer.d_model // transformer.nhead == 32
                    and isinstance(layer0.attn_q, nn.Linear) and layer0.activation is F.relu
                    and dec.query_pos_type == "sine" and spec.seq_len <= 1024)
        if fused and not can_fuse:
            raise NotImplementedError("the fused decode step needs d_model 256, head dim 32, q/k/v projections, ReLU and "
                                      "sine query positions (the CAPE configuration)")
        self.fused = can_fuse if fused is None else fused
        # MSDeformAttn sampling + output_proj + residual + norm1 as ONE launch (cape_msda_output_proj: 15 launches per layer);
        # False: the sampling op followed by the skinny linear (16 launches, identical results)
        self.fuse_output_proj = True
        self._fprep = None

    def invalidate(self):
        super().invalidate()
        self._fprep = None

    def _prepare_fused(self):
        """Transposed / concatenated weights of the hand-written decode step (csrc/decode_step.cu), derived once from the
        modules' own parameters (call ``invalidate()`` after changing them)."""
        dec = self.transformer.decoder
        t = lambda w: w.detach().t().contiguous()
        layers = []
        for lid, layer in enumerate(self.layers):
            base = self._prep[lid]
            sa, xa, ca = layer.self_attn, layer.support_attn, layer.cross_attn
            head = dec.coords_embed[lid]
            layers.append({
                "wt_qkv": t(base["w_qkv"]), "b_qkv": base["b_qkv"], "wt_qin": t(base["wq_in"]), "b_qin": base["bq_in"],
                "wt_o": t(sa.out_proj.weight), "b_o": sa.out_proj.bias, "n2": (layer.norm2.weight, layer.norm2.bias, layer.norm2.eps),
                "wt_qs": t(base["wq_s"]), "b_qs": base["bq_s"], "wt_so": t(xa.out_proj.weight), "b_so": xa.out_proj.bias,
                "ns": (layer.norm_support.weight, layer.norm_support.bias, layer.norm_support.eps),
                "wt_ol": t(base["w_ol"]), "b_ol": base["b_ol"], "n_off": base["n_off"],
                "wt_out": t(ca.output_proj.weight), "b_out": ca.output_proj.bias,
                "n1": (layer.norm1.weight, layer.norm1.bias, layer.norm1.eps),
                "wt_f1": t(layer.linear1.weight), "b_f1": layer.linear1.bias, "wt_f2": t(layer.linear2.weight),
                "b_f2": layer.linear2.bias, "n3": (layer.norm3.weight, layer.norm3.bias, layer.norm3.eps),
                "wt_c1": t(head.layers[0].weight), "b_c1": head.layers[0].bias, "wt_c2": t(head.layers[1].weight),
                "b_c2": head.layers[1].bias, "w_c3": head.layers[2].weight.detach().contiguous(), "b_c3": head.layers[2].bias,
            })
        dim_t = torch.arange(128, dtype=torch.float32, device=self.device)
        dim_t = 10000 ** (2 * (dim_t // 2) / 128)                               # get_query_pos_embed, :1010-1011
        return {"layers": layers, "dim_t": dim_t.contiguous(), "wt_pos": t(dec.pos_trans.weight), "b_pos": dec.pos_trans.bias,
                "npos": (dec.pos_trans_norm.weight, dec.pos_trans_norm.bias, dec.pos_trans_norm.eps),
                "w_cls": dec.class_embed[-1].weight.detach().contiguous(), "b_cls": dec.class_embed[-1].bias}

    def _run_fused(self):
        """One token through every decoder layer with the hand-written step kernels: 15 launches per layer, no library
        GEMM / attention call.  Same arithmetic as :meth:`_run` (fp32 FMA; only the summation order differs)."""
        from . import decode_ops as K
        dec, st = self.transformer.decoder, self.state
        n = self.batch
        fp = self._fprep
        shapes, starts, sup, _ = self._ctx
        heads = self.layers[0].self_attn.num_heads
        x = dec._seq_embed(*st.seq, *st.delta).view(n, -1)                                     # (B, C)
        ref = self.ref_table.index_select(0, self.pos).expand(n, -1).contiguous()              # (B, 2), :1049-1050
        ref_levels = (ref[:, None, :] * self.valid_ratios).contiguous()                        # (B, L, 2), :1072
        for lid, layer in enumerate(self.layers):
            w = fp["layers"][lid]
            ca = layer.cross_attn
            g, b_, eps = fp["npos"]
            qpos = K.skinny_linear(ref, fp["wt_pos"], fp["b_pos"], gamma=g, beta=b_, eps=eps, sine_dim_t=fp["dim_t"])
            # causal self-attention on the K/V cache (:322-341)
            qkv = K.skinny_linear(x, w["wt_qkv"], w["b_qkv"])                                  # [attn_q(x) | k_in | v_in]
            q_in = K.skinny_linear(qkv[:, :256], w["wt_qin"], w["b_qin"], x2=qpos)
            attn = K.decode_attention(q_in, self.k_cache[lid], self.v_cache[lid], qkv[:, 256:512], qkv[:, 512:], self.pos,
                                      n_heads=heads)
            g, b_, eps = w["n2"]
            x = K.skinny_linear(attn, w["wt_o"], w["b_o"], residual=x, gamma=g, beta=b_, eps=eps)
            if sup is not None:                                                                # support cross-attention (:350-357)
                q_s = K.skinny_linear(x, w["wt_qs"], w["b_qs"])
                xs = K.decode_attention(q_s, self._sup_k_rows[lid], self._sup_v_rows[lid], key_bias=self._sup_bias_rows,
                                        n_heads=heads)
                g, b_, eps = w["ns"]
                x = K.skinny_linear(xs, w["wt_so"], w["b_so"], residual=x, gamma=g, beta=b_, eps=eps)
            # MSDeformAttn on the cached projected value (:360-363)
            offsets, logits = K.skinny_linear_split(x, w["wt_ol"], w["b_ol"], w["n_off"], x2=qpos)   # :99-100
            g, b_, eps = w["n1"]
            off6 = offsets.view(n, 1, ca.n_heads, ca.n_levels, ca.n_points, 2)
            lg4 = logits.view(n, 1, ca.n_heads, ca.n_levels * ca.n_points)
            if self.fuse_output_proj and self.values[lid].dtype == torch.float32 and ca.n_levels == 4 and ca.n_points == 4:
                # sampling + output_proj + residual + norm1 in one launch: the sampled rows never leave the SM (:112-113)
                x = K.msda_output_proj(self.values[lid], shapes, starts, ref_levels.view(n, 1, ca.n_levels, 2), off6, lg4,
                                       w["wt_out"], w["b_out"], residual=x, gamma=g, beta=b_, eps=eps)
            else:
                sampled = torch.ops.cape.ms_deform_attn_decode(
                    self.values[lid], shapes, starts, ref_levels.view(n, 1, ca.n_levels, 2), off6, lg4).view(n, -1)
                x = K.skinny_linear(sampled, w["wt_out"], w["b_out"], residual=x, gamma=g, beta=b_, eps=eps)
            hidden = K.skinny_linear(x, w["wt_f1"], w["b_f1"], relu=True)                      # FFN (:366-368)
            g, b_, eps = w["n3"]
            x = K.skinny_linear(hidden, w["wt_f2"], w["b_f2"], residual=x, gamma=g, beta=b_, eps=eps)
            # iterative refinement of the reference point by this layer's coordinate head (:1096-1102)
            c = K.skinny_linear(x, w["wt_c1"], w["b_c1"], relu=True)
            ref, ref_levels = K.coord_head_refine(c, w["wt_c2"], w["b_c2"], w["w_c3"], w["b_c3"], ref, self.valid_ratios)
        cls = K.tiny_linear(x, fp["w_cls"], fp["b_cls"])                                       # :1117-1121
        self.hidden.index_copy_(1, self.pos, x.view(n, 1, -1))
        st.advance(cls, ref)
        return x

    def _run(self):
        if self.fused:
            return self._run_fused()
        dec, st = self.transformer.decoder, self.state
        n = self.batch
        mask = torch.zeros(1, 1, 1, self.max_len, device=self.device).masked_fill_(
            (self.positions > self.pos)[None, None, None], float("-inf"))
        x = dec._seq_embed(*st.seq, *st.delta)                                               # (B, 1, C)
        ref = self.ref_table.index_select(0, self.pos).unsqueeze(0).expand(n, -1, -1)        # (B, 1, 2), :1049-1050
        zero_pos = torch.zeros_like(x)
        for lid, layer in enumerate(self.layers):
            ref_input = ref[:, :, None] * self.valid_ratios[:, None]                         # (B, 1, L, 2), :1072
            qpos = dec.query_pos_from(ref) if dec.query_pos_type == "sine" else zero_pos     # :1075-1079
            x = self._layer_step(lid, layer, x, qpos, mask, ref_input.contiguous())
            ref = dec.refine(lid, x, ref)
        cls = dec.class_embed[-1](x)                                                         # :1117-1121
        self.hidden.index_copy_(1, self.pos, x.view(n, 1, -1))
        st.advance(cls.contiguous(), ref.contiguous())
        return x

    @torch.no_grad()
    def generate(self, srcs, masks, pos_embeds, query_embed, support_features=None, support_mask=None,
                 use_graph: bool = True, poll_every: int = 16, enc_cache: Optional[dict] = None):
        tr = self.transformer
        if enc_cache is None:
            enc_cache = tr.encode(srcs, masks, pos_embeds)
        memory = enc_cache["memory"]
        n = memory.shape[0]
        if n != self.batch:
            raise ValueError(f"generator was built for batch {self.batch}, got {n}")
        if support_features is None:
            support_features = getattr(tr.decoder, "support_features", None)
        if support_mask is None:
            support_mask = getattr(tr.decoder, "support_mask", None)
        sig_before = getattr(self, "_sig", None)
        self.reset(memory, enc_cache["spatial_shapes"], enc_cache["level_start_index"], support_features, support_mask,
                   padding_mask=enc_cache["mask_flatten"])
        if self.valid_ratios is None or self.valid_ratios.shape != enc_cache["valid_ratios"].shape \
                or sig_before != self._sig:
            self.valid_ratios = torch.empty_like(enc_cache["valid_ratios"])
            self.graph = None
        self.valid_ratios.copy_(enc_cache["valid_ratios"])
        self.ref_table.copy_(query_embed[:self.spec.seq_len].sigmoid())                      # :227 (the loop stops at tokenizer.seq_len)
        if self.fused:
            fresh = self._prepare_fused()                 # re-derived per batch, written into the captured buffers
            if self._fprep is None:
                self._fprep = fresh
                self.graph = None
            elif not self._refresh(self._fprep, fresh):
                self.graph = None
            if self._sup is not None:   # support K / V as (B, T, C) rows for the attention kernel
                to_rows = lambda t: t.transpose(1, 2).reshape(t.shape[0], t.shape[2], -1)
                rows_shape = to_rows(self._sup_k[0]).shape
                if getattr(self, "_sup_k_rows", None) is None or self._sup_k_rows[0].shape != rows_shape:
                    # buffers the captured graph reads: allocated once per shape, refilled per batch
                    self._sup_k_rows = [torch.empty(rows_shape, device=self.device) for _ in self.layers]
                    self._sup_v_rows = [torch.empty(rows_shape, device=self.device) for _ in self.layers]
                    self._sup_bias_rows = torch.empty(rows_shape[0], rows_shape[1], device=self.device)
                    self.graph = None
                for dst, src in zip(self._sup_k_rows + self._sup_v_rows, self._sup_k + self._sup_v):
                    dst.copy_(to_rows(src))
                self._sup_bias_rows.copy_(self._sup_bias.view(rows_shape[0], rows_shape[1]))
        self.state.reset()
        max_len = self.spec.seq_len
        if use_graph and self.graph is None:
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                self._run()                       # warm-up outside capture (cuBLAS handles, autotuning)
            torch.cuda.current_stream(self.device).wait_stream(side)
            for buf in self.k_cache + self.v_cache:
                buf.zero_()
            self.state.reset()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self._run()
            for buf in self.k_cache + self.v_cache:
                buf.zero_()
            self.state.reset()                    # capture itself launches nothing, but leave a clean slate regardless
        done = 0
        while done < max_len:                     # `while i < max_len and unfinish_flag.any()` (:481), polled in chunks
            burst = min(poll_every, max_len - done)
            for _ in range(burst):
                if use_graph:
                    self.graph.replay()
                else:
                    self._run()
            done += burst
            if done < max_len and not bool(self.state.unfinished.any()):
                break
        return self._collect()

    def _collect(self):
        st = self.state
        finish = st.finish_step.cpu()
        # the reference stops after the step in which the last sample emitted <eos>; steps run past that in the last
        # polling burst only appended <pad> bookkeeping and are cut off here
        steps = self.spec.seq_len if bool((finish < 0).any()) else int(finish.max()) + 1
        steps = min(steps, int(st.step.item()))
        logits = st.pred_logits[:, :steps].clone()
        coords = st.pred_coords[:, :steps].clone()
        kind = st.gen_kind[:, :steps].cpu().tolist()                 # one D2H copy each, then plain Python lists
        xy = st.gen_xy[:, :steps].cpu().tolist()
        gen_out = [[p if k == 0 else k for k, p in zip(kinds, points)] for kinds, points in zip(kind, xy)]
        return {"pred_logits": logits, "pred_coords": coords, "gen_out": gen_out, "sequences": logits.argmax(-1),
                "steps": steps, "hidden": self.hidden[:, :steps].clone()}


@torch.no_grad()
def generate_eager(transformer: DeformableTransformer, spec: TokenizerSpec, srcs, masks, pos_embeds, query_embed,
                   support_features=None, support_mask=None):
    """The reference's loop shape (one ``transformer.forward(..., decode_token_pos=i)`` per token with the KV caches,
    roomformer_v2.py:481-535) with the token bookkeeping on the device; works for every configuration the mirror
    supports.  One host read of the unfinished flags per step, as in the reference."""
    device = srcs[0].device
    n = srcs[0].shape[0]
    dec = transformer.decoder
    transformer._setup_caches(n, spec.seq_len, dtype=srcs[0].dtype, device=device)
    state = TokenState(n, spec, dec.class_embed[-1].out_features, device)
    keys = ("seq11", "seq12", "seq21", "seq22", "delta_x1", "delta_x2", "delta_y1", "delta_y2")
    enc_cache = None
    steps = 0
    try:
        while steps < spec.seq_len and bool(state.unfinished.any()):
            seq_kwargs = dict(zip(keys, state.seq + state.delta))
            _, _, reg, cls, enc_cache = transformer(srcs, masks, pos_embeds, query_embed, None, None, seq_kwargs,
                                                    force_simple_returns=True, return_enc_cache=True, enc_cache=enc_cache,
                                                    decode_token_pos=steps, support_features=support_features,
                                                    support_mask=support_mask)
            state.advance(cls.contiguous(), reg.contiguous())
            steps += 1
    finally:
        for layer in dec.layers:
            layer.kv_cache = None
            layer.cross_attn.cache = None
    return {"pred_logits": state.pred_logits[:, :steps].clone(), "pred_coords": state.pred_coords[:, :steps].clone(),
            "sequences": state.pred_logits[:, :steps].argmax(-1), "steps": steps}


def to_cape_predictions(generated: dict) -> dict:
    """The dict ``CAPEModel.forward_inference`` returns (cape_model.py:200-209) from :meth:`AutoregressiveGenerator.generate`
    / :func:`generate_eager` output: ``sequences`` = argmax token types, ``coordinates``, ``logits``."""
    logits = generated["pred_logits"]
    return {"sequences": logits.argmax(dim=-1), "coordinates": generated["pred_coords"], "logits": logits}


def load_reference_checkpoint(transformer: DeformableTransformer, model_state: dict, prefix: Optional[str] = None):
    """Load the transformer part of a reference checkpoint (``checkpoint['model']`` of train_cape_episodic.py:863-888, i.e. a
    ``CAPEModel.state_dict()``) into the mirror.

    Reference keys look like ``base_model.transformer.encoder.layers.0...`` (CAPEModel wraps RoomFormerV2 as
    ``base_model``, cape_model.py:40-60); the prediction heads the decoder uses live at ``base_model.class_embed.*`` /
    ``base_model.coords_embed.*`` as well as under ``...transformer.decoder.*`` (same modules, roomformer_v2.py:245-246).
    Cache buffers that the reference leaks into its state dict (``kv_cache.*``, ``cross_attn.cache.*``; SURVEY.md
    Appendix A.2) are dropped.  Returns ``(query_embed_weight or None, missing_keys, unexpected_keys)``."""
    if prefix is None:
        for cand in ("base_model.transformer.", "transformer.", ""):
            if any(k.startswith(cand + "encoder.layers.") for k in model_state):
                prefix = cand
                break
        else:
            raise KeyError("no '...encoder.layers.*' keys found: not a RoomFormerV2 / CAPEModel state dict")
    state = {}
    for k, v in model_state.items():
        if not k.startswith(prefix):
            continue
        name = k[len(prefix):]
        if ".kv_cache." in name or ".cross_attn.cache." in name:
            continue
        state[name] = v
    outer = prefix[:-len("transformer.")] if prefix.endswith("transformer.") else None
    if outer is not None:   # heads saved on the owning model only (older checkpoints)
        for head in ("class_embed", "coords_embed"):
            for k, v in model_state.items():
                if k.startswith(outer + head + ".") and ("decoder." + k[len(outer):]) not in state:
                    state["decoder." + k[len(outer):]] = v
    missing, unexpected = transformer.load_state_dict(state, strict=False)
    query = model_state.get((outer or "") + "query_embed.weight")
    return query, list(missing), list(unexpected)


# ---- binding the mirror to a live reference model (patch_reference(..., swap_forward_inference=True)) -------------
_ACTIVATION_NAMES = {F.relu: "relu", F.gelu: "gelu", F.glu: "glu"}


def _reference_transformer_config(ref) -> dict:
    """Constructor arguments of the reference ``DeformableTransformer`` (deformable_transformer_v2.py:56-62) read back
    from a live instance.  Raises NotImplementedError for the configurations the mirror does not cover."""
    dec, enc = ref.decoder, ref.encoder
    layer0 = dec.layers[0]
    if type(layer0).__name__ != "TransformerDecoderLayer":
        raise NotImplementedError(f"decoder layer {type(layer0).__name__}: only 'v1' is mirrored (the only one CAPE can run)")
    if getattr(ref, "inject_cls_embed", False) or getattr(dec, "room_class_embed", None) is not None:
        raise NotImplementedError("inject_cls_embed / room classes are outside the CAPE path")
    ca = layer0.cross_attn
    return dict(
        d_model=ref.d_model, nhead=ref.nhead, num_encoder_layers=enc.num_layers, num_decoder_layers=dec.num_layers,
        dim_feedforward=layer0.linear1.out_features, dropout=float(layer0.dropout1.p),
        activation=_ACTIVATION_NAMES[layer0.activation], poly_refine=bool(ref.poly_refine),
        return_intermediate_dec=bool(dec.return_intermediate), aux_loss=bool(dec.aux_loss),
        num_feature_levels=int(ref.level_embed.shape[0]), dec_n_points=int(ca.n_points),
        enc_n_points=int(enc.layers[0].self_attn.n_points), query_pos_type=dec.query_pos_type,
        vocab_size=int(dec.token_embed.num_embeddings), seq_len=int(ref.pos_embed.shape[1]),
        pre_decoder_pos_embed=bool(ref.pre_decoder_pos_embed), learnable_dec_pe=bool(ref.pos_embed.requires_grad),
        dec_attn_concat_src=bool(ref.dec_attn_concat_src), dec_qkv_proj=isinstance(layer0.attn_q, nn.Linear),
        dec_layer_type="v1", pad_idx=dec.token_embed.padding_idx, use_anchor=bool(ref.use_anchor))


def _live_transformer_state(ref) -> dict:
    """``ref.state_dict()`` without the cache buffers the reference's ``_setup_caches`` registers on every
    ``forward_inference`` and without the heads (attached as live modules instead)."""
    # ... and without the two module lists CAPEModel hangs on the decoder for the duration of a call
    # (cape_model.py:127-132, 187-194): attribute assignment registers them as sub-modules, the decoder never runs them
    # 
    skip = ("decoder.class_embed.", "decoder.coords_embed.", "decoder.support_cross_attn_layers.",
            "decoder.support_attn_norms.")
    return {k: v for k, v in ref.state_dict().items()
            if ".kv_cache." not in k and ".cross_attn.cache." not in k and not k.startswith(skip)}


def mirror_from_reference(ref) -> DeformableTransformer:
    """A :class:`DeformableTransformer` whose parameters ARE the live reference transformer's tensors (same storage:
    ``load_state_dict(assign=True)``), so optimizer steps and in-place loads on the reference are seen without a copy.
    The prediction heads are the reference's own modules (roomformer_v2.py:245-246).  Use :func:`mirror_is_current`
    to detect a re-allocation (``model.to(...)``, a non-in-place load) and rebuild."""
    with torch.device("meta"):
        mirror = DeformableTransformer(**_reference_transformer_config(ref))
    state = _live_transformer_state(ref)
    missing, unexpected = mirror.load_state_dict(state, strict=False, assign=True)
    missing = [k for k in missing if not k.startswith(("decoder.class_embed.", "decoder.coords_embed."))]
    if missing or unexpected:
        raise RuntimeError(f"reference transformer does not match the mirror: missing {missing[:5]}, "
                           f"unexpected {unexpected[:5]}")
    mirror.attach_heads(ref.decoder.class_embed, ref.decoder.coords_embed)
    mirror.train(ref.training)
    mirror._bound_ptrs = {k: v.data_ptr() for k, v in state.items()}
    return mirror


def mirror_is_current(mirror: DeformableTransformer, ref) -> bool:
    ptrs = getattr(mirror, "_bound_ptrs", None)
    if ptrs is None:
        return False
    state = _live_transformer_state(ref)
    return len(state) == len(ptrs) and all(ptrs.get(k) == v.data_ptr() for k, v in state.items())
