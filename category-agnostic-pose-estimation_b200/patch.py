"""Zero-edit integration with the reference code base.

``MSDeformAttn.forward`` looks ``ms_deform_attn_core_pytorch`` up in its module's globals at call time
(``/root/reference/models/deformable_transformer.py:112``), so rebinding that one name swaps the hot path under an
unmodified ``CAPEModel`` — parameters, ``state_dict`` keys, ``forward`` / ``forward_inference`` API all untouched.

Levels (combinable):

* ``patch_reference()``                              the sampling core only (training and inference);
* ``patch_reference(swap_module_class=True)``        also the ``MSDeformAttn`` class for models built afterwards
                                                     (fused prologue, ``use_cache`` honoured);
* ``patch_reference(swap_layer_classes=True)``       also the encoder layer / encoder / decoder layer v1 classes (mirrors with
                                                     identical ``state_dict`` keys), which opens their FFNs to the opt-in
                                                     tensor-core linears;
* ``patch_reference(swap_forward_inference=True)``   also ``RoomFormerV2.forward_inference`` (``roomformer_v2.py:385-677``):
                                                     the autoregressive loop runs device-resident on a mirror bound to the
                                                     live module's own parameter tensors, and returns the reference's dict,
                                                     so ``CAPEModel.forward_inference`` (``cape_model.py:142-209``) is
                                                     called unchanged.
"""
from __future__ import annotations

import importlib
from typing import Optional

import torch
import torch.nn.functional as F

from . import functional as CF
from .modules import MSDeformAttn

_ORIGINALS = {}
_INFERENCE_ORIGINALS = {}
_LAYER_ORIGINALS = {}
_GEN_ATTR = "_cape_b200_generation"      # plain attribute on the live module: never a submodule, buffer or parameter


_LAYER_CLASSES = ("DeformableTransformerEncoderLayer", "DeformableTransformerEncoder", "TransformerDecoderLayer")


def patch_reference(module=None, swap_module_class: bool = False, swap_forward_inference: bool = False,
                    swap_layer_classes: bool = False):
    """Rebind the reference's sampling core to the B200 op.

    module: the imported ``models.deformable_transformer`` module object (or its dotted name; default
        ``"models.deformable_transformer"``).
    swap_module_class: also replace the ``MSDeformAttn`` class (in that module and in
        ``models.deformable_transformer_v2``, which imports it by name, :17) by the mirror that honours
        ``use_cache`` and fuses the decode prologue.  Only affects models built after the call.
    swap_forward_inference: also replace ``RoomFormerV2.forward_inference`` in the sibling ``roomformer_v2`` module
        by :func:`forward_inference` below.  Affects existing model objects too (the method is looked up on the class).
    swap_layer_classes: implies ``swap_module_class`` and also replaces the callers either side of the op —
        ``DeformableTransformerEncoderLayer`` / ``DeformableTransformerEncoder`` (``deformable_transformer.py:155-291``) and
        decoder layer v1 ``TransformerDecoderLayer`` (``deformable_transformer_v2.py:262-370``) — by the mirrors of
        ``layers.py`` (same constructor arguments, parameter names and ``state_dict`` keys).  Their ``nn.Linear`` calls
        go through ``gemm.linear``, so ``cape_b200.set_linear_mode("tf32x3")`` then puts the encoder / decoder FFNs and the
        MSDeformAttn projections of a model built afterwards on the tcgen05 3xTF32 GEMM.  Only affects models built after
        the call.
    Returns the patched module.
    """
    if module is None or isinstance(module, str):
        module = importlib.import_module(module or "models.deformable_transformer")
    key = id(module)
    if key not in _ORIGINALS:
        _ORIGINALS[key] = (module, module.ms_deform_attn_core_pytorch, getattr(module, "MSDeformAttn", None))
    module.ms_deform_attn_core_pytorch = CF.ms_deform_attn_core_pytorch
    package = module.__name__.rsplit(".", 1)[0] if "." in module.__name__ else ""
    sibling = lambda name: importlib.import_module((package + "." if package else "") + name)
    if swap_layer_classes:
        swap_module_class = True
    if swap_module_class:
        module.MSDeformAttn = MSDeformAttn
        try:
            v2 = sibling("deformable_transformer_v2")
            if hasattr(v2, "MSDeformAttn"):
                _ORIGINALS.setdefault(id(v2), (v2, None, v2.MSDeformAttn))
                v2.MSDeformAttn = MSDeformAttn
        except Exception:
            pass
    if swap_layer_classes:
        from . import layers as CL
        targets = [module]
        try:
            targets.append(sibling("deformable_transformer_v2"))
        except Exception:
            pass
        for mod in targets:
            saved = _LAYER_ORIGINALS.setdefault(id(mod), (mod, {}))[1]
            for name in _LAYER_CLASSES:
                if hasattr(mod, name) and hasattr(CL, name):
                    saved.setdefault(name, getattr(mod, name))
                    setattr(mod, name, getattr(CL, name))
    if swap_forward_inference:
        rf = sibling("roomformer_v2")
        cls = rf.RoomFormerV2
        if id(cls) not in _INFERENCE_ORIGINALS:
            _INFERENCE_ORIGINALS[id(cls)] = (cls, cls.forward_inference)
            cls.forward_inference = _make_forward_inference(rf, cls.forward_inference)
    return module


def unpatch_reference(module: Optional[object] = None):
    """Undo :func:`patch_reference` (all patched modules when ``module`` is None)."""
    keys = [id(module)] if module is not None else list(_ORIGINALS)
    for k in keys:
        if k in _ORIGINALS:
            mod, core, cls = _ORIGINALS.pop(k)
            if core is not None:
                mod.ms_deform_attn_core_pytorch = core
            if cls is not None:
                mod.MSDeformAttn = cls
    if module is None:
        for cls, original in _INFERENCE_ORIGINALS.values():
            cls.forward_inference = original
        _INFERENCE_ORIGINALS.clear()
        for mod, saved in _LAYER_ORIGINALS.values():
            for name, cls in saved.items():
                setattr(mod, name, cls)
        _LAYER_ORIGINALS.clear()


# ---- RoomFormerV2.forward_inference, device-resident ----------------------------------------------------------------
def _feature_pyramid(model, samples, rf):
    """Backbone + input projections + extra levels: what ``forward_inference`` does before the decoder part
    (roomformer_v2.py:403-440).  Uses the live model's own backbone / ``input_proj`` modules."""
    nested = rf.NestedTensor
    if not isinstance(samples, nested):
        samples = rf.nested_tensor_from_tensor_list(samples)
    features, pos = model.backbone(samples)
    srcs, masks = [], []
    for l, feat in enumerate(features):
        src, mask = feat.decompose()
        src = model.input_proj[l](src)
        srcs.append(src)
        if model.patch_size != 1:
            mask = F.interpolate(mask[None].float(), size=src.shape[-2:]).to(torch.bool)[0]
            pos[l] = model.backbone[1](nested(src, mask)).to(src.dtype)
        masks.append(mask)
    for l in range(len(srcs), model.num_feature_levels):
        src = model.input_proj[l](features[-1].tensors if l == len(features) else srcs[-1])
        mask = F.interpolate(samples.mask[None].float(), size=src.shape[-2:]).to(torch.bool)[0]
        pos.append(model.backbone[1](nested(src, mask)).to(src.dtype))
        srcs.append(src)
        masks.append(mask)
    return srcs, masks, pos


def _generation_state(model, batch: int, device):
    """Mirror transformer bound to ``model.transformer``'s own tensors + one generator per batch size, cached on the
    live module as a plain attribute (``state_dict()`` unaffected).  Rebuilt when the live parameters were re-allocated."""
    from .sequence import TokenizerSpec
    from .transformer import AutoregressiveGenerator, mirror_from_reference, mirror_is_current
    state = model.__dict__.get(_GEN_ATTR)
    if state is None or not mirror_is_current(state["mirror"], model.transformer):
        state = {"mirror": mirror_from_reference(model.transformer), "generators": {}}
        model.__dict__[_GEN_ATTR] = state
    mirror = state["mirror"]
    mirror.train(False)
    spec = TokenizerSpec.from_tokenizer(model.tokenizer)
    key = (batch, str(device), spec.seq_len, spec.num_bins, spec.add_cls)
    gen = state["generators"].get(key)
    if gen is None:
        gen = AutoregressiveGenerator(mirror, spec, batch, device)
        state["generators"] = {key: gen}          # one resident generator: its K/V + value caches are batch-sized
    return mirror, gen


def _make_forward_inference(rf, original):
    def forward_inference(self, samples, use_cache=True, support_graphs=None, support_mask=None):
        """Drop-in for ``RoomFormerV2.forward_inference`` (roomformer_v2.py:385-677): same arguments, same result dict
        (``pred_logits`` (B, steps, n_classes), ``pred_coords`` (B, steps, 2), ``gen_out``).  The decoder part — the
        ``while i < max_len and unfinish_flag.any()`` loop with its per-sample Python bookkeeping (:481-598) — runs as
        one CUDA graph per token on the device.  Configurations outside the mirror (CPU tensors, ``cape_mode`` with an
        internal support encoder, ``inject_cls_embed``, ``use_cache=False``, decoder layers other than v1) take the reference's
        own loop, which still samples through the patched core."""
        tensors = samples.tensors if hasattr(samples, "tensors") else samples
        unsupported = (not isinstance(tensors, torch.Tensor) or not tensors.is_cuda or not use_cache
                       or (getattr(self, "cape_mode", False) and support_graphs is not None)
                       or getattr(self, "inject_cls_embed", False) or self.query_embed is None
                       or getattr(self, "tokenizer", None) is None)
        if not unsupported:
            try:
                from .transformer import _reference_transformer_config
                _reference_transformer_config(self.transformer)
            except NotImplementedError:
                unsupported = True
        if unsupported:
            return original(self, samples, use_cache=use_cache, support_graphs=support_graphs, support_mask=support_mask)
        with torch.no_grad():
            srcs, masks, pos = _feature_pyramid(self, samples, rf)
            bs = srcs[0].shape[0]
            mirror, gen = _generation_state(self, bs, srcs[0].device)
            dec = self.transformer.decoder
            out = gen.generate(srcs, masks, pos, self.query_embed.weight,
                               support_features=getattr(dec, "support_features", None),
                               support_mask=getattr(dec, "support_mask", None))
            result = {"pred_logits": out["pred_logits"], "pred_coords": out["pred_coords"], "gen_out": out["gen_out"]}
            if getattr(self, "room_class_embed", None) is not None:          # :647-654 (semantic_classes > 0, the CAPE default)
                result = {"pred_logits": out["pred_logits"], "pred_coords": out["pred_coords"],
                          "pred_room_logits": self.room_class_embed(out["hidden"]), "gen_out": out["gen_out"],
                          "anchors": self.query_embed.weight.detach()}
        return result

    forward_inference.__wrapped__ = original
    return forward_inference
