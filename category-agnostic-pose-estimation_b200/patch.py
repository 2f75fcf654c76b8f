"""Zero-edit integration with the reference code base.

``MSDeformAttn.forward`` looks ``ms_deform_attn_core_pytorch`` up in its module's globals at call time
(``/root/reference/models/deformable_transformer.py:112``), so rebinding that one name swaps the hot path under an
unmodified ``CAPEModel`` — parameters, ``state_dict`` keys, ``forward`` / ``forward_inference`` API all untouched.
"""
from __future__ import annotations

import importlib
from typing import Optional

from . import functional as CF
from .modules import MSDeformAttn

_ORIGINALS = {}


def patch_reference(module=None, swap_module_class: bool = False):
    """Rebind the reference's sampling core to the B200 op.

    module: the imported ``models.deformable_transformer`` module object (or its dotted name; default
        ``"models.deformable_transformer"``).
    swap_module_class: also replace the ``MSDeformAttn`` class (in that module and in
        ``models.deformable_transformer_v2``, which imports it by name, :17) by the mirror that honours
        ``use_cache`` and fuses the decode prologue.  Only affects models built after the call.
    Returns the patched module.
    """
    if module is None or isinstance(module, str):
        module = importlib.import_module(module or "models.deformable_transformer")
    key = id(module)
    if key not in _ORIGINALS:
        _ORIGINALS[key] = (module, module.ms_deform_attn_core_pytorch, getattr(module, "MSDeformAttn", None))
    module.ms_deform_attn_core_pytorch = CF.ms_deform_attn_core_pytorch
    if swap_module_class:
        module.MSDeformAttn = MSDeformAttn
        try:
            v2 = importlib.import_module(module.__name__.rsplit(".", 1)[0] + ".deformable_transformer_v2")
            if hasattr(v2, "MSDeformAttn"):
                _ORIGINALS.setdefault(id(v2), (v2, None, v2.MSDeformAttn))
                v2.MSDeformAttn = MSDeformAttn
        except Exception:
            pass
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
