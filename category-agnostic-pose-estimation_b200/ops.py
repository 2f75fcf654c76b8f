"""torch.library registration of the MSDeformAttn ops (namespace ``cape``).

    cape::ms_deform_attn(value, spatial_shapes, level_start_index, sampling_locations, attention_weights) -> out
    cape::ms_deform_attn_backward(grad_out, value, spatial_shapes, level_start_index, sampling_locations,
                                  attention_weights) -> (grad_value, grad_sampling_loc, grad_attn_weight)
    cape::ms_deform_attn_decode(value_cache, spatial_shapes, level_start_index, reference_points,
                                sampling_offsets, attention_logits) -> out
    cape::ms_deform_attn_fused_backward(grad_out, value, spatial_shapes, level_start_index, reference_points,
                                        sampling_offsets, attention_logits) -> (grad_value, grad_offsets, grad_logits)

``ms_deform_attn_decode`` is the module-level fused op: softmax over the L*P logits and ``ref + off / (W_l, H_l)``
(``/root/reference/models/deformable_transformer.py:100-105``) happen inside the sampling kernel.  It serves the decode
step (one new token on a cached projected value) and, being differentiable, the training path of the module too —
``sampling_locations`` / ``attention_weights`` and their gradients then never exist in HBM.

The argument order is upstream Deformable-DETR's ``MSDeformAttnFunction`` (what the reference's vestigial
``im2col_step`` at ``/root/reference/models/deformable_transformer.py:51`` was for); semantics are those of
``ms_deform_attn_core_pytorch`` (``:115-141``).  CUDA only: the ops are registered for the CUDA dispatch key alone,
so CPU tensors raise ``NotImplementedError`` from the dispatcher, and a missing shared library raises
``CapeLibraryError`` — there is no fallback path.

Host work per call: one ctypes call on the current torch stream.  No device synchronisation (``spatial_shapes`` and
``level_start_index`` are read on the device, unlike the reference which iterates the CUDA tensor in Python,
``:130,133``).
"""
from __future__ import annotations

import ctypes
from typing import Tuple

import torch

from . import _lib

_DTYPE_CODE = {torch.float32: _lib.DTYPE_F32, torch.bfloat16: _lib.DTYPE_BF16, torch.float16: _lib.DTYPE_F16}


def _dims(value, loc) -> "_lib.Dims":
    n, s, m, d = value.shape
    lq, l, p = loc.shape[1], loc.shape[3], loc.shape[4]
    return _lib.Dims(n, s, m, d, lq, l, p)


def _check_shapes(value, spatial_shapes, level_start_index, loc, attn):
    if value.dim() != 4:
        raise ValueError(f"value must be (N, S, M, D), got {tuple(value.shape)}")
    if loc.dim() != 6 or loc.shape[-1] != 2:
        raise ValueError(f"sampling_locations must be (N, Lq, M, L, P, 2), got {tuple(loc.shape)}")
    n, _, m, _ = value.shape
    if loc.shape[0] != n or loc.shape[2] != m:
        raise ValueError(f"sampling_locations {tuple(loc.shape)} does not match value {tuple(value.shape)}")
    if tuple(attn.shape) != tuple(loc.shape[:5]):
        raise ValueError(f"attention_weights must be {tuple(loc.shape[:5])}, got {tuple(attn.shape)}")
    l = loc.shape[3]
    if tuple(spatial_shapes.shape) != (l, 2) or tuple(level_start_index.shape) != (l,):
        raise ValueError(f"spatial_shapes must be ({l}, 2) and level_start_index ({l},), got "
                         f"{tuple(spatial_shapes.shape)} and {tuple(level_start_index.shape)}")


def _meta(t: torch.Tensor, device) -> torch.Tensor:
    """int64, contiguous, on the op's device (spatial_shapes / level_start_index)."""
    if t.dtype != torch.int64 or t.device != device or not t.is_contiguous():
        t = t.to(device=device, dtype=torch.int64, non_blocking=True).contiguous()
    return t


def _aux(loc, attn, value_dtype):
    """The kernels take loc/attn either both fp32 or both in the value dtype."""
    if loc.dtype == attn.dtype and (loc.dtype == torch.float32 or loc.dtype == value_dtype):
        return _c16(loc), _c16(attn)
    return _c16(loc.float()), _c16(attn.float())


def _c16(t: torch.Tensor) -> torch.Tensor:
    """Contiguous AND 16-byte aligned (the kernels use 128-bit accesses): ``.contiguous()`` is a no-op on a contiguous
    view with a storage offset (``loc[1:]``, a slice of a fused buffer), which the ABI would reject as misaligned."""
    t = t.contiguous()
    return t.clone() if t.data_ptr() % 16 else t


def _ptr(t) -> ctypes.c_void_p:
    return ctypes.c_void_p(t.data_ptr())


def _stream(device) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


@torch.library.custom_op("cape::ms_deform_attn", mutates_args=(), device_types="cuda")
def ms_deform_attn(value: torch.Tensor, spatial_shapes: torch.Tensor, level_start_index: torch.Tensor,
                   sampling_locations: torch.Tensor, attention_weights: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    _check_shapes(value, spatial_shapes, level_start_index, sampling_locations, attention_weights)
    if value.dtype not in _DTYPE_CODE:
        raise TypeError(f"unsupported value dtype {value.dtype}")
    value = _c16(value)
    loc, attn = _aux(sampling_locations, attention_weights, value.dtype)
    shapes = _meta(spatial_shapes, value.device)
    starts = _meta(level_start_index, value.device)
    dims = _dims(value, loc)
    out = torch.empty((dims.N, dims.Lq, dims.M * dims.D), dtype=value.dtype, device=value.device)
    with torch.cuda.device(value.device):
        rc = lib.cape_msda_forward(_ptr(value), _ptr(shapes), _ptr(starts), _ptr(loc), _ptr(attn), _ptr(out),
                                   ctypes.byref(dims), _DTYPE_CODE[value.dtype], _DTYPE_CODE[loc.dtype],
                                   _stream(value.device))
    _lib.check(rc, "cape_msda_forward")
    return out


@ms_deform_attn.register_fake
def _(value, spatial_shapes, level_start_index, sampling_locations, attention_weights):
    n, _, m, d = value.shape
    return value.new_empty((n, sampling_locations.shape[1], m * d))


@torch.library.custom_op("cape::ms_deform_attn_backward", mutates_args=(), device_types="cuda")
def ms_deform_attn_backward(grad_out: torch.Tensor, value: torch.Tensor, spatial_shapes: torch.Tensor,
                            level_start_index: torch.Tensor, sampling_locations: torch.Tensor,
                            attention_weights: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    lib = _lib.load()
    _check_shapes(value, spatial_shapes, level_start_index, sampling_locations, attention_weights)
    value = _c16(value)
    loc, attn = _aux(sampling_locations, attention_weights, value.dtype)
    shapes = _meta(spatial_shapes, value.device)
    starts = _meta(level_start_index, value.device)
    dims = _dims(value, loc)
    grad_out = _c16(grad_out.to(value.dtype))
    grad_value = torch.empty(value.shape, dtype=torch.float32, device=value.device)   # zeroed by the library
    grad_loc = torch.empty_like(loc)
    grad_attn = torch.empty_like(attn)
    with torch.cuda.device(value.device):
        rc = lib.cape_msda_backward(_ptr(grad_out), _ptr(value), _ptr(shapes), _ptr(starts), _ptr(loc), _ptr(attn),
                                    _ptr(grad_value), _ptr(grad_loc), _ptr(grad_attn), ctypes.byref(dims),
                                    _DTYPE_CODE[value.dtype], _DTYPE_CODE[loc.dtype], 1, _stream(value.device))
    _lib.check(rc, "cape_msda_backward")
    return (grad_value.to(value.dtype), grad_loc.to(sampling_locations.dtype), grad_attn.to(attention_weights.dtype))


@ms_deform_attn_backward.register_fake
def _(grad_out, value, spatial_shapes, level_start_index, sampling_locations, attention_weights):
    return (torch.empty_like(value), torch.empty_like(sampling_locations), torch.empty_like(attention_weights))


def _setup_context(ctx, inputs, output):
    value, spatial_shapes, level_start_index, loc, attn = inputs
    ctx.save_for_backward(value, spatial_shapes, level_start_index, loc, attn)


def _backward(ctx, grad_out):
    value, spatial_shapes, level_start_index, loc, attn = ctx.saved_tensors
    gv, gl, ga = torch.ops.cape.ms_deform_attn_backward(grad_out, value, spatial_shapes, level_start_index, loc, attn)
    return gv, None, None, gl, ga


ms_deform_attn.register_autograd(_backward, setup_context=_setup_context)

# The reference runs this path in fp32 under autocast (grid_sampler is on autocast's fp32 list):
# same policy here, so AMP training sees identical dtypes at the seam.
torch.library.register_autocast("cape::ms_deform_attn", "cuda", torch.float32)


@torch.library.custom_op("cape::ms_deform_attn_decode", mutates_args=(), device_types="cuda")
def ms_deform_attn_decode(value_cache: torch.Tensor, spatial_shapes: torch.Tensor, level_start_index: torch.Tensor,
                          reference_points: torch.Tensor, sampling_offsets: torch.Tensor,
                          attention_logits: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    if value_cache.dim() != 4 or sampling_offsets.dim() != 6 or sampling_offsets.shape[-1] != 2:
        raise ValueError("value_cache must be (B,S,M,D) and sampling_offsets (B,k,M,L,P,2)")
    b, k, m, l, p, _ = sampling_offsets.shape
    if tuple(reference_points.shape) != (b, k, l, 2):
        raise ValueError(f"reference_points must be {(b, k, l, 2)}, got {tuple(reference_points.shape)}")
    if attention_logits.numel() != b * k * m * l * p:
        raise ValueError("attention_logits must have B*k*M*L*P elements")
    if value_cache.dtype not in _DTYPE_CODE:
        raise TypeError(f"unsupported value dtype {value_cache.dtype}")
    value_cache = _c16(value_cache)
    ref = _c16(reference_points.float())
    off = _c16(sampling_offsets.float())
    logits = _c16(attention_logits.float())
    shapes = _meta(spatial_shapes, value_cache.device)
    starts = _meta(level_start_index, value_cache.device)
    dims = _lib.Dims(b, value_cache.shape[1], m, value_cache.shape[3], k, l, p)
    out = torch.empty((b, k, m * value_cache.shape[3]), dtype=value_cache.dtype, device=value_cache.device)
    with torch.cuda.device(value_cache.device):
        rc = lib.cape_msda_decode(_ptr(value_cache), _ptr(shapes), _ptr(starts), _ptr(ref), _ptr(off), _ptr(logits),
                                  _ptr(out), ctypes.byref(dims), _DTYPE_CODE[value_cache.dtype],
                                  _stream(value_cache.device))
    _lib.check(rc, "cape_msda_decode")
    return out


@ms_deform_attn_decode.register_fake
def _(value_cache, spatial_shapes, level_start_index, reference_points, sampling_offsets, attention_logits):
    b, k = sampling_offsets.shape[:2]
    return value_cache.new_empty((b, k, value_cache.shape[2] * value_cache.shape[3]))


@torch.library.custom_op("cape::ms_deform_attn_fused_backward", mutates_args=(), device_types="cuda")
def ms_deform_attn_fused_backward(grad_out: torch.Tensor, value: torch.Tensor, spatial_shapes: torch.Tensor,
                                  level_start_index: torch.Tensor, reference_points: torch.Tensor,
                                  sampling_offsets: torch.Tensor, attention_logits: torch.Tensor
                                  ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    lib = _lib.load()
    b, k, m, l, p, _ = sampling_offsets.shape
    value = _c16(value)
    dims = _lib.Dims(b, value.shape[1], m, value.shape[3], k, l, p)
    shapes = _meta(spatial_shapes, value.device)
    starts = _meta(level_start_index, value.device)
    ref = _c16(reference_points.float())
    off = _c16(sampling_offsets.float())
    logits = _c16(attention_logits.float())
    grad_out = _c16(grad_out.to(value.dtype))
    if lib.cape_msda_fused_supported(ctypes.byref(dims)):
        grad_value = torch.empty(value.shape, dtype=torch.float32, device=value.device)
        grad_off = torch.empty_like(off)
        grad_logits = torch.empty_like(logits)
        with torch.cuda.device(value.device):
            rc = lib.cape_msda_fused_backward(_ptr(grad_out), _ptr(value), _ptr(shapes), _ptr(starts), _ptr(ref),
                                              _ptr(off), _ptr(logits), _ptr(grad_value), _ptr(grad_off),
                                              _ptr(grad_logits), ctypes.byref(dims), _DTYPE_CODE[value.dtype], 1,
                                              _stream(value.device))
        _lib.check(rc, "cape_msda_fused_backward")
    else:
        # dimensions outside the fused fast path: same mathematics composed from the unfused backward
        attn = torch.softmax(logits.reshape(b, k, m, l * p), -1)
        wh = torch.stack([shapes[:, 1], shapes[:, 0]], -1).to(off.dtype)
        loc = ref[:, :, None, :, None, :] + off / wh[None, None, None, :, None, :]
        grad_value, grad_loc, grad_attn = torch.ops.cape.ms_deform_attn_backward(
            grad_out, value, shapes, starts, loc.contiguous(), attn.view(b, k, m, l, p))
        grad_off = grad_loc / wh[None, None, None, :, None, :]
        ga = grad_attn.reshape(b, k, m, l * p)
        grad_logits = attn * (ga - (attn * ga).sum(-1, keepdim=True))
    return (grad_value.to(value.dtype), grad_off.to(sampling_offsets.dtype),
            grad_logits.reshape(attention_logits.shape).to(attention_logits.dtype))


@ms_deform_attn_fused_backward.register_fake
def _(grad_out, value, spatial_shapes, level_start_index, reference_points, sampling_offsets, attention_logits):
    return (torch.empty_like(value), torch.empty_like(sampling_offsets), torch.empty_like(attention_logits))


def _fused_setup_context(ctx, inputs, output):
    ctx.save_for_backward(*inputs)
    ctx.needs_ref_grad = inputs[3].requires_grad


def _fused_backward(ctx, grad_out):
    value, shapes, starts, ref, off, logits = ctx.saved_tensors
    gv, goff, glogits = torch.ops.cape.ms_deform_attn_fused_backward(grad_out, value, shapes, starts, ref, off, logits)
    gref = None
    if ctx.needs_ref_grad:   # d loc / d ref = 1: sum the location gradient over heads and points
        wh = torch.stack([shapes[:, 1], shapes[:, 0]], -1).to(goff.dtype)
        gref = (goff * wh[None, None, None, :, None, :]).sum(dim=(2, 4)).to(ref.dtype)
    return gv, None, None, gref, goff, glogits


ms_deform_attn_decode.register_autograd(_fused_backward, setup_context=_fused_setup_context)
# No autocast rule for the fused op on purpose: it takes `value` in whatever dtype the caller holds it (fp32, or the
# fp16 / bf16 that an autocast value_proj produced), converts on load, accumulates in fp32 and casts the small tensors
# itself.  A blanket cast-to-fp32 rule would copy the whole (B, S, M, D) value — for the decode step that is the entire
# projected-value cache — on every call.  The result equals the reference's AMP flow, whose fp32 sampling output is
# rounded to half precision by the next autocast Linear anyway.


# ---- padding-mask fill of the projected value (deformable_transformer.py:96-97) ---------------------------------------
def _zero_masked_rows_(t: torch.Tensor, mask: torch.Tensor) -> None:
    lib = _lib.load()
    rows = mask.numel()
    row_bytes = (t.numel() // max(rows, 1)) * t.element_size()
    with torch.cuda.device(t.device):
        rc = lib.cape_zero_masked_rows(_ptr(t), _ptr(mask), rows, row_bytes, _stream(t.device))
    _lib.check(rc, "cape_zero_masked_rows")


class _MaskedFillRows(torch.autograd.Function):
    """``value.masked_fill(mask[..., None], 0)`` done IN PLACE on a tensor the caller owns (the fresh output of
    value_proj), forward and backward, by a kernel that inspects the mask on the device and leaves un-masked blocks
    untouched: for the all-False mask CAPE always passes, neither direction reads or writes the value tensor, and there is
    no ``mask.any()`` host synchronisation.  The gradient arriving in backward is the freshly allocated ``grad_value`` of
    the sampling op (single consumer), so it is zeroed in place as well."""

    @staticmethod
    def forward(ctx, value, mask):
        ctx.mark_dirty(value)
        ctx.save_for_backward(mask)
        _zero_masked_rows_(value, mask)
        return value

    @staticmethod
    def backward(ctx, grad):
        (mask,) = ctx.saved_tensors
        if not grad.is_contiguous() or grad.data_ptr() % 16:
            grad = grad.contiguous().clone()
        _zero_masked_rows_(grad, mask)
        return grad, None


def masked_fill_rows_(value: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """value (..., C) contiguous CUDA tensor produced by the caller (not a leaf, not shared), mask (...) bool."""
    if mask.dtype != torch.bool:
        mask = mask.bool()
    mask = mask.contiguous()
    row_bytes = value.shape[-1] * value.element_size()
    if (not value.is_cuda or not value.is_contiguous() or value.data_ptr() % 16 or row_bytes % 16
            or value.numel() != mask.numel() * value.shape[-1] or value.is_leaf and value.requires_grad):
        return value.masked_fill(mask[..., None], float(0))
    if value.requires_grad and torch.is_grad_enabled():
        return _MaskedFillRows.apply(value, mask)
    _zero_masked_rows_(value, mask)
    return value
