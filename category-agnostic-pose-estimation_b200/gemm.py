"""fp32-accurate linear layers on the tensor cores (``csrc/linear_tf32x3.cu``: tcgen05.mma.kind::tf32 fed by TMA, every
operand used as hi + lo, three products accumulated in fp32 in tensor memory).

The reference runs the projections around the sampling op (``value_proj`` / ``output_proj`` / ``sampling_offsets`` /
``attention_weights`` of MSDeformAttn, the encoder FFN; ``/root/reference/models/deformable_transformer.py:95-113,219-224``)
as strict-fp32 ``nn.Linear``, which cuBLAS serves with SIMT kernels on B200.  ``linear()`` is what the mirrors call
instead of ``module(x)``: by default it IS ``module(x)`` (bit-for-bit the reference's arithmetic); after
``set_linear_mode("tf32x3")`` calls with supported shapes (K % 32 == 0, N % 128 == 0, fp32, CUDA, >= 128 rows, no
autocast) go through the 3xTF32 kernel — max error about 2e-6 of the output scale at K = 256 (cuBLAS fp32: 5e-7), ~5x
faster.  With autograd on, forward, the input gradient (the same GEMM with W^T) AND the weight gradient (transposed
operands, reduction split over the SMs, partial tiles added by the TMA) all use the kernel (``_LinearTF32x3``).

Derived operands (the lo part of a weight, W^T) are cached per weight tensor and refreshed when its ``_version`` changes,
which every optimizer step and ``load_state_dict`` does.  Writes through ``weight.data`` (``w.data.copy_()``, hand-rolled
EMA) do NOT bump the version: call :func:`clear_caches` after such an update.
"""
from __future__ import annotations

import ctypes
import weakref

import torch
import torch.nn.functional as F

from . import _lib
from .ops import _ptr, _stream

_MODE = "fp32"
_LO_CACHE: dict = {}


def set_linear_mode(mode: str) -> str:
    """"fp32" (default: plain nn.Linear) or "tf32x3" (3xTF32 tensor-core kernel, inference and training).  Returns the old mode."""
    global _MODE
    if mode not in ("fp32", "tf32x3"):
        raise ValueError(f"unknown linear mode {mode!r}")
    old, _MODE = _MODE, mode
    return old


def linear_mode() -> str:
    return _MODE


def clear_caches() -> None:
    """Forget the cached lo parts / transposed weights (needed only after weight updates that bypass the version counter,
    e.g. writes through ``.data``)."""
    _LO_CACHE.clear()
    _WT_CACHE.clear()


def _weight_lo(weight: torch.Tensor) -> torch.Tensor:
    """lo part of a weight (cached per tensor object and version; recomputed after an in-place update)."""
    key = id(weight)
    hit = _LO_CACHE.get(key)
    same_tensor = hit is not None and hit[0]() is weight and hit[2] == weight.data_ptr()
    if same_tensor and hit[1] == weight._version:
        return hit[3]
    lib = _lib.load()
    w = weight.detach()
    # after an optimizer step only the version changes: refill the same buffer (a fresh allocation per weight and step
    # keeps the caching allocator from settling and costs cudaMallocs inside the training loop)
    lo = hit[3] if same_tensor and hit[3].shape == w.shape else torch.empty_like(w)
    with torch.cuda.device(w.device):
        _lib.check(lib.cape_tf32_split_lo(_ptr(w), _ptr(lo), w.numel(), _stream(w.device)), "cape_tf32_split_lo")
    if len(_LO_CACHE) > 512:
        _LO_CACHE.clear()
    _LO_CACHE[key] = (weakref.ref(weight), weight._version, weight.data_ptr(), lo)
    return lo


def supported(x: torch.Tensor, weight: torch.Tensor) -> bool:
    n, k = weight.shape
    return (x.is_cuda and x.dtype == torch.float32 and weight.dtype == torch.float32 and weight.is_contiguous()
            and k % 32 == 0 and n % 128 == 0 and x.shape[-1] == k and x.numel() > 0)


def linear_tf32x3(x: torch.Tensor, weight: torch.Tensor, bias=None, relu: bool = False) -> torch.Tensor:
    """act(x @ weight.T + bias) through the 3xTF32 kernel.  x (..., K) fp32 CUDA; weight (N, K) contiguous."""
    if not supported(x, weight):
        raise ValueError(f"linear_tf32x3: unsupported operands x {tuple(x.shape)} {x.dtype}, weight {tuple(weight.shape)}")
    lib = _lib.load()
    n, k = weight.shape
    x2 = x.detach().reshape(-1, k)
    if not x2.is_contiguous() or x2.data_ptr() % 16 != 0:
        x2 = x2.contiguous()
    w = weight.detach()
    y = torch.empty(*x.shape[:-1], n, dtype=torch.float32, device=x.device)   # final shape: not a view (callers may write in place)
    b = None if bias is None else bias.detach().contiguous()
    with torch.cuda.device(x.device):
        rc = lib.cape_linear_tf32x3(_ptr(x2), _ptr(w), _ptr(_weight_lo(weight)), None if b is None else _ptr(b), _ptr(y),
                                    x2.shape[0], n, k, 1 if relu else 0, _stream(x.device))
    _lib.check(rc, "cape_linear_tf32x3")
    return y


_WT_CACHE: dict = {}


def _weight_t(weight: torch.Tensor) -> torch.Tensor:
    """W^T (K, N) contiguous, cached per tensor version: the operand of the input-gradient GEMM g . W."""
    key = id(weight)
    hit = _WT_CACHE.get(key)
    same_tensor = hit is not None and hit[0]() is weight and hit[2] == weight.data_ptr()
    if same_tensor and hit[1] == weight._version:
        return hit[3]
    if same_tensor and hit[3].shape == (weight.shape[1], weight.shape[0]):
        wt = hit[3]
        wt.copy_(weight.detach().t())                      # in place: same buffer, new version (its lo part follows)
    else:
        wt = weight.detach().t().contiguous()
    if len(_WT_CACHE) > 512:
        _WT_CACHE.clear()
    _WT_CACHE[key] = (weakref.ref(weight), weight._version, weight.data_ptr(), wt)
    return wt


_WORKSPACE: dict = {}


def _workspace(numel: int, device) -> torch.Tensor:
    """Grow-only scratch for the transposed operands of the weight gradient (up to 0.7 GB at the FFN shapes): one buffer
    per device and stream instead of an allocation per call.  Safe to share: every use is enqueued on the stream it is
    keyed by, and the kernels that read it are enqueued before the next use overwrites it."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    buf = _WORKSPACE.get(key)
    if buf is None or buf.numel() < numel:
        buf = torch.empty(numel, dtype=torch.float32, device=device)
        _WORKSPACE[key] = buf
    return buf


def wgrad_supported(g2: torch.Tensor, x2: torch.Tensor) -> bool:
    rows, n = g2.shape
    k = x2.shape[1]
    return (g2.is_cuda and g2.dtype == torch.float32 and x2.dtype == torch.float32 and rows >= 1024
            and n % 128 == 0 and k % 128 == 0)


def linear_tf32x3_wgrad(grad_out: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """grad_w (N, K) = grad_out (rows, N)^T @ x (rows, K) through the 3xTF32 kernel: both operands are read where they are
    (MN-major tiles, lo parts split in shared memory), the rows are split over the SMs in chains of <= 1024 whose partial tiles
    the TMA adds into grad_w.  (Tuning knob WGRAD_TRANSPOSE=1: the earlier form on transposed copies, kept for A/B runs.)"""
    lib = _lib.load()
    g2, x2 = grad_out.detach().contiguous(), x.detach().contiguous()
    rows, n = g2.shape
    k = x2.shape[1]
    grad_w = torch.empty(n, k, dtype=torch.float32, device=g2.device)
    transposed = lib.cape_get_tuning(b"WGRAD_TRANSPOSE") == 1
    workspace = _workspace((n + 2 * k) * rows, g2.device) if transposed else None
    with torch.cuda.device(g2.device):
        rc = lib.cape_linear_tf32x3_wgrad(_ptr(g2), _ptr(x2), _ptr(grad_w), None if workspace is None else _ptr(workspace),
                                          rows, n, k, _stream(g2.device))
    _lib.check(rc, "cape_linear_tf32x3_wgrad")
    return grad_w


class _LinearTF32x3(torch.autograd.Function):
    """Training form: forward, the input gradient (g . W, the same K-major GEMM with W^T as the weight) and the weight
    gradient (g^T . x with both operands read in place as MN-major tiles, reduction split over the SMs) all run on the 3xTF32
    kernel."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return linear_tf32x3(x, weight, bias)

    @staticmethod
    def backward(ctx, grad_out):
        x, weight = ctx.saved_tensors
        g = grad_out.contiguous()
        grad_x = grad_w = grad_b = None
        if ctx.needs_input_grad[0]:
            wt = _weight_t(weight)
            if supported(g, wt):
                grad_x = linear_tf32x3(g, wt, None)
            else:
                grad_x = g.matmul(weight)
        if ctx.needs_input_grad[1]:
            g2, x2 = g.reshape(-1, g.shape[-1]), x.reshape(-1, x.shape[-1])
            if wgrad_supported(g2, x2):
                grad_w = linear_tf32x3_wgrad(g2, x2)
            else:
                grad_w = g2.t().matmul(x2)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            grad_b = g.reshape(-1, g.shape[-1]).sum(0)
        return grad_x, grad_w, grad_b


def linear(module: torch.nn.Linear, x: torch.Tensor, relu: bool = False) -> torch.Tensor:
    """``module(x)`` (optionally followed by ReLU), routed to the tensor-core kernel when the mode allows it."""
    if _MODE == "tf32x3" and supported(x, module.weight) and x.numel() // x.shape[-1] >= 128 \
            and not torch.is_autocast_enabled():          # under AMP nn.Linear runs in fp16 / bf16: leave that to autocast
        if not torch.is_grad_enabled():
            return linear_tf32x3(x, module.weight, module.bias, relu)
        y = _LinearTF32x3.apply(x, module.weight, module.bias)
        return F.relu(y) if relu else y
    y = module(x)
    return F.relu(y) if relu else y
