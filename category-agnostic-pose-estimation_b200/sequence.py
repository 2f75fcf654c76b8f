"""Sequence side of the decoder on the device.

    cape::seq_embed(table, seq11, seq12, seq21, seq22, delta_x1, delta_x2, delta_y1, delta_y2, padding_idx) -> (B, T, C)
        TransformerDecoder._seq_embed, /root/reference/models/deformable_transformer_v2.py:984-997 — four embedding
        lookups and the bilinear blend in one kernel; differentiable w.r.t. the table (the deltas and token ids are data).
    token_step(...)
        the per-sample bookkeeping of RoomFormerV2.forward_inference (/root/reference/models/roomformer_v2.py:548-597).

CUDA only, no fallback (the ops are registered for the CUDA dispatch key alone).
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass

import torch

from . import _lib
from .ops import _ptr, _stream


def _seq_args(table, seqs, deltas):
    if table.dim() != 2 or table.dtype != torch.float32:
        raise ValueError(f"embedding table must be fp32 (V, C), got {table.dtype} {tuple(table.shape)}")
    shape = seqs[0].shape
    for t in list(seqs) + list(deltas):
        if t.shape != shape:
            raise ValueError(f"token / delta tensors must share one shape, got {tuple(t.shape)} vs {tuple(shape)}")
    seqs = [s.to(torch.int64).contiguous() for s in seqs]
    deltas = [d.to(torch.float32).contiguous() for d in deltas]
    return table.contiguous(), seqs, deltas, shape


@torch.library.custom_op("cape::seq_embed", mutates_args=(), device_types="cuda")
def seq_embed(table: torch.Tensor, seq11: torch.Tensor, seq12: torch.Tensor, seq21: torch.Tensor, seq22: torch.Tensor,
              delta_x1: torch.Tensor, delta_x2: torch.Tensor, delta_y1: torch.Tensor, delta_y2: torch.Tensor,
              padding_idx: int = -1) -> torch.Tensor:
    lib = _lib.load()
    table, seqs, deltas, shape = _seq_args(table, (seq11, seq12, seq21, seq22), (delta_x1, delta_x2, delta_y1, delta_y2))
    v, c = table.shape
    out = torch.empty(tuple(shape) + (c,), dtype=torch.float32, device=table.device)
    with torch.cuda.device(table.device):
        rc = lib.cape_seq_embed_forward(_ptr(table), *[_ptr(s) for s in seqs], *[_ptr(d) for d in deltas], _ptr(out),
                                        seqs[0].numel(), c, v, _stream(table.device))
    _lib.check(rc, "cape_seq_embed_forward")
    return out


@seq_embed.register_fake
def _(table, seq11, seq12, seq21, seq22, delta_x1, delta_x2, delta_y1, delta_y2, padding_idx=-1):
    return table.new_empty(tuple(seq11.shape) + (table.shape[1],))


@torch.library.custom_op("cape::seq_embed_backward", mutates_args=(), device_types="cuda")
def seq_embed_backward(grad_out: torch.Tensor, table: torch.Tensor, seq11: torch.Tensor, seq12: torch.Tensor,
                       seq21: torch.Tensor, seq22: torch.Tensor, delta_x1: torch.Tensor, delta_x2: torch.Tensor,
                       delta_y1: torch.Tensor, delta_y2: torch.Tensor, padding_idx: int) -> torch.Tensor:
    lib = _lib.load()
    table, seqs, deltas, _ = _seq_args(table, (seq11, seq12, seq21, seq22), (delta_x1, delta_x2, delta_y1, delta_y2))
    v, c = table.shape
    g = grad_out.to(torch.float32).contiguous()
    grad_table = torch.empty_like(table)
    with torch.cuda.device(table.device):
        rc = lib.cape_seq_embed_backward(_ptr(g), *[_ptr(s) for s in seqs], *[_ptr(d) for d in deltas],
                                         _ptr(grad_table), seqs[0].numel(), c, v, padding_idx, 1, _stream(table.device))
    _lib.check(rc, "cape_seq_embed_backward")
    return grad_table


@seq_embed_backward.register_fake
def _(grad_out, table, seq11, seq12, seq21, seq22, delta_x1, delta_x2, delta_y1, delta_y2, padding_idx):
    return torch.empty_like(table)


def _se_setup(ctx, inputs, output):
    ctx.save_for_backward(*inputs[:9])
    ctx.padding_idx = inputs[9]


def _se_backward(ctx, grad_out):
    saved = ctx.saved_tensors
    return (torch.ops.cape.seq_embed_backward(grad_out, *saved, ctx.padding_idx),) + (None,) * 9


seq_embed.register_autograd(_se_backward, setup_context=_se_setup)


@dataclass
class TokenizerSpec:
    """Constants of the reference's ``DiscreteTokenizer`` (datasets/discrete_tokenizer.py:7-28) and ``TokenType``
    (datasets/token_types.py) that the generation loop needs."""
    num_bins: int
    seq_len: int
    add_cls: bool = False
    min_len: int = 6                       # roomformer_v2.py:456

    def __post_init__(self):
        v = self.num_bins * self.num_bins
        self.bos, self.eos, self.sep, self.pad = v, v + 1, v + 2, v + 3
        self.cls = v + 4 if self.add_cls else -1
        self.vocab_size = v + (5 if self.add_cls else 4)

    @classmethod
    def from_tokenizer(cls, tokenizer, min_len: int = 6):
        """From a reference tokenizer object (anything with num_bins / seq_len / add_cls)."""
        return cls(int(tokenizer.num_bins), int(tokenizer.seq_len), bool(getattr(tokenizer, "add_cls", False)), min_len)

    def as_struct(self) -> "_lib.Tokenizer":
        return _lib.Tokenizer(self.num_bins, self.min_len, self.bos, self.eos, self.sep, self.pad, self.cls, 0, 1, 2, 3)


class TokenState:
    """Device-resident state of one autoregressive generation (what the reference keeps in Python lists,
    roomformer_v2.py:445-480): the next step's ``_seq_embed`` inputs, the unfinished flags, and the per-step records."""

    def __init__(self, batch: int, spec: TokenizerSpec, n_classes: int, device):
        self.spec, self.batch, self.n_classes = spec, batch, n_classes
        dev = torch.device(device)
        t = spec.seq_len
        self.step = torch.zeros(1, dtype=torch.int64, device=dev)
        self.unfinished = torch.ones(batch, dtype=torch.int32, device=dev)
        self.finish_step = torch.full((batch,), -1, dtype=torch.int64, device=dev)
        self.seq = [torch.full((batch, 1), spec.bos, dtype=torch.int64, device=dev) for _ in range(4)]   # 11, 12, 21, 22
        self.delta = [torch.zeros(batch, 1, device=dev) for _ in range(4)]                                # x1, x2, y1, y2
        self.pred_logits = torch.zeros(batch, t, n_classes, device=dev)
        self.pred_coords = torch.zeros(batch, t, 2, device=dev)
        self.gen_kind = torch.full((batch, t), -1, dtype=torch.int32, device=dev)
        self.gen_xy = torch.zeros(batch, t, 2, device=dev)
        self._tok = spec.as_struct()
        self._st = _lib.TokenState(*[t_.data_ptr() for t_ in (
            self.unfinished, self.finish_step, *self.seq, *self.delta, self.pred_logits, self.pred_coords, self.gen_kind,
            self.gen_xy)], t)
        self.reset()

    def reset(self):
        """roomformer_v2.py:365-385 (``_prepare_sequences``): every sequence starts from <bos> with deltas (0, 1, 0, 1)."""
        self.step.zero_()
        self.unfinished.fill_(1)
        self.finish_step.fill_(-1)
        for s in self.seq:
            s.fill_(self.spec.bos)
        for d, v in zip(self.delta, (0.0, 1.0, 0.0, 1.0)):
            d.fill_(v)
        self.gen_kind.fill_(-1)
        for t in (self.pred_logits, self.pred_coords, self.gen_xy):
            t.zero_()

    def advance(self, cls_logits: torch.Tensor, reg: torch.Tensor):
        """Consume this step's head outputs ((B, 1, n_classes) and (B, 1, 2), fp32, contiguous) and produce the next
        step's inputs; increments the device step counter.  No host synchronisation."""
        if cls_logits.dtype != torch.float32 or reg.dtype != torch.float32 or not cls_logits.is_contiguous() \
                or not reg.is_contiguous() or cls_logits.numel() != self.batch * self.n_classes \
                or reg.numel() != self.batch * 2:
            raise ValueError("token_step expects contiguous fp32 (B, 1, n_classes) logits and (B, 1, 2) coordinates")
        lib = _lib.load()
        dev = self.step.device
        with torch.cuda.device(dev):
            rc = lib.cape_token_step(_ptr(cls_logits), _ptr(reg), _ptr(self.step), ctypes.byref(self._st),
                                     ctypes.byref(self._tok), self.batch, self.n_classes, _stream(dev))
        _lib.check(rc, "cape_token_step")
