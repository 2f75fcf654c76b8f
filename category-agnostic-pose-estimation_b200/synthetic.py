"""Deterministic synthetic inputs for the MSDeformAttn hot path.

Shapes follow the CAPE configuration: 512² query images at ``--image_size 256`` give the pyramid
(64,64),(32,32),(16,16),(8,8) (S = 5440), d_model 256 = 8 heads × 32, 4 levels, 4 points
(``/root/reference/models/train_cape_episodic.py:168-188``, ``models/roomformer_v2.py:187-208``).

Two sampling-location distributions:

* ``"encoder"``  reference points are the pixel centres of the pyramid in flattened order
  (what ``DeformableTransformerEncoder.get_reference_points`` builds,
  ``models/deformable_transformer.py:248-271``), offsets are the module's initial bias pattern
  (8 head directions × (p+1) px, ``:63-69``) plus N(0,1) px noise — spatially coherent, a few % of
  corners out of bounds;
* ``"uniform"``  i.i.d. U(-0.1, 1.1) — the worst case for a gather.
"""
from __future__ import annotations

import math
from typing import Sequence

import torch

CAPE_PYRAMID = ((64, 64), (32, 32), (16, 16), (8, 8))
CAPE_PYRAMID_512 = ((32, 32), (16, 16), (8, 8), (4, 4))   # --image_size 512 (patch 2), S = 1360


def level_start_index(spatial_shapes: Sequence[Sequence[int]]) -> list:
    """Row offset of each level (``models/deformable_transformer_v2.py:204``)."""
    starts, acc = [], 0
    for h, w in spatial_shapes:
        starts.append(acc)
        acc += int(h) * int(w)
    return starts


def pyramid_reference_points(spatial_shapes, lq: int) -> torch.Tensor:
    """(Lq, 2) normalised (x, y) pixel centres of every level, flattened level 0 first, tiled to Lq."""
    pts = []
    for h, w in spatial_shapes:
        ys = (torch.arange(h, dtype=torch.float32) + 0.5) / h
        xs = (torch.arange(w, dtype=torch.float32) + 0.5) / w
        yy, xx = torch.meshgrid(ys, xs, indexing="ij")
        pts.append(torch.stack([xx.reshape(-1), yy.reshape(-1)], -1))
    pts = torch.cat(pts, 0)
    reps = (lq + pts.shape[0] - 1) // pts.shape[0]
    return pts.repeat(reps, 1)[:lq].contiguous()


def head_direction_offsets(n_heads: int, n_levels: int, n_points: int) -> torch.Tensor:
    """(M, L, P, 2) initial offset pattern of ``MSDeformAttn._reset_parameters`` (:63-69), in pixels."""
    thetas = torch.arange(n_heads, dtype=torch.float32) * (2.0 * math.pi / n_heads)
    grid = torch.stack([thetas.cos(), thetas.sin()], -1)
    grid = grid / grid.abs().max(-1, keepdim=True)[0]
    grid = grid.view(n_heads, 1, 1, 2).repeat(1, n_levels, n_points, 1)
    scale = torch.arange(1, n_points + 1, dtype=torch.float32).view(1, 1, n_points, 1)
    return grid * scale


def make_inputs(n: int, lq: int, spatial_shapes=CAPE_PYRAMID, n_heads: int = 8, head_dim: int = 32,
                n_points: int = 4, dist: str = "encoder", seed: int = 0,
                dtype: torch.dtype = torch.float32, device="cpu", with_grad_out: bool = True) -> dict:
    """Seeded inputs for one op call.  Generated on the CPU generator, then moved to ``device``.

    Returns a dict with ``value (N,S,M,D)``, ``spatial_shapes (L,2) int64``, ``level_start_index (L,)
    int64``, ``sampling_locations (N,Lq,M,L,P,2)``, ``attention_weights (N,Lq,M,L,P)`` and
    ``grad_output (N,Lq,M*D)``.
    """
    g = torch.Generator(device="cpu").manual_seed(int(seed))
    shapes = [(int(h), int(w)) for h, w in spatial_shapes]
    n_levels = len(shapes)
    s = sum(h * w for h, w in shapes)
    value = torch.randn(n, s, n_heads, head_dim, generator=g)
    logits = torch.randn(n, lq, n_heads, n_levels * n_points, generator=g)
    attn = torch.softmax(logits, -1).view(n, lq, n_heads, n_levels, n_points)
    wh = torch.tensor([[w, h] for h, w in shapes], dtype=torch.float32)          # (L, 2)
    if dist == "encoder":
        ref = pyramid_reference_points(shapes, lq)                               # (Lq, 2)
        off = head_direction_offsets(n_heads, n_levels, n_points)                # (M, L, P, 2) px
        noise = torch.randn(n, lq, n_heads, n_levels, n_points, 2, generator=g)
        loc = ref.view(1, lq, 1, 1, 1, 2) + (off.view(1, 1, n_heads, n_levels, n_points, 2) + noise) \
            / wh.view(1, 1, 1, n_levels, 1, 2)
    elif dist == "uniform":
        loc = torch.rand(n, lq, n_heads, n_levels, n_points, 2, generator=g) * 1.2 - 0.1
    else:
        raise ValueError(f"unknown location distribution {dist!r}")
    out = {
        "value": value.to(dtype).to(device),
        "spatial_shapes": torch.tensor(shapes, dtype=torch.int64, device=device),
        "level_start_index": torch.tensor(level_start_index(shapes), dtype=torch.int64, device=device),
        "sampling_locations": loc.contiguous().to(dtype).to(device),
        "attention_weights": attn.contiguous().to(dtype).to(device),
    }
    if with_grad_out:
        out["grad_output"] = torch.randn(n, lq, n_heads * head_dim, generator=g).to(dtype).to(device)
    return out


def algorithmic_bytes(n: int, lq: int, s: int, n_heads: int = 8, head_dim: int = 32, n_levels: int = 4,
                      n_points: int = 4, e_value: int = 4, e_aux: int = 4, e_grad: int = 4):
    """(A_fwd, A_bwd): each tensor touched once (BASELINE.md §3 / SURVEY.md §8d)."""
    c = n_heads * head_dim
    mlp = n_heads * n_levels * n_points
    a_fwd = e_value * n * s * c + e_aux * 3 * n * lq * mlp + e_value * n * lq * c
    a_bwd = (e_value * n * lq * c + e_value * n * s * c + e_aux * 3 * n * lq * mlp
             + e_grad * n * s * c + e_aux * 3 * n * lq * mlp)
    return a_fwd, a_bwd


# ---- name-keyed deterministic arrays (model-level fixtures) ------------------------------------------------------
# Model-level fixtures would need megabytes of weights; instead both sides (the reference model inside
# oracle/make_golden.py, the mirror inside the tests) are filled from the same recipe, keyed by parameter NAME, so the
# fixture only stores outputs plus a checksum of what the recipe produced.  Integer hashing only — no libm — so the
# values are identical on every machine.
def seeded_array(name: str, shape, seed: int = 0, low: float = -1.0, high: float = 1.0):
    """float32 array U(low, high), a pure function of (name, shape, seed)."""
    import zlib

    import numpy as np
    count = 1
    for s in shape:
        count *= int(s)
    key = np.uint64(zlib.crc32(name.encode()) ^ (int(seed) * 0x9E3779B1 & 0xFFFFFFFF))
    x = np.arange(count, dtype=np.uint64) + (key << np.uint64(32)) + np.uint64(0x9E3779B97F4A7C15)
    # splitmix64 finaliser
    x ^= x >> np.uint64(30)
    x *= np.uint64(0xBF58476D1CE4E5B9)
    x ^= x >> np.uint64(27)
    x *= np.uint64(0x94D049BB133111EB)
    x ^= x >> np.uint64(31)
    u = (x >> np.uint64(40)).astype(np.float64) / float(1 << 24)            # 24-bit uniform in [0, 1)
    return (low + (high - low) * u).astype(np.float32).reshape(tuple(int(s) for s in shape))


def fill_parameters_(module, seed: int = 0, gain: float = 1.0) -> float:
    """Overwrite every parameter of ``module`` from :func:`seeded_array`, keyed by its ``named_parameters`` name:
    matrices U(+-gain*sqrt(3/fan_in)), LayerNorm / GroupNorm weights 1 + U(+-0.1), other vectors U(+-0.1).
    Returns a checksum (sum of |w|) the fixtures record."""
    total = 0.0
    with torch.no_grad():
        for name, p in module.named_parameters():
            if p.dim() >= 2:
                fan_in = p.shape[1] if p.dim() == 2 else p[0].numel()
                a = gain * math.sqrt(3.0 / max(fan_in, 1))
                w = seeded_array(name, p.shape, seed, -a, a)
            elif "norm" in name and name.endswith("weight"):
                w = seeded_array(name, p.shape, seed, 0.9, 1.1)
            else:
                w = seeded_array(name, p.shape, seed, -0.1, 0.1)
            p.copy_(torch.from_numpy(w).to(p.device, p.dtype))
            total += float(abs(w.astype("float64")).sum())
    return total
