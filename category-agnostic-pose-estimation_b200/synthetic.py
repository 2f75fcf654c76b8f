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


# ---- synthetic MP-100-shaped episodes (BASELINE.json configs 1, 3, 4, 5) ---------------------------
def tokenize_keypoints(kpts_norm, num_bins: int = 44, seq_len: int = 200) -> dict:
    """Targets of one query exactly as the reference's dataset builds them (``datasets/mp100_cape.py:625-832`` with
    ``DiscreteTokenizerV2(num_bins, seq_len)``, ``datasets/discrete_tokenizer.py``): 4 index sequences + 4 deltas for the
    bilinear token embedding, ``target_seq``, ``token_labels`` (0 coord, 2 eos, -1 pad), ``mask``, ``visibility_mask``.
    ``kpts_norm``: (K, 2) float64 array of (x, y) in [0, 1]; all keypoints visible."""
    import numpy as np
    k = int(kpts_norm.shape[0])
    vocab = num_bins * num_bins
    bos, pad = vocab, vocab + 3
    q = np.clip(np.asarray(kpts_norm, dtype=np.float64) * (num_bins - 1), 0, num_bins - 1)
    fl = np.clip(np.floor(q), 0, num_bins - 1).astype(np.int64)
    ce = np.clip(np.ceil(q), 0, num_bins - 1).astype(np.int64)

    def seq(ix, iy):
        body = (ix * num_bins + iy).tolist()
        if 1 + len(body) + 1 > seq_len:                       # the tokenizer drops a polygon that does not fit
            body = []
        out = [bos] + body
        return torch.tensor(out + [pad] * (seq_len - len(out)), dtype=torch.long)

    def padded(values, fill, dtype):
        values = list(values)
        return torch.tensor(np.array(values + [fill] * (seq_len - len(values))), dtype=dtype)

    labels = [0] * k + [2]                                    # coords then <eos> (the last <sep> becomes <eos>)
    target = [list(map(float, p)) for p in np.asarray(kpts_norm, dtype=np.float64)] + [[0.0, 0.0]]
    mask = torch.zeros(seq_len, dtype=torch.bool)
    mask[:len(labels)] = True
    vis = torch.zeros(seq_len, dtype=torch.bool)
    vis[:k + 1] = True                                        # every keypoint visible + the first <eos>
    dx = [0.0] + (q[:, 0] - np.floor(q[:, 0])).tolist()
    dy = [0.0] + (q[:, 1] - np.floor(q[:, 1])).tolist()
    delta_x1, delta_y1 = padded(dx, 0, torch.float32), padded(dy, 0, torch.float32)
    poly_labels = torch.full((seq_len,), -1, dtype=torch.long)
    poly_labels[:min(k, seq_len)] = 0
    return {"seq11": seq(fl[:, 0], fl[:, 1]), "seq21": seq(ce[:, 0], fl[:, 1]), "seq12": seq(fl[:, 0], ce[:, 1]),
            "seq22": seq(ce[:, 0], ce[:, 1]), "target_seq": padded(target, [0.0, 0.0], torch.float32),
            "token_labels": padded(labels, -1, torch.long), "mask": mask, "visibility_mask": vis,
            "target_polygon_labels": poly_labels, "delta_x1": delta_x1, "delta_x2": 1 - delta_x1,
            "delta_y1": delta_y1, "delta_y2": 1 - delta_y1}


def make_episode_batch(num_episodes: int, queries_per_episode: int = 2, num_keypoints: int = 17, shots: int = 1,
                       image_size: int = 512, seed: int = 0, num_bins: int = 44, seq_len: int = 200,
                       with_images: bool = True) -> dict:
    """One collated batch of synthetic MP-100-shaped episodes, laid out like ``episodic_collate_fn``'s output
    (``datasets/episodic_sampler.py:353-480``): every tensor has first dimension episodes x queries; the support of an
    episode is the mean of its ``shots`` support draws (:439-442) repeated once per query; mask convention True = ignore
    (:280-284), all keypoints valid; chain skeleton.  Query images U(0, 1); query keypoints = support + N(0, 0.05)."""
    import numpy as np
    rng = np.random.RandomState(seed)
    n = num_episodes * queries_per_episode
    support = rng.uniform(0.05, 0.95, size=(num_episodes, shots, num_keypoints, 2)).mean(1)      # (B, K, 2)
    support_coords = torch.from_numpy(np.repeat(support, queries_per_episode, axis=0)).float()
    support_masks = torch.zeros(n, num_keypoints, dtype=torch.bool)
    skeleton = [[i, i + 1] for i in range(num_keypoints - 1)]
    targets = []
    for e in range(num_episodes):
        for _ in range(queries_per_episode):
            kp = np.clip(support[e] + rng.normal(0, 0.05, size=(num_keypoints, 2)), 0.0, 1.0)
            targets.append(tokenize_keypoints(kp, num_bins, seq_len))
    batch = {"support_images": None, "support_coords": support_coords, "support_masks": support_masks,
             "support_skeletons": [skeleton for _ in range(n)],
             "query_targets": {k: torch.stack([t[k] for t in targets]) for k in targets[0]},
             "category_ids": torch.arange(num_episodes).repeat_interleave(queries_per_episode)}
    if with_images:
        g = torch.Generator().manual_seed(seed)
        batch["query_images"] = torch.rand(n, 3, image_size, image_size, generator=g)
    return batch
