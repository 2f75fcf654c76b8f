"""Generate tests/golden/*.npz by running the REAL reference.  Build-container only.

    python -m oracle.make_golden            (needs /root/reference; the GPU box never runs this)

The reference's own tests hold no vectors for MSDeformAttn, so the parity pin is
the reference's own outputs on committed inputs: this script imports
``/root/reference/models/deformable_transformer.py`` by file path (the ``models`` package itself
needs pycocotools, which is absent) and records, for every case, the inputs, the forward output of
``ms_deform_attn_core_pytorch`` (:115-141) and the autograd gradients, both from an fp32 run (what a
user of the reference gets) and from an fp64 run (tight target for the closed-form oracles).
One extra case records ``MSDeformAttn.forward`` (:76-114) with seeded random weights.
"""
from __future__ import annotations

import importlib.util
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_ROOT = "/root/reference"
OUT_DIR = os.path.join(REPO, "tests", "golden")


def load_reference():
    sys.path.insert(0, REF_ROOT)
    spec = importlib.util.spec_from_file_location(
        "ref_deformable_transformer", os.path.join(REF_ROOT, "models", "deformable_transformer.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_synthetic():
    path = os.path.join(REPO, "category-agnostic-pose-estimation_b200", "synthetic.py")
    spec = importlib.util.spec_from_file_location("cape_synthetic", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_reference_v2():
    """models.deformable_transformer_v2 needs the ``models`` package, whose __init__ pulls in pycocotools and timm
    (absent here): stub them."""
    import types
    sys.path.insert(0, REF_ROOT)
    for name in ("pycocotools", "pycocotools.coco", "pycocotools.mask"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.COCO = object
            sys.modules[name] = m
    if "timm" not in sys.modules:
        timm = types.ModuleType("timm")
        layers = types.ModuleType("timm.layers")

        class _Identity(torch.nn.Module):
            def __init__(self, *a, **k):
                super().__init__()

            def forward(self, x):
                return x
        layers.DropPath = _Identity
        layers.Mlp = _Identity
        timm.layers = layers
        sys.modules["timm"] = timm
        sys.modules["timm.layers"] = layers
    for name in ("albumentations", "matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    import importlib
    return (importlib.import_module("models.deformable_transformer"),
            importlib.import_module("models.deformable_transformer_v2"),
            importlib.import_module("models.kv_cache"))


def record_layer_cases(syn):
    """Encoder stack (deformable_transformer.py:155-291) and decoder layer v1 (deformable_transformer_v2.py:262-370),
    teacher-forced and token-by-token with the reference's KVCache, dropout 0, seeded weights."""
    dt, v2, kvc = load_reference_v2()
    torch.manual_seed(4321)
    d_model, d_ffn, n_levels, n_heads, n_points = 64, 96, 4, 2, 4
    shapes = ((6, 8), (3, 4), (2, 2), (1, 2))
    s = sum(h * w for h, w in shapes)
    shapes_t = torch.tensor(shapes, dtype=torch.int64)
    starts_t = torch.tensor(syn.level_start_index(shapes), dtype=torch.int64)
    n = 2

    def randomise(module):
        with torch.no_grad():
            for p in module.parameters():
                p.add_(torch.randn_like(p) * 0.05)

    # ---- encoder, 2 layers
    enc = dt.DeformableTransformerEncoder(
        dt.DeformableTransformerEncoderLayer(d_model, d_ffn, 0.0, "relu", n_levels, n_heads, n_points), 2)
    randomise(enc)
    src = torch.randn(n, s, d_model)
    pos = torch.randn(n, s, d_model) * 0.5
    valid = torch.ones(n, n_levels, 2)
    valid[1] = torch.rand(n_levels, 2) * 0.3 + 0.7
    x = src.clone().requires_grad_(True)
    out = enc(x, shapes_t, starts_t, valid, pos, None)
    gout = torch.randn_like(out)
    gsrc, = torch.autograd.grad(out, x, gout)
    rec = {"src": src.numpy(), "pos": pos.numpy(), "valid_ratios": valid.numpy(), "spatial_shapes": shapes_t.numpy(),
           "level_start_index": starts_t.numpy(), "out": out.detach().numpy(), "grad_output": gout.numpy(),
           "grad_src": gsrc.numpy(), "d_model": d_model, "d_ffn": d_ffn, "n_levels": n_levels, "n_heads": n_heads,
           "n_points": n_points, "reference_points": enc.get_reference_points(shapes_t, valid, "cpu").numpy()}
    for k, v in enc.state_dict().items():
        rec["param." + k] = v.numpy()
    path = os.path.join(OUT_DIR, "encoder_stack.npz")
    np.savez_compressed(path, **rec)
    print(f"encoder_stack: {os.path.getsize(path) / 1024:.1f} KiB")

    # ---- decoder layer v1
    layer = v2.TransformerDecoderLayer(d_model, d_ffn, 0.0, "relu", n_levels, n_heads, n_points)
    randomise(layer)
    layer.eval()
    t_len, n_sup = 7, 5
    tgt = torch.randn(n, t_len, d_model)
    qpos = torch.randn(n, t_len, d_model) * 0.5
    refp = torch.rand(n, t_len, n_levels, 2)
    memory = torch.randn(n, s, d_model)
    sup = torch.randn(n, n_sup, d_model)
    sup_mask = torch.zeros(n, n_sup, dtype=torch.bool)
    sup_mask[1, -2:] = True
    causal = torch.triu(torch.full((t_len, t_len), float("-inf")), diagonal=1)
    t = tgt.clone().requires_grad_(True)
    mem = memory.clone().requires_grad_(True)
    out, _ = layer(t, qpos, refp, mem, shapes_t, starts_t, None, causal, support_features=sup, support_mask=sup_mask)
    gout = torch.randn_like(out)
    g_t, g_mem = torch.autograd.grad(out, (t, mem), gout)
    rec = {"tgt": tgt.numpy(), "query_pos": qpos.numpy(), "reference_points": refp.numpy(), "memory": memory.numpy(),
           "support_features": sup.numpy(), "support_mask": sup_mask.numpy(), "causal_mask": causal.numpy(),
           "spatial_shapes": shapes_t.numpy(), "level_start_index": starts_t.numpy(),
           "out_teacher_forced": out.detach().numpy(), "grad_output": gout.numpy(), "grad_tgt": g_t.numpy(),
           "grad_memory": g_mem.numpy(), "d_model": d_model, "d_ffn": d_ffn, "n_levels": n_levels, "n_heads": n_heads,
           "n_points": n_points}
    # token by token with the reference's own caches (what forward_inference drives, roomformer_v2.py:481-598)
    layer.kv_cache = kvc.KVCache(n, t_len, d_model, torch.float32)
    layer.cross_attn.cache = kvc.VCache(n, s, n_heads, d_model // n_heads, torch.float32)
    steps = []
    with torch.no_grad():
        for i in range(t_len):
            o, _ = layer(tgt[:, i:i + 1], qpos[:, i:i + 1], refp[:, i:i + 1], memory, shapes_t, starts_t, None,
                         torch.zeros(1, i + 1), input_pos=torch.tensor([i]), support_features=sup, support_mask=sup_mask)
            steps.append(o)
    rec["out_incremental"] = torch.cat(steps, 1).numpy()
    layer.kv_cache = None
    del layer.cross_attn.cache
    for k, v in layer.state_dict().items():
        rec["param." + k] = v.numpy()
    path = os.path.join(OUT_DIR, "decoder_layer.npz")
    np.savez_compressed(path, **rec)
    print(f"decoder_layer: {os.path.getsize(path) / 1024:.1f} KiB; teacher-forced vs incremental max diff "
          f"{float((out.detach() - torch.cat(steps, 1)).abs().max()):.2e}")


def record_variant_cases(syn):
    """Decoder V4's inline sampler (deformable_transformer_v2.py:661-687) and MSDeformablePoints
    (deformable_points.py:31-130), run from the reference's own code."""
    import types
    import importlib
    dt, v2, _ = load_reference_v2()
    dp = importlib.import_module("models.deformable_points")
    torch.manual_seed(777)
    # ---- V4 sampler: call the reference method on a stand-in object that carries the attributes it reads
    d_model, n_heads, n_levels, n_points = 64, 2, 3, 4
    shapes = ((6, 8), (3, 4), (2, 2))
    s = sum(h * w for h, w in shapes)
    shapes_t = torch.tensor(shapes, dtype=torch.int64)
    starts_t = torch.tensor(syn.level_start_index(shapes), dtype=torch.int64)
    holder = types.SimpleNamespace(
        sampling_offsets=torch.nn.Linear(d_model, n_heads * n_levels * n_points * 2),
        attention_weights=torch.nn.Linear(d_model, n_heads * n_levels * n_points),
        source_proj=torch.nn.Linear(d_model, d_model), n_heads=n_heads, n_levels=n_levels, n_points=n_points,
        d_model=d_model)
    with torch.no_grad():
        holder.sampling_offsets.weight.mul_(3.0)          # spread the samples over the maps (incl. out of bounds)
        holder.sampling_offsets.bias.add_(torch.rand_like(holder.sampling_offsets.bias) * 4)
    n, lq = 2, 7
    query = torch.randn(n, lq, d_model, requires_grad=True)
    src = torch.randn(n, s, d_model, requires_grad=True)
    out = v2.TransformerDecoderLayerV4._sample_reference_points(holder, query, src, shapes_t, starts_t)
    gout = torch.randn_like(out)
    mods = [holder.sampling_offsets, holder.attention_weights, holder.source_proj]
    params = [p for m_ in mods for p in m_.parameters()]
    grads = torch.autograd.grad(out, [query, src] + params, gout)
    rec = {"query": query.detach().numpy(), "src": src.detach().numpy(), "spatial_shapes": shapes_t.numpy(),
           "level_start_index": starts_t.numpy(), "out": out.detach().numpy(), "grad_output": gout.numpy(),
           "grad_query": grads[0].numpy(), "grad_src": grads[1].numpy(), "d_model": d_model, "n_heads": n_heads,
           "n_levels": n_levels, "n_points": n_points}
    names = ["sampling_offsets", "attention_weights", "source_proj"]
    gi = 2
    for name, m_ in zip(names, mods):
        for pn, p in m_.named_parameters():
            rec[f"param.{name}.{pn}"] = p.detach().numpy()
            rec[f"grad_param.{name}.{pn}"] = grads[gi].numpy()
            gi += 1
    path = os.path.join(OUT_DIR, "v4_sampler.npz")
    np.savez_compressed(path, **rec)
    print(f"v4_sampler: {os.path.getsize(path) / 1024:.1f} KiB")

    # ---- MSDeformablePoints
    embed, n_levels, n_heads = 64, 4, 2
    shapes = ((32, 16), (16, 8), (8, 4), (4, 2))
    s = sum(h * w for h, w in shapes)
    shapes_l = [list(hw) for hw in shapes]
    for tag, factor in (("clamp", -1), ("tanh", 2.0)):
        mod = dp.MSDeformablePoints(embed, n_levels, n_heads, offset_range_factor=factor)
        with torch.no_grad():
            for p in mod.parameters():
                p.add_(torch.randn_like(p) * 0.3)
        x = torch.randn(2, s, embed, requires_grad=True)
        out = mod(x, shapes_l, None)
        gout = torch.randn_like(out)
        params = dict(mod.named_parameters())
        grads = torch.autograd.grad(out, [x] + list(params.values()), gout)
        rec = {"x": x.detach().numpy(), "spatial_shapes": np.array(shapes, dtype=np.int64), "out": out.detach().numpy(),
               "grad_output": gout.numpy(), "grad_x": grads[0].numpy(), "embed_dim": embed, "n_levels": n_levels,
               "n_heads": n_heads, "offset_range_factor": factor}
        for (k, v), g_ in zip(params.items(), grads[1:]):
            rec["param." + k] = v.detach().numpy()
            rec["grad_param." + k] = g_.numpy()
        path = os.path.join(OUT_DIR, f"deformable_points_{tag}.npz")
        np.savez_compressed(path, **rec)
        print(f"deformable_points_{tag}: {os.path.getsize(path) / 1024:.1f} KiB, out {tuple(out.shape)}")


def build_reference_model(syn, seq_len, num_bins, feats, masks_full, seed):
    """The reference's RoomFormerV2 (roomformer_v2.py:149-270) around its v2 DeformableTransformer with CAPE's defaults
    (poly_refine, sine query positions, aux loss, decoder layer v1; train_cape_episodic.py:182-225), on a stub backbone
    that hands back pre-made 256-channel feature maps — the backbone and ``input_proj`` are not on the path under test."""
    import importlib
    load_reference_v2()
    rf = importlib.import_module("models.roomformer_v2")
    v2 = importlib.import_module("models.deformable_transformer_v2")
    misc = importlib.import_module("util.misc")
    pe = importlib.import_module("models.position_encoding")
    tok_mod = importlib.import_module("datasets.discrete_tokenizer")
    tokenizer = tok_mod.DiscreteTokenizerV2(num_bins=num_bins, seq_len=seq_len, add_cls=False)

    class StubBody(torch.nn.Module):
        strides = [8, 16, 32, 64]
        num_channels = [256, 256, 256, 256]

        def forward(self, tensor_list):
            out = {}
            for i, f in enumerate(feats):
                m = torch.nn.functional.interpolate(tensor_list.mask[None].float(), size=f.shape[-2:]).to(torch.bool)[0]
                out[str(i)] = misc.NestedTensor(f, m)
            return out

    class StubJoiner(torch.nn.Sequential):
        def __init__(self):
            super().__init__(StubBody(), pe.PositionEmbeddingSine(128, normalize=True))
            self.strides = StubBody.strides
            self.num_channels = StubBody.num_channels

        def forward(self, tensor_list):
            xs = self[0](tensor_list)
            out = [x for _, x in sorted(xs.items())]
            return out, [self[1](x).to(x.tensors.dtype) for x in out]

    transformer = v2.DeformableTransformer(
        d_model=256, nhead=8, num_encoder_layers=1, num_decoder_layers=2, dim_feedforward=64, dropout=0.0,
        activation="relu", poly_refine=True, return_intermediate_dec=True, aux_loss=True, num_feature_levels=4,
        dec_n_points=4, enc_n_points=4, query_pos_type="sine", vocab_size=len(tokenizer), seq_len=seq_len,
        pre_decoder_pos_embed=False, learnable_dec_pe=False, dec_attn_concat_src=False, dec_qkv_proj=True,
        dec_layer_type="v1", pad_idx=tokenizer.pad)
    model = rf.RoomFormerV2(StubJoiner(), transformer, num_classes=3, num_queries=seq_len, num_polys=1,
                            num_feature_levels=4, aux_loss=True, with_poly_refine=True, seq_len=seq_len,
                            tokenizer=tokenizer)
    model.input_proj = torch.nn.ModuleList([torch.nn.Identity() for _ in range(4)])
    checksum = syn.fill_parameters_(model.transformer, seed)
    with torch.no_grad():
        model.query_embed.weight.copy_(torch.from_numpy(syn.seeded_array("query_embed.weight", (seq_len, 2), seed, -2, 2)))
        # class logits: make all three token types win somewhere so the generated sequences are ragged
        for i, head in enumerate(model.class_embed):
            head.weight.mul_(6.0)
            head.bias.copy_(torch.tensor([0.8, 0.0, -0.3]))
    model.eval()
    return model, tokenizer, checksum


def record_model_cases(syn):
    """DeformableTransformer.forward (deformable_transformer_v2.py:177-254) + TransformerDecoder.forward (:1024-1131)
    teacher-forced with gradients, and RoomFormerV2.forward_inference's autoregressive loop (roomformer_v2.py:381-676)
    with the KV cache, both from the reference's own classes.  Weights / feature maps come from synthetic.seeded_array
    (keyed by name), so the fixture stores only masks, positional encodings, outputs and a weight checksum."""
    import importlib
    seed, seq_len, num_bins, n_sup = 31, 20, 6, 5
    sizes = ((64, 96), (48, 80))                              # second image is padded by the nested-tensor batching
    level_shapes = ((8, 12), (4, 6), (2, 3), (1, 2))
    n = len(sizes)
    feats = [torch.from_numpy(syn.seeded_array(f"feat{i}", (n, 256, h, w), seed)) for i, (h, w) in enumerate(level_shapes)]
    model, tokenizer, checksum = build_reference_model(syn, seq_len, num_bins, feats, None, seed)
    misc = importlib.import_module("util.misc")
    images = [torch.zeros(3, h, w) for h, w in sizes]
    samples = misc.nested_tensor_from_tensor_list(images)
    sup = torch.from_numpy(syn.seeded_array("support_features", (n, n_sup, 256), seed))
    sup_mask = torch.zeros(n, n_sup, dtype=torch.bool)
    sup_mask[1, -2:] = True
    dec = model.transformer.decoder

    # capture what the transformer is fed (masks and positional encodings come from the stub backbone)
    captured = {}
    orig_forward = model.transformer.forward

    def spy(srcs, masks, pos_embeds, *a, **k):
        captured.setdefault("masks", [m.clone() for m in masks])
        captured.setdefault("pos", [p.clone() for p in pos_embeds])
        return orig_forward(srcs, masks, pos_embeds, *a, **k)
    model.transformer.forward = spy

    # ---- teacher-forced forward + backward (RoomFormerV2.forward, :283-361)
    vocab_coords = num_bins * num_bins
    rng = np.random.default_rng(seed)
    seqs = rng.integers(0, vocab_coords, size=(4, n, seq_len))
    seqs[:, :, 0] = tokenizer.bos
    seqs[:, 0, 7] = tokenizer.sep
    seqs[:, 0, 15:] = tokenizer.pad
    seqs[:, 0, 14] = tokenizer.eos
    seqs[:, 1, -1] = tokenizer.eos
    dx = rng.random((n, seq_len)).astype(np.float32)
    dy = rng.random((n, seq_len)).astype(np.float32)
    seq_kwargs = {"seq11": torch.from_numpy(seqs[0]), "seq12": torch.from_numpy(seqs[1]),
                  "seq21": torch.from_numpy(seqs[2]), "seq22": torch.from_numpy(seqs[3]),
                  "delta_x1": torch.from_numpy(dx), "delta_x2": torch.from_numpy(1 - dx),
                  "delta_y1": torch.from_numpy(dy), "delta_y2": torch.from_numpy(1 - dy)}
    dec.support_features, dec.support_mask = sup, sup_mask             # what CAPEModel.forward does (cape_model.py:123-126)
    for f in feats:
        f.requires_grad_(True)
    out = model(samples, seq_kwargs=seq_kwargs)
    logits = torch.stack([a["pred_logits"] for a in out["aux_outputs"]] + [out["pred_logits"]])
    coords = torch.stack([a["pred_coords"] for a in out["aux_outputs"]] + [out["pred_coords"]])
    g_logits = torch.from_numpy(syn.seeded_array("grad_logits", logits.shape, seed))
    g_coords = torch.from_numpy(syn.seeded_array("grad_coords", coords.shape, seed))
    loss = (logits * g_logits).sum() + (coords * g_coords).sum()
    grad_names = ["decoder.token_embed.weight", "level_embed", "decoder.pos_trans.weight",
                  "encoder.layers.0.self_attn.value_proj.weight", "encoder.layers.0.self_attn.sampling_offsets.bias",
                  "decoder.layers.0.cross_attn.sampling_offsets.weight", "decoder.layers.1.cross_attn.attention_weights.bias",
                  "decoder.layers.0.attn_k.weight", "decoder.layers.1.support_attn.in_proj_weight",
                  "decoder.coords_embed.0.layers.2.weight", "decoder.class_embed.1.weight", "decoder.layers.1.linear2.weight"]
    params = dict(model.transformer.named_parameters())
    grads = torch.autograd.grad(loss, [params[k] for k in grad_names] + [model.query_embed.weight] + feats)
    rec = {"seed": seed, "seq_len": seq_len, "num_bins": num_bins, "n_sup": n_sup, "weight_checksum": checksum,
           "level_shapes": np.array(level_shapes, dtype=np.int64), "support_mask": sup_mask.numpy(),
           "pred_logits": logits.detach().numpy(), "pred_coords": coords.detach().numpy(),
           "grad_query_embed": grads[len(grad_names)].numpy()}
    for k in ("seq11", "seq12", "seq21", "seq22", "delta_x1", "delta_x2", "delta_y1", "delta_y2"):
        rec[k] = seq_kwargs[k].numpy()
    for i in range(4):
        rec[f"mask{i}"] = captured["masks"][i].numpy()
        rec[f"pos{i}"] = captured["pos"][i].numpy()
        rec[f"grad_feat{i}"] = grads[len(grad_names) + 1 + i].numpy()
    for k, g_ in zip(grad_names, grads):
        rec["grad_param." + k] = g_.numpy()[:8]               # first 8 rows: enough to pin the gradient, keeps the file small
    rec["state_dict_keys"] = np.array(sorted(model.transformer.state_dict().keys()))
    for f in feats:
        f.requires_grad_(False)

    # ---- autoregressive generation with the KV cache (RoomFormerV2.forward_inference, :381-676)
    with torch.no_grad():
        gen = model.forward_inference(samples, use_cache=True)
    dec.support_features = dec.support_mask = None
    steps = gen["pred_logits"].shape[1]
    kind = np.full((n, steps), -1, dtype=np.int64)             # gen_out: [x, y] -> 0, sep -> 2, everything else -1
    xy = np.zeros((n, steps, 2), dtype=np.float32)
    for j, row in enumerate(gen["gen_out"]):
        assert len(row) == steps
        for t, item in enumerate(row):
            if isinstance(item, list):
                kind[j, t] = 0
                xy[j, t] = item
            else:
                kind[j, t] = int(item)
    rec.update(gen_logits=gen["pred_logits"].numpy(), gen_coords=gen["pred_coords"].numpy(), gen_kind=kind, gen_xy=xy)
    # second run with the class bias tilted towards <eos>: every sequence finishes early, so the loop stops before max_len
    with torch.no_grad():
        for head in model.class_embed:
            head.bias.copy_(torch.tensor([0.0, 0.0, 7.0]))
        dec.support_features, dec.support_mask = sup, sup_mask
        gen2 = model.forward_inference(samples, use_cache=True)
        dec.support_features = dec.support_mask = None
    rec.update(gen2_logits=gen2["pred_logits"].numpy(), gen2_coords=gen2["pred_coords"].numpy(),
               gen2_bias=np.array([0.0, 0.0, 7.0], dtype=np.float32))
    print("second run:", gen2["pred_logits"].shape[1], "steps", gen2["pred_logits"].argmax(-1).tolist())
    path = os.path.join(OUT_DIR, "transformer_model.npz")
    np.savez_compressed(path, **rec)
    print(f"transformer_model: {os.path.getsize(path) / 1024:.1f} KiB; generated {steps} steps; token types per sample:",
          gen["pred_logits"].argmax(-1).tolist())


def run_reference(ref, value, shapes, loc, attn, gout, dtype):
    v = value.to(dtype).clone().requires_grad_(True)
    l = loc.to(dtype).clone().requires_grad_(True)
    a = attn.to(dtype).clone().requires_grad_(True)
    out = ref.ms_deform_attn_core_pytorch(v, shapes, l, a)
    gv, gl, ga = torch.autograd.grad(out, (v, l, a), gout.to(dtype))
    return out.detach(), gv, gl, ga


def record_core_case(ref, name, inp):
    shapes = inp["spatial_shapes"]
    rec = {
        "value": inp["value"].numpy(), "spatial_shapes": shapes.numpy(),
        "level_start_index": inp["level_start_index"].numpy(),
        "sampling_locations": inp["sampling_locations"].numpy(),
        "attention_weights": inp["attention_weights"].numpy(),
        "grad_output": inp["grad_output"].numpy(),
    }
    for tag, dt in (("32", torch.float32), ("64", torch.float64)):
        out, gv, gl, ga = run_reference(ref, inp["value"], shapes, inp["sampling_locations"],
                                        inp["attention_weights"], inp["grad_output"], dt)
        rec["out" + tag] = out.numpy()
        rec["grad_value" + tag] = gv.numpy()
        rec["grad_loc" + tag] = gl.numpy()
        rec["grad_attn" + tag] = ga.numpy()
    path = os.path.join(OUT_DIR, name + ".npz")
    np.savez_compressed(path, **rec)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


def edge_case_inputs(syn):
    """Hand-placed locations: pixel centres, cell borders, the exact -1 / W pixel-coordinate ends,
    fully out of bounds, plus zero-weighted samples."""
    shapes = ((6, 8), (3, 4), (2, 2), (1, 2))
    base = syn.make_inputs(1, 24, shapes, dist="uniform", seed=11)
    loc = base["sampling_locations"]
    L = len(shapes)
    specials = []
    for h, w in shapes:
        specials.append([
            (0.5 / w, 0.5 / h),            # centre of pixel (0,0): x = 0 exactly
            ((w - 0.5) / w, (h - 0.5) / h),  # centre of the last pixel
            (1.0 / w, 1.0 / h),            # cell border: x = 0.5
            (0.0, 0.0),                    # image corner: x = -0.5
            (1.0, 1.0),                    # x = W - 0.5
            (-0.5 / w, -0.5 / h),          # x = -1 exactly (weight 0, gradient not 0)
            (1.0 + 0.5 / w, 1.0 + 0.5 / h),  # x = W exactly
            (-1.0, 0.3), (0.4, 2.0), (-3.0, -3.0), (5.0, 5.0),   # far outside
            (0.5, -0.25 / h), (-0.25 / w, 0.5), (0.5, 1.0 + 0.25 / h), (1.0 + 0.25 / w, 0.5),
            (0.5, 0.5),
        ])
    for q in range(16):
        for l in range(L):
            x, y = specials[l][q]
            loc[0, q, :, l, :, 0] = x
            loc[0, q, :, l, :, 1] = y
    # queries 16..19: first two levels weighted zero, 20..23: only point 0 carries weight
    attn = base["attention_weights"]
    attn[0, 16:20, :, :2] = 0.0
    attn[0, 20:24, :, :, 1:] = 0.0
    base["sampling_locations"] = loc.contiguous()
    base["attention_weights"] = attn.contiguous()
    return base


def record_module_case(ref, syn):
    torch.manual_seed(1234)
    d_model, n_levels, n_heads, n_points = 64, 4, 2, 4
    shapes = ((6, 8), (3, 4), (2, 2), (1, 2))
    s = sum(h * w for h, w in shapes)
    mod = ref.MSDeformAttn(d_model, n_levels, n_heads, n_points)
    with torch.no_grad():                     # the zero-initialised projections would hide bugs
        for p in mod.parameters():
            p.add_(torch.randn_like(p) * 0.1)
    n, lq = 2, 9
    query = torch.randn(n, lq, d_model)
    src = torch.randn(n, s, d_model)
    refpts = torch.rand(n, lq, n_levels, 2)
    mask = torch.zeros(n, s, dtype=torch.bool)
    mask[1, -5:] = True
    shapes_t = torch.tensor(shapes, dtype=torch.int64)
    starts_t = torch.tensor(syn.level_start_index(shapes), dtype=torch.int64)
    rec = {"query": query.numpy(), "input_flatten": src.numpy(), "reference_points": refpts.numpy(),
           "padding_mask": mask.numpy(), "spatial_shapes": shapes_t.numpy(),
           "level_start_index": starts_t.numpy(),
           "d_model": d_model, "n_levels": n_levels, "n_heads": n_heads, "n_points": n_points}
    for k, v in mod.state_dict().items():
        rec["param." + k] = v.numpy()
    q = query.clone().requires_grad_(True)
    x = src.clone().requires_grad_(True)
    out = mod(q, refpts, x, shapes_t, starts_t, mask)
    gout = torch.randn_like(out)
    params = list(mod.parameters())
    grads = torch.autograd.grad(out, [q, x] + params, gout)
    rec["out"] = out.detach().numpy()
    rec["grad_output"] = gout.numpy()
    rec["grad_query"] = grads[0].numpy()
    rec["grad_input_flatten"] = grads[1].numpy()
    for (k, _), gval in zip(mod.named_parameters(), grads[2:]):
        rec["grad_param." + k] = gval.numpy()
    # 4-d reference points branch (:106-108)
    ref4 = torch.cat([refpts, torch.rand(n, lq, n_levels, 2) * 0.3 + 0.05], -1)
    rec["reference_points4"] = ref4.numpy()
    rec["out_ref4"] = mod(query, ref4, src, shapes_t, starts_t, None).detach().numpy()
    path = os.path.join(OUT_DIR, "module_forward.npz")
    np.savez_compressed(path, **rec)
    print(f"module_forward: {os.path.getsize(path) / 1024:.1f} KiB")


def main():
    os.makedirs(OUT_DIR, exist_ok=True)
    ref = load_reference()
    syn = load_synthetic()
    torch.set_num_threads(1)      # fixed summation order inside ATen
    only = set(sys.argv[1:])      # e.g. `python -m oracle.make_golden model` regenerates one group
    if only:
        groups = {"module": lambda: record_module_case(ref, syn), "layers": lambda: record_layer_cases(syn),
                  "variants": lambda: record_variant_cases(syn), "model": lambda: record_model_cases(syn)}
        for name in sorted(only):
            groups[name]()
        return
    record_core_case(ref, "core_uniform", syn.make_inputs(
        2, 37, ((8, 8), (4, 4), (2, 2), (1, 1)), dist="uniform", seed=1))
    enc_shapes = ((6, 8), (3, 4), (2, 2), (1, 2))
    record_core_case(ref, "core_encoder_nonsquare", syn.make_inputs(
        1, sum(h * w for h, w in enc_shapes), enc_shapes, dist="encoder", seed=2))
    record_core_case(ref, "core_edges", edge_case_inputs(syn))
    record_core_case(ref, "core_generic_dims", syn.make_inputs(
        2, 11, ((5, 7), (3, 2)), n_heads=3, head_dim=16, n_points=3, dist="uniform", seed=3))
    record_core_case(ref, "core_lq1", syn.make_inputs(
        3, 1, ((8, 8), (4, 4), (2, 2), (1, 1)), dist="uniform", seed=4))
    record_core_case(ref, "core_d64_p8", syn.make_inputs(
        1, 5, ((4, 4), (2, 2), (1, 1)), n_heads=2, head_dim=64, n_points=8, dist="uniform", seed=5))
    record_module_case(ref, syn)
    record_layer_cases(syn)
    record_variant_cases(syn)
    record_model_cases(syn)


if __name__ == "__main__":
    main()
