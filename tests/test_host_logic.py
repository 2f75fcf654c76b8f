"""Host-side mirror of the reference interface: module layout, error behaviour, shape inference, sharding maths."""
import math
import os
import types

import numpy as np
import pytest
import torch

import cape_b200
from cape_b200 import synthetic
from tests.conftest import GOLDEN


def test_state_dict_layout_matches_reference_module():
    g = np.load(os.path.join(GOLDEN, "module_forward.npz"))
    ref = {k[len("param."):]: g[k].shape for k in g.files if k.startswith("param.")}
    mod = cape_b200.MSDeformAttn(int(g["d_model"]), int(g["n_levels"]), int(g["n_heads"]), int(g["n_points"]))
    mine = {k: tuple(v.shape) for k, v in mod.state_dict().items()}
    assert mine == {k: tuple(v) for k, v in ref.items()}
    assert [n for n, _ in mod.named_buffers()] == []          # no persistent buffers under new names
    # CAPE configuration: 230,272 parameters per module
    assert sum(p.numel() for p in cape_b200.MSDeformAttn(256, 4, 8, 4).parameters()) == 230272


def test_initialisation_follows_reference_reset_parameters():
    torch.manual_seed(0)
    m = cape_b200.MSDeformAttn(256, 4, 8, 4)
    assert float(m.sampling_offsets.weight.detach().abs().max()) == 0.0
    assert float(m.attention_weights.weight.abs().max()) == 0.0 and float(m.attention_weights.bias.abs().max()) == 0.0
    want = synthetic.head_direction_offsets(8, 4, 4).reshape(-1)
    assert torch.allclose(m.sampling_offsets.bias, want)
    assert float(m.value_proj.bias.abs().max()) == 0.0 and float(m.output_proj.bias.abs().max()) == 0.0
    bound = math.sqrt(6.0 / (256 + 256))
    assert float(m.value_proj.weight.abs().max()) <= bound + 1e-6


def test_constructor_and_forward_errors():
    with pytest.raises(ValueError, match="divisible"):
        cape_b200.MSDeformAttn(100, 4, 8, 4)
    m = cape_b200.MSDeformAttn(32, 2, 2, 2)
    q = torch.zeros(1, 3, 32)
    src = torch.zeros(1, 8, 32)
    shapes = torch.tensor([[2, 2], [2, 2]])
    starts = torch.tensor([0, 4])
    with pytest.raises(ValueError, match="2 or 4"):
        m(q, torch.zeros(1, 3, 2, 3), src, shapes, starts)
    with pytest.raises(AssertionError):
        m(q, torch.zeros(1, 3, 2, 2), torch.zeros(1, 9, 32), shapes, starts)


def test_cpu_tensors_are_rejected_not_computed():
    inp = synthetic.make_inputs(1, 2, ((2, 2),), n_heads=1, head_dim=4, n_points=1)
    with pytest.raises(NotImplementedError):
        cape_b200.ms_deform_attn_core_pytorch(inp["value"], inp["spatial_shapes"], inp["sampling_locations"],
                                              inp["attention_weights"])


def test_fake_tensor_shape_inference():
    v = torch.empty(3, 85, 8, 32, device="meta")
    loc = torch.empty(3, 7, 8, 4, 4, 2, device="meta")
    attn = torch.empty(3, 7, 8, 4, 4, device="meta")
    shapes = torch.empty(4, 2, dtype=torch.int64, device="meta")
    starts = torch.empty(4, dtype=torch.int64, device="meta")
    out = torch.ops.cape.ms_deform_attn(v, shapes, starts, loc, attn)
    assert out.shape == (3, 7, 256) and out.dtype == v.dtype
    gv, gl, ga = torch.ops.cape.ms_deform_attn_backward(out, v, shapes, starts, loc, attn)
    assert gv.shape == v.shape and gl.shape == loc.shape and ga.shape == attn.shape
    dec = torch.ops.cape.ms_deform_attn_decode(v, shapes, starts, torch.empty(3, 1, 4, 2, device="meta"),
                                               torch.empty(3, 1, 8, 4, 4, 2, device="meta"),
                                               torch.empty(3, 1, 8, 16, device="meta"))
    assert dec.shape == (3, 1, 256)


def test_level_start_index_helpers():
    shapes = torch.tensor(synthetic.CAPE_PYRAMID)
    assert cape_b200.level_start_index_from_shapes(shapes).tolist() == [0, 4096, 5120, 5376]
    assert synthetic.level_start_index(synthetic.CAPE_PYRAMID) == [0, 4096, 5120, 5376]
    assert synthetic.level_start_index(synthetic.CAPE_PYRAMID_512) == [0, 1024, 1280, 1344]


def test_algorithmic_bytes_match_baseline_table():
    # BASELINE.md §3 worked values (fp32)
    for (n, lq), (fwd_mb, bwd_mb) in {(2, 5440): (39.0, 66.8), (4, 5440): (78.0, 133.7), (20, 5440): (389.9, 668.5),
                                      (2, 1000): (16.3, 30.5), (20, 200): (121.7, 239.2)}.items():
        a_fwd, a_bwd = synthetic.algorithmic_bytes(n, lq, 5440)
        assert abs(a_fwd / 1e6 - fwd_mb) < 0.06 and abs(a_bwd / 1e6 - bwd_mb) < 0.06


def test_synthetic_inputs_are_deterministic_and_well_formed():
    a = synthetic.make_inputs(2, 50, seed=3)
    b = synthetic.make_inputs(2, 50, seed=3)
    for k in a:
        assert torch.equal(a[k], b[k])
    assert a["value"].shape == (2, 5440, 8, 32) and a["sampling_locations"].shape == (2, 50, 8, 4, 4, 2)
    assert torch.allclose(a["attention_weights"].flatten(3).sum(-1), torch.ones(2, 50, 8), atol=1e-5)
    loc = a["sampling_locations"]
    assert -1.5 < float(loc.min()) < 0.0 and 1.0 < float(loc.max()) < 2.5      # some corners out of bounds


def test_patch_reference_rebinds_the_seam_and_restores_it():
    fake = types.ModuleType("fake_models.deformable_transformer")
    original = lambda *a: "reference"
    fake.ms_deform_attn_core_pytorch = original
    fake.MSDeformAttn = object
    cape_b200.patch_reference(fake, swap_module_class=True)
    assert fake.ms_deform_attn_core_pytorch is cape_b200.ms_deform_attn_core_pytorch
    assert fake.MSDeformAttn is cape_b200.MSDeformAttn
    cape_b200.unpatch_reference(fake)
    assert fake.ms_deform_attn_core_pytorch is original and fake.MSDeformAttn is object


def test_value_cache_protocol_without_buffers():
    m = cape_b200.MSDeformAttn(32, 1, 1, 1)
    m.cache = cape_b200.ValueCache()
    assert "cache" not in "".join(m.state_dict().keys())
    v = torch.zeros(2, 4, 1, 32)
    assert m._cache_load(2, 4) is None           # nothing stored by this module yet
    m._cache_store(v)
    assert m._cache_load(2, 4) is v
    assert m._cache_load(3, 4) is None           # batch changed: do not trust the cache
    m.cache = cape_b200.ValueCache()             # _setup_caches attaches a fresh holder every forward_inference
    assert m._cache_load(2, 4) is None


def _params(g):
    return {k[len("param."):]: g[k] for k in g.files if k.startswith("param.")}


def test_layer_mirrors_have_the_reference_state_dict_layout():
    g = np.load(os.path.join(GOLDEN, "encoder_stack.npz"))
    kw = dict(d_model=int(g["d_model"]), d_ffn=int(g["d_ffn"]), dropout=0.0, n_levels=int(g["n_levels"]),
              n_heads=int(g["n_heads"]), n_points=int(g["n_points"]))
    enc = cape_b200.DeformableTransformerEncoder(cape_b200.DeformableTransformerEncoderLayer(**kw), 2)
    assert {k: tuple(v.shape) for k, v in enc.state_dict().items()} == {k: v.shape for k, v in _params(g).items()}
    enc.load_state_dict({k: torch.from_numpy(v) for k, v in _params(g).items()})
    # reference points are host logic: compare on the CPU
    ref = enc.get_reference_points(torch.from_numpy(g["spatial_shapes"]), torch.from_numpy(g["valid_ratios"]), "cpu")
    assert torch.allclose(ref, torch.from_numpy(g["reference_points"]), atol=1e-7)

    g = np.load(os.path.join(GOLDEN, "decoder_layer.npz"))
    layer = cape_b200.TransformerDecoderLayer(**kw)
    assert {k: tuple(v.shape) for k, v in layer.state_dict().items()} == {k: v.shape for k, v in _params(g).items()}
    layer.setup_caches(2, 7)
    assert set(layer.state_dict()) == set(_params(g))          # caches never enter the checkpoint


def test_kv_cache_update_returns_prefix():
    kv = cape_b200.KVCache(2, 5, 4)
    for i in range(3):
        k, v = kv.update(torch.tensor([i]), torch.full((2, 1, 4), float(i)), torch.full((2, 1, 4), float(-i)))
        assert k.shape == (2, i + 1, 4) and float(k[0, i, 0]) == i and float(v[1, i, 3]) == -i
    k, _ = kv.update(1, torch.full((2, 1, 4), 9.0), torch.zeros(2, 1, 4))      # python-int position
    assert k.shape == (2, 2, 4) and float(k[0, 1, 0]) == 9.0


def test_patch_reference_routes_the_real_reference_module_to_the_op():
    """With the actual reference checked out (build container only), rebinding the seam makes the reference's own
    MSDeformAttn.forward (models/deformable_transformer.py:76-114) dispatch to cape::ms_deform_attn: on CPU tensors
    that shows as the dispatcher's NotImplementedError (there is no CPU path), and unpatching restores eager."""
    ref_root = "/root/reference"
    path = os.path.join(ref_root, "models", "deformable_transformer.py")
    if not os.path.exists(path):
        pytest.skip("reference checkout not present (GPU box)")
    import importlib.util
    import sys
    sys.path.insert(0, ref_root)
    try:
        spec = importlib.util.spec_from_file_location("ref_dt_for_patch_test", path)
        ref = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref)
    finally:
        sys.path.remove(ref_root)
    torch.manual_seed(0)
    mod = ref.MSDeformAttn(64, 2, 2, 4)
    mine = cape_b200.MSDeformAttn(64, 2, 2, 4)
    assert list(mod.state_dict()) == list(mine.state_dict())                      # checkpoint-compatible
    mine.load_state_dict(mod.state_dict())
    q, src = torch.randn(1, 3, 64), torch.randn(1, 20, 64)
    refp = torch.rand(1, 3, 2, 2)
    shapes, starts = torch.tensor([[4, 4], [2, 2]]), torch.tensor([0, 16])
    eager = mod(q, refp, src, shapes, starts)
    cape_b200.patch_reference(ref)
    try:
        with pytest.raises(NotImplementedError):
            mod(q, refp, src, shapes, starts)
    finally:
        cape_b200.unpatch_reference(ref)
    assert torch.equal(mod(q, refp, src, shapes, starts), eager)


# ---- transformer-level mirrors and the generation state --------------------------
def _small_transformer(spec, **kw):
    args = dict(d_model=256, nhead=8, num_encoder_layers=1, num_decoder_layers=2, dim_feedforward=64, dropout=0.0,
                poly_refine=True, return_intermediate_dec=True, aux_loss=True, num_feature_levels=4,
                query_pos_type="sine", vocab_size=spec.vocab_size, seq_len=spec.seq_len, pad_idx=spec.pad)
    args.update(kw)
    return cape_b200.DeformableTransformer(**args)


def test_transformer_mirror_has_the_reference_state_dict_layout_and_seeded_weights():
    g = np.load(os.path.join(GOLDEN, "transformer_model.npz"))
    spec = cape_b200.TokenizerSpec(int(g["num_bins"]), int(g["seq_len"]))
    tr = _small_transformer(spec).attach_heads(*cape_b200.build_prediction_heads(256, 3, 2, True))
    assert sorted(tr.state_dict().keys()) == list(g["state_dict_keys"])
    assert synthetic.fill_parameters_(tr, int(g["seed"])) == pytest.approx(float(g["weight_checksum"]), rel=1e-12)
    # pos_embed is a frozen parameter unless learnable_dec_pe (deformable_transformer_v2.py:132)
    assert not tr.pos_embed.requires_grad and _small_transformer(spec, learnable_dec_pe=True).pos_embed.requires_grad


def test_transformer_mirror_rejects_configurations_the_reference_cannot_run_in_cape():
    spec = cape_b200.TokenizerSpec(6, 20)
    with pytest.raises(ValueError, match="v1"):
        _small_transformer(spec, dec_layer_type="v4")
    with pytest.raises(NotImplementedError):
        _small_transformer(spec, inject_cls_embed=True)
    tr = _small_transformer(spec, poly_refine=False, query_pos_type="none")
    assert tr.decoder.pos_trans is None
    with pytest.raises(NotImplementedError):                   # the one-graph generator covers the CAPE configuration only
        cape_b200.AutoregressiveGenerator(tr, spec, 1, "cpu")


def test_prediction_heads_follow_roomformer_initialisation():
    cls, coords = cape_b200.build_prediction_heads(256, 3, 6, with_poly_refine=True)
    assert len(cls) == len(coords) == 6 and cls[0] is not cls[1]                     # clones (roomformer_v2.py:231-233)
    assert torch.allclose(cls[0].bias, torch.full((3,), -math.log(99.0)))            # prior 0.01 (:219-221)
    assert float(coords[0].layers[-1].weight.abs().max()) == 0.0 and float(coords[0].layers[-1].bias.abs().max()) == 0.0
    cls, coords = cape_b200.build_prediction_heads(256, 3, 6, with_poly_refine=False)
    assert cls[0] is cls[5] and coords[0] is coords[5]                               # shared (:236-237)


def test_tokenizer_spec_matches_the_reference_vocabulary():
    spec = cape_b200.TokenizerSpec(num_bins=44, seq_len=200)                         # train_cape_episodic defaults
    assert (spec.bos, spec.eos, spec.sep, spec.pad, spec.vocab_size) == (1936, 1937, 1938, 1939, 1940)
    assert spec.min_len == 6 and spec.cls == -1
    with_cls = cape_b200.TokenizerSpec(num_bins=44, seq_len=200, add_cls=True)
    assert with_cls.cls == 1940 and with_cls.vocab_size == 1941
    ref_like = types.SimpleNamespace(num_bins=6, seq_len=20, add_cls=False)
    assert cape_b200.TokenizerSpec.from_tokenizer(ref_like).pad == 39


def test_sincos_table_and_query_pos_embedding_shapes():
    from cape_b200.transformer import sincos_position_table, inverse_sigmoid
    tab = sincos_position_table(256, 200)
    assert tab.shape == (200, 256) and np.allclose(tab[0, :128], 0.0) and np.allclose(tab[0, 128:], 1.0)
    pos = cape_b200.TransformerDecoder.get_query_pos_embed(torch.rand(2, 5, 2))
    assert pos.shape == (2, 5, 256)
    x = torch.tensor([0.0, 0.25, 1.0])
    assert torch.allclose(inverse_sigmoid(x), torch.log(torch.tensor([1e-5 / 1.0, 0.25 / 0.75, 1.0 / 1e-5])))


def test_seeded_array_is_a_pure_function_of_name_shape_seed():
    a = synthetic.seeded_array("decoder.layers.0.linear1.weight", (4, 3), 7)
    assert a.dtype == np.float32 and np.array_equal(a, synthetic.seeded_array("decoder.layers.0.linear1.weight", (4, 3), 7))
    assert not np.array_equal(a, synthetic.seeded_array("decoder.layers.1.linear1.weight", (4, 3), 7))
    assert not np.array_equal(a, synthetic.seeded_array("decoder.layers.0.linear1.weight", (4, 3), 8))
    big = synthetic.seeded_array("x", (1000, 100), 0, -2.0, 2.0)
    assert -2.0 <= big.min() < -1.99 and 1.99 < big.max() < 2.0 and abs(float(big.mean())) < 0.02
    # a fixed known answer so a change of the recipe cannot go unnoticed
    assert synthetic.seeded_array("x.weight", (3, 4), 1)[0, 0] == np.float32(0.62825286)


def test_linear_mode_switch_and_cpu_fallthrough():
    """gemm.linear is nn.Linear unless the tensor-core mode is on AND the operands are CUDA fp32 with supported shapes."""
    from cape_b200 import gemm
    assert cape_b200.linear_mode() == "fp32"
    with pytest.raises(ValueError):
        cape_b200.set_linear_mode("bf16")
    lin = torch.nn.Linear(256, 256)
    x = torch.randn(130, 256)
    want = lin(x)
    old = cape_b200.set_linear_mode("tf32x3")
    try:
        assert old == "fp32" and cape_b200.linear_mode() == "tf32x3"
        assert not gemm.supported(x, lin.weight)                          # CPU tensors never reach the kernel
        assert torch.equal(gemm.linear(lin, x), want)
        assert torch.equal(gemm.linear(lin, x, relu=True), want.relu())
        with pytest.raises(ValueError):
            cape_b200.linear_tf32x3(x, lin.weight, lin.bias)              # the explicit entry point does not fall back
    finally:
        cape_b200.set_linear_mode(old)
    assert not gemm.wgrad_supported(torch.zeros(1000, 256, device="meta"), torch.zeros(1000, 256, device="meta"))


def test_reference_checkpoint_keys_load_into_the_mirror():
    """A CAPEModel-style state dict (prefix base_model.transformer., heads also on the owner, leaked cache buffers) loads
    into the mirror with nothing missing."""
    spec = cape_b200.TokenizerSpec(6, 20)
    src = _small_transformer(spec).attach_heads(*cape_b200.build_prediction_heads(256, 3, 2, True))
    synthetic.fill_parameters_(src, seed=3)
    ckpt = {"base_model.transformer." + k: v.clone() for k, v in src.state_dict().items()}
    for k, v in src.decoder.class_embed.state_dict().items():                      # the owner's copy (roomformer_v2.py:178)
        ckpt["base_model.class_embed." + k] = v.clone()
    ckpt["base_model.query_embed.weight"] = torch.randn(20, 2)
    ckpt["base_model.transformer.decoder.layers.0.kv_cache.k_cache"] = torch.zeros(2, 20, 256)     # Appendix A.2 leak
    ckpt["base_model.transformer.decoder.layers.0.cross_attn.cache.v_cache"] = torch.zeros(1)
    ckpt["base_model.backbone.0.body.conv1.weight"] = torch.zeros(1)              # not ours: ignored
    dst = _small_transformer(spec).attach_heads(*cape_b200.build_prediction_heads(256, 3, 2, True))
    query, missing, unexpected = cape_b200.load_reference_checkpoint(dst, ckpt)
    assert missing == [] and unexpected == [] and torch.equal(query, ckpt["base_model.query_embed.weight"])
    assert all(torch.equal(a, b) for a, b in zip(src.state_dict().values(), dst.state_dict().values()))
    # heads present only on the owner (decoder.* copies absent): still filled
    ckpt2 = {k: v for k, v in ckpt.items() if ".decoder.class_embed." not in k}
    dst2 = _small_transformer(spec).attach_heads(*cape_b200.build_prediction_heads(256, 3, 2, True))
    _, missing, _ = cape_b200.load_reference_checkpoint(dst2, ckpt2)
    assert missing == []
    assert torch.equal(dst2.decoder.class_embed[1].weight, src.decoder.class_embed[1].weight)
    with pytest.raises(KeyError):
        cape_b200.load_reference_checkpoint(dst, {"foo.weight": torch.zeros(1)})
    out = cape_b200.to_cape_predictions({"pred_logits": torch.tensor([[[0.1, 2.0, 0.3]]]), "pred_coords": torch.zeros(1, 1, 2)})
    assert out["sequences"].tolist() == [[1]] and set(out) == {"sequences", "coordinates", "logits"}


@pytest.mark.skipif(not os.path.isdir("/root/reference/models"), reason="needs the reference sources (build container only)")
def test_real_reference_state_dict_loads_into_the_mirror():
    """The reference's own RoomFormerV2 (stub backbone, oracle/make_golden.py) after `_setup_caches` — i.e. with the leaked
    KV / V cache buffers of Appendix A.2 in its state dict — loads into the mirror with nothing missing or unexpected."""
    import sys
    from oracle import make_golden
    syn = make_golden.load_synthetic()
    shapes = ((8, 12), (4, 6), (2, 3), (1, 2))
    feats = [torch.zeros(2, 256, h, w) for h, w in shapes]
    saved_path = list(sys.path)
    try:
        model, tokenizer, _ = make_golden.build_reference_model(syn, 20, 6, feats, None, 31)
    finally:
        sys.path[:] = saved_path          # make_golden puts the reference checkout (with its own `tests/`) in front
    model._setup_caches(2, sum(h * w for h, w in shapes))
    state = model.state_dict()
    assert any(".kv_cache." in k for k in state)                                  # the leak is real
    spec = cape_b200.TokenizerSpec.from_tokenizer(tokenizer)
    tr = _small_transformer(spec).attach_heads(*cape_b200.build_prediction_heads(256, 3, 2, True))
    query, missing, unexpected = cape_b200.load_reference_checkpoint(tr, state)
    assert missing == [] and unexpected == []
    assert torch.equal(query, model.query_embed.weight)
    ref_tr = model.transformer.state_dict()
    for k, v in tr.state_dict().items():
        assert torch.equal(v, ref_tr[k]), k
