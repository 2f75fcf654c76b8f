"""Model-level parity on the GPU (§8c iv): the DeformableTransformer / TransformerDecoder
mirrors teacher-forced (outputs of every decoder layer + gradients) and the device-resident autoregressive generation,
against outputs of the reference's own RoomFormerV2 + DeformableTransformer classes (tests/golden/transformer_model.npz,
made by oracle/make_golden.py).  Weights and feature maps are regenerated from synthetic.seeded_array; the fixture's
checksum guards that both sides saw the same numbers."""
import os

import numpy as np
import pytest
import torch

import cape_b200
from cape_b200 import synthetic
from tests.conftest import GOLDEN, rel_err

pytestmark = pytest.mark.gpu

FWD_TOL_F32 = 1e-5
GRAD_TOL_F32 = 1e-4
SEQ_KEYS = ("seq11", "seq12", "seq21", "seq22", "delta_x1", "delta_x2", "delta_y1", "delta_y2")


@pytest.fixture(scope="module")
def fx():
    g = np.load(os.path.join(GOLDEN, "transformer_model.npz"))
    seed = int(g["seed"])
    spec = cape_b200.TokenizerSpec(int(g["num_bins"]), int(g["seq_len"]))
    tr = cape_b200.DeformableTransformer(
        d_model=256, nhead=8, num_encoder_layers=1, num_decoder_layers=2, dim_feedforward=64, dropout=0.0,
        poly_refine=True, return_intermediate_dec=True, aux_loss=True, num_feature_levels=4, dec_n_points=4,
        enc_n_points=4, query_pos_type="sine", vocab_size=spec.vocab_size, seq_len=spec.seq_len, pad_idx=spec.pad)
    tr.attach_heads(*cape_b200.build_prediction_heads(256, 3, 2, with_poly_refine=True))
    assert sorted(tr.state_dict().keys()) == list(g["state_dict_keys"])           # checkpoint layout of the reference
    assert synthetic.fill_parameters_(tr, seed) == pytest.approx(float(g["weight_checksum"]), rel=1e-12)
    with torch.no_grad():                                                          # the generator script's head recipe
        for head in tr.decoder.class_embed:
            head.weight.mul_(6.0)
            head.bias.copy_(torch.tensor([0.8, 0.0, -0.3]))
    tr = tr.cuda().eval()
    shapes = [tuple(int(v) for v in hw) for hw in g["level_shapes"]]
    n = g["mask0"].shape[0]
    feats = [torch.from_numpy(synthetic.seeded_array(f"feat{i}", (n, 256, h, w), seed)).cuda()
             for i, (h, w) in enumerate(shapes)]
    masks = [torch.from_numpy(g[f"mask{i}"]).cuda() for i in range(4)]
    pos = [torch.from_numpy(g[f"pos{i}"]).cuda() for i in range(4)]
    query_embed = torch.from_numpy(synthetic.seeded_array("query_embed.weight", (spec.seq_len, 2), seed, -2, 2)).cuda()
    sup = torch.from_numpy(synthetic.seeded_array("support_features", (n, int(g["n_sup"]), 256), seed)).cuda()
    sup_mask = torch.from_numpy(g["support_mask"]).cuda()
    return dict(g=g, seed=seed, spec=spec, tr=tr, feats=feats, masks=masks, pos=pos, query_embed=query_embed, sup=sup,
                sup_mask=sup_mask)


def test_teacher_forced_forward_and_gradients_match_reference_model(fx):
    g, tr = fx["g"], fx["tr"]
    seq_kwargs = {k: torch.from_numpy(g[k]).cuda() for k in SEQ_KEYS}
    feats = [f.clone().requires_grad_(True) for f in fx["feats"]]
    qe = fx["query_embed"].clone().requires_grad_(True)
    causal = tr._create_causal_attention_mask(fx["spec"].seq_len).cuda()             # RoomFormerV2.attention_mask (:266)
    hs, init_ref, refs, classes = tr(feats, fx["masks"], fx["pos"], qe, None, causal, seq_kwargs,
                                     support_features=fx["sup"], support_mask=fx["sup_mask"])
    assert hs.shape == (2, 2, fx["spec"].seq_len, 256)
    assert rel_err(init_ref.detach().cpu().numpy()[0], torch.from_numpy(
        synthetic.seeded_array("query_embed.weight", (fx["spec"].seq_len, 2), fx["seed"], -2, 2)).sigmoid().numpy()) < 1e-6
    assert rel_err(classes.detach().cpu().numpy(), g["pred_logits"]) < FWD_TOL_F32
    assert rel_err(refs.detach().cpu().numpy(), g["pred_coords"]) < FWD_TOL_F32
    seed = fx["seed"]
    loss = (classes * torch.from_numpy(synthetic.seeded_array("grad_logits", classes.shape, seed)).cuda()).sum() \
        + (refs * torch.from_numpy(synthetic.seeded_array("grad_coords", refs.shape, seed)).cuda()).sum()
    names = [k[len("grad_param."):] for k in g.files if k.startswith("grad_param.")]
    params = dict(tr.named_parameters())
    grads = torch.autograd.grad(loss, [params[k] for k in names] + [qe] + feats)
    errs = {k: rel_err(gr.cpu().numpy()[:8], g["grad_param." + k]) for k, gr in zip(names, grads)}
    errs["query_embed"] = rel_err(grads[len(names)].cpu().numpy(), g["grad_query_embed"])
    for i in range(4):
        errs[f"feat{i}"] = rel_err(grads[len(names) + 1 + i].cpu().numpy(), g[f"grad_feat{i}"])
    assert max(errs.values()) < GRAD_TOL_F32, errs


def _check_generation(out, g, prefix):
    want_logits, want_coords = g[prefix + "_logits"], g[prefix + "_coords"]
    assert out["pred_logits"].shape == want_logits.shape, (out["pred_logits"].shape, want_logits.shape)
    assert np.array_equal(out["sequences"].cpu().numpy(), want_logits.argmax(-1))    # same generated token types
    assert rel_err(out["pred_logits"].cpu().numpy(), want_logits) < 1e-4             # 20 chained steps of fp32 GEMMs
    assert rel_err(out["pred_coords"].cpu().numpy(), want_coords) < 1e-4


def test_device_resident_generation_matches_reference_forward_inference(fx):
    g, tr, spec = fx["g"], fx["tr"], fx["spec"]
    gen = cape_b200.AutoregressiveGenerator(tr, spec, max_batch_size=2, device="cuda")
    before = cape_b200.launch_count()
    out = gen.generate(fx["feats"], fx["masks"], fx["pos"], fx["query_embed"], fx["sup"], fx["sup_mask"], poll_every=4)
    assert cape_b200.launch_count() > before
    _check_generation(out, g, "gen")
    kind = np.array([[0 if isinstance(e, list) else e for e in row] for row in out["gen_out"]])
    assert np.array_equal(kind, g["gen_kind"])
    xy = np.array([[e if isinstance(e, list) else [0.0, 0.0] for e in row] for row in out["gen_out"]], dtype=np.float32)
    assert rel_err(xy, g["gen_xy"]) < 1e-4
    # same object, next batch (graph reused), eager stepping: identical tokens
    out2 = gen.generate(fx["feats"], fx["masks"], fx["pos"], fx["query_embed"], fx["sup"], fx["sup_mask"], use_graph=False)
    assert torch.equal(out2["sequences"], out["sequences"])
    assert rel_err(out2["pred_coords"].cpu().numpy(), out["pred_coords"].cpu().numpy()) < 1e-5


def test_generation_stops_when_every_sequence_emitted_eos(fx):
    g, tr, spec = fx["g"], fx["tr"], fx["spec"]
    saved = [h.bias.detach().clone() for h in tr.decoder.class_embed]
    try:
        with torch.no_grad():
            for h in tr.decoder.class_embed:
                h.bias.copy_(torch.from_numpy(g["gen2_bias"]).cuda())
        gen = cape_b200.AutoregressiveGenerator(tr, spec, max_batch_size=2, device="cuda")
        for poll in (1, 3, 16):                      # polling granularity must not change the result
            out = gen.generate(fx["feats"], fx["masks"], fx["pos"], fx["query_embed"], fx["sup"], fx["sup_mask"],
                               poll_every=poll)
            assert out["steps"] == g["gen2_logits"].shape[1] < spec.seq_len
            _check_generation(out, g, "gen2")
        eager = cape_b200.generate_eager(tr, spec, fx["feats"], fx["masks"], fx["pos"], fx["query_embed"], fx["sup"],
                                         fx["sup_mask"])
        _check_generation(eager, g, "gen2")
    finally:
        with torch.no_grad():
            for h, b in zip(tr.decoder.class_embed, saved):
                h.bias.copy_(b)


def test_seq_embed_op_matches_the_eager_expression_bitwise_and_in_gradient():
    gen = torch.Generator().manual_seed(2)
    v, c, b, t, pad = 40, 256, 3, 17, 39
    table = torch.randn(v, c, generator=gen)
    seqs = [torch.randint(0, v, (b, t), generator=gen) for _ in range(4)]
    seqs[0][0, 3] = pad
    dx, dy = torch.rand(b, t, generator=gen), torch.rand(b, t, generator=gen)
    deltas = [dx, 1 - dx, dy, 1 - dy]                                    # x1, x2, y1, y2
    emb = torch.nn.Embedding(v, c, padding_idx=pad)
    with torch.no_grad():
        emb.weight.copy_(table)
    e11, e12, e21, e22 = (emb(s) for s in seqs)
    x1, x2, y1, y2 = (d[..., None] for d in deltas)
    want = e11 * x2 * y2 + e21 * x1 * y2 + e12 * x2 * y1 + e22 * x1 * y1   # deformable_transformer_v2.py:991-995
    gout = torch.randn(b, t, c, generator=gen)
    (want_grad,) = torch.autograd.grad(want, emb.weight, gout)
    tab = table.cuda().requires_grad_(True)
    got = cape_b200.seq_embed(tab, *[s.cuda() for s in seqs], *[d.cuda() for d in deltas], pad)
    assert torch.equal(got.detach().cpu(), want.detach())
    (grad,) = torch.autograd.grad(got, tab, gout.cuda())
    assert rel_err(grad.cpu().numpy(), want_grad.numpy()) < 1e-6
    assert float(grad[pad].abs().max()) == 0.0
    torch.library.opcheck(torch.ops.cape.seq_embed.default,
                          (tab, *[s.cuda() for s in seqs], *[d.cuda() for d in deltas], pad),
                          test_utils=("test_schema", "test_faketensor", "test_autograd_registration"))


def test_token_step_follows_the_reference_bookkeeping_rules():
    """Scripted head outputs through cape_token_step vs a literal Python transcription of roomformer_v2.py:548-597."""
    import math
    spec = cape_b200.TokenizerSpec(num_bins=44, seq_len=12)
    b = 5
    state = cape_b200.TokenState(b, spec, 3, "cuda")
    rng = np.random.default_rng(0)
    unfinished = np.ones(b)
    for i in range(spec.seq_len):
        logits = rng.standard_normal((b, 1, 3)).astype(np.float32)
        reg = rng.random((b, 1, 2)).astype(np.float32)
        reg[0, 0] = (1.0, 0.0)
        state.advance(torch.from_numpy(logits).cuda(), torch.from_numpy(reg).cuda())
        torch.cuda.synchronize()
        seq = [s.cpu().numpy()[:, 0] for s in state.seq]
        delta = [d.cpu().numpy()[:, 0] for d in state.delta]
        for j in range(b):
            cls = int(np.argmax(logits[j, 0]))
            dx = dy = 0
            if unfinished[j] == 1:
                if cls == 0 or (cls == 2 and i < 6):
                    x, y = reg[j, 0]
                    x, y = min(x, 1) * (spec.num_bins - 1), min(y, 1) * (spec.num_bins - 1)
                    want = [math.floor(x) * 44 + math.floor(y), math.floor(x) * 44 + math.ceil(y),
                            math.ceil(x) * 44 + math.floor(y), math.ceil(x) * 44 + math.ceil(y)]
                    dx, dy = x - math.floor(x), y - math.floor(y)
                elif cls == 1:
                    want = [spec.sep] * 4
                else:
                    unfinished[j] = 0
                    want = [spec.eos] * 4
            else:
                want = [spec.pad] * 4
            assert [int(s[j]) for s in seq] == want, (i, j)
            assert [float(d[j]) for d in delta] == [np.float32(dx), np.float32(1 - np.float32(dx)), np.float32(dy),
                                                    np.float32(1 - np.float32(dy))], (i, j)
        assert np.array_equal(state.unfinished.cpu().numpy(), unfinished.astype(np.int32))
    assert int(state.step.item()) == spec.seq_len
    state.advance(torch.zeros(b, 1, 3).cuda(), torch.zeros(b, 1, 2).cuda())        # past max_len: a no-op
    torch.cuda.synchronize()
    assert int(state.step.item()) == spec.seq_len + 1


# ---- hand-written decode-step kernels (csrc/decode_step.cu) vs plain PyTorch fp32 ----------------------------------
def test_generation_with_library_kernels_and_with_step_kernels_agree_with_the_reference(fx):
    g, tr, spec = fx["g"], fx["tr"], fx["spec"]
    for fused in (False, True):
        gen = cape_b200.AutoregressiveGenerator(tr, spec, max_batch_size=2, device="cuda", fused=fused)
        assert gen.fused is fused
        out = gen.generate(fx["feats"], fx["masks"], fx["pos"], fx["query_embed"], fx["sup"], fx["sup_mask"])
        _check_generation(out, g, "gen")


@pytest.mark.parametrize("rows,k,n", [(128, 256, 256), (5, 256, 768), (2, 1024, 256), (7, 256, 1024), (3, 256, 128), (1, 64, 36)])
def test_skinny_linear_epilogues_vs_torch(rows, k, n):
    from cape_b200 import decode_ops as K
    gen = torch.Generator().manual_seed(rows * 1000 + n)
    lin = torch.nn.Linear(k, n)
    x, x2, res = torch.randn(rows, k, generator=gen), torch.randn(rows, k, generator=gen), torch.randn(rows, n, generator=gen)
    ln = torch.nn.LayerNorm(n)
    with torch.no_grad():
        ln.weight.add_(torch.randn(n, generator=gen) * 0.1)
        ln.bias.add_(torch.randn(n, generator=gen) * 0.1)
    lin, ln = lin.cuda(), ln.cuda()
    wt = lin.weight.detach().t().contiguous()
    xc, x2c, resc = x.cuda(), x2.cuda(), res.cuda()
    with torch.no_grad():
        assert rel_err(K.skinny_linear(xc, wt, lin.bias).cpu(), lin(xc).cpu()) < FWD_TOL_F32
        assert rel_err(K.skinny_linear(xc, wt, None, x2=x2c).cpu(), torch.nn.functional.linear(xc + x2c, lin.weight).cpu()) < FWD_TOL_F32
        assert rel_err(K.skinny_linear(xc, wt, lin.bias, relu=True).cpu(), lin(xc).relu().cpu()) < FWD_TOL_F32
        # strided input rows (a slice of a wider projection output)
        wide = torch.randn(rows, k + 64, generator=gen).cuda()
        assert rel_err(K.skinny_linear(wide[:, 32:32 + k], wt, lin.bias).cpu(), lin(wide[:, 32:32 + k]).cpu()) < FWD_TOL_F32
        if n <= 256:
            got = K.skinny_linear(xc, wt, lin.bias, residual=resc, gamma=ln.weight, beta=ln.bias, eps=ln.eps)
            assert rel_err(got.cpu(), ln(resc + lin(xc)).cpu()) < FWD_TOL_F32
            got = K.skinny_linear(xc, wt, lin.bias, gamma=ln.weight, beta=ln.bias, eps=ln.eps)
            assert rel_err(got.cpu(), ln(lin(xc)).cpu()) < FWD_TOL_F32


def test_skinny_linear_sine_input_matches_query_pos_embedding():
    from cape_b200 import decode_ops as K
    torch.manual_seed(0)
    lin, ln = torch.nn.Linear(256, 256).cuda(), torch.nn.LayerNorm(256).cuda()
    ref = torch.rand(9, 2, device="cuda")
    dim_t = torch.arange(128, dtype=torch.float32, device="cuda")
    dim_t = 10000 ** (2 * (dim_t // 2) / 128)
    with torch.no_grad():
        want = ln(lin(cape_b200.TransformerDecoder.get_query_pos_embed(ref[:, None])[:, 0]))
        got = K.skinny_linear(ref, lin.weight.t().contiguous(), lin.bias, gamma=ln.weight, beta=ln.bias, eps=ln.eps,
                              sine_dim_t=dim_t)
    assert rel_err(got.cpu(), want.cpu()) < FWD_TOL_F32


def test_tiny_linear_and_refinement_vs_torch():
    from cape_b200 import decode_ops as K
    from cape_b200.transformer import inverse_sigmoid
    torch.manual_seed(1)
    for n in (2, 3, 8):
        lin = torch.nn.Linear(256, n).cuda()
        x = torch.randn(11, 256, device="cuda")
        ref = torch.rand(11, n, device="cuda")
        ref[0, 0], ref[1, 0] = 0.0, 1.0                                   # the clamps of inverse_sigmoid
        with torch.no_grad():
            assert rel_err(K.tiny_linear(x, lin.weight, lin.bias).cpu(), lin(x).cpu()) < FWD_TOL_F32
            got = K.tiny_linear(x, lin.weight, lin.bias, refine_ref=ref)
            assert rel_err(got.cpu(), (lin(x) + inverse_sigmoid(ref)).sigmoid().cpu()) < FWD_TOL_F32


def test_decode_attention_vs_multihead_attention_with_kv_cache():
    """Self-attention form against nn.MultiheadAttention over the growing prefix (what the reference layer does with its
    KVCache, deformable_transformer_v2.py:322-341), and the cross-attention form with a key-padding mask (:350-357)."""
    from cape_b200 import decode_ops as K
    torch.manual_seed(5)
    b, t_max, c, heads = 3, 40, 256, 8
    mha = torch.nn.MultiheadAttention(c, heads, batch_first=True).cuda().eval()
    wq, wk, wv = mha.in_proj_weight.chunk(3)
    bq, bk, bv = mha.in_proj_bias.chunk(3)
    xs = torch.randn(b, t_max, c, device="cuda")
    k_cache = torch.zeros(b, t_max, c, device="cuda")
    v_cache = torch.zeros(b, t_max, c, device="cuda")
    pos = torch.zeros(1, dtype=torch.int64, device="cuda")
    lin = torch.nn.functional.linear
    with torch.no_grad():
        for i in (0, 1, 2, 31, 32, 33, 39):                               # crosses the 32-key lane blocks
            # fill the cache up to i-1 exactly, then decode token i
            k_cache[:, :i] = lin(xs[:, :i], wk, bk)
            v_cache[:, :i] = lin(xs[:, :i], wv, bv)
            k_cache[:, i:], v_cache[:, i:] = 7.0, 7.0                      # stale rows beyond the position must not matter
            pos.fill_(i)
            x = xs[:, i]
            proj = torch.cat([lin(x, wq, bq), lin(x, wk, bk), lin(x, wv, bv)], 1)          # strided slices, as the generator passes
            got = K.decode_attention(proj[:, :c], k_cache, v_cache, proj[:, c:2 * c], proj[:, 2 * c:], pos, n_heads=heads)
            want = mha(xs[:, i:i + 1], xs[:, :i + 1], xs[:, :i + 1], need_weights=False)[0][:, 0]
            assert rel_err(mha.out_proj(got).cpu(), want.cpu()) < FWD_TOL_F32, i
            assert torch.equal(k_cache[:, i], proj[:, c:2 * c]) and torch.equal(v_cache[:, i], proj[:, 2 * c:])
        # cross-attention over 17 support keys, two of them padded for one sample
        sup = torch.randn(b, 17, c, device="cuda")
        pad = torch.zeros(b, 17, dtype=torch.bool, device="cuda")
        pad[1, -2:] = True
        bias = torch.zeros(b, 17, device="cuda").masked_fill_(pad, float("-inf"))
        q = torch.randn(b, c, device="cuda")
        got = K.decode_attention(lin(q, wq, bq), lin(sup, wk, bk).contiguous(), lin(sup, wv, bv).contiguous(), key_bias=bias,
                                 n_heads=heads)
        want = mha(q[:, None], sup, sup, key_padding_mask=pad, need_weights=False)[0][:, 0]
        assert rel_err(mha.out_proj(got).cpu(), want.cpu()) < FWD_TOL_F32
        pos.fill_(t_max)                                                   # out-of-range position: a no-op, not a fault
        K.decode_attention(proj[:, :c], k_cache, v_cache, proj[:, c:2 * c], proj[:, 2 * c:], pos, n_heads=heads)
        torch.cuda.synchronize()


def test_split_linear_and_coordinate_head_fusions_vs_torch():
    from cape_b200 import decode_ops as K
    from cape_b200.transformer import MLP, inverse_sigmoid
    torch.manual_seed(3)
    for rows in (1, 6, 128):
        lin = torch.nn.Linear(256, 384).cuda()
        x, x2 = torch.randn(rows, 256, device="cuda"), torch.randn(rows, 256, device="cuda")
        with torch.no_grad():
            want = lin(x + x2)
            y, y2 = K.skinny_linear_split(x, lin.weight.t().contiguous(), lin.bias, 256, x2=x2)
        assert y.shape == (rows, 256) and y2.shape == (rows, 128) and y.is_contiguous() and y2.is_contiguous()
        assert rel_err(torch.cat([y, y2], 1).cpu(), want.cpu()) < FWD_TOL_F32
        # coordinate MLP (roomformer_v2.py:956-968) layers 2 + 3, refinement (:1096-1102), per-level centres (:1072)
        mlp = MLP(256, 256, 2, 3).cuda()
        ref = torch.rand(rows, 2, device="cuda")
        ref[0, 0] = 1.0
        valid = torch.rand(rows, 4, 2, device="cuda") * 0.5 + 0.5
        with torch.no_grad():
            h1 = mlp.layers[0](x).relu()
            want_ref = (mlp.layers[2](mlp.layers[1](h1).relu()) + inverse_sigmoid(ref)).sigmoid()
            got_ref, got_levels = K.coord_head_refine(h1, mlp.layers[1].weight.t().contiguous(), mlp.layers[1].bias,
                                                      mlp.layers[2].weight.contiguous(), mlp.layers[2].bias, ref, valid)
        assert rel_err(got_ref.cpu(), want_ref.cpu()) < FWD_TOL_F32
        assert rel_err(got_levels.cpu(), (want_ref[:, None, :] * valid).cpu()) < FWD_TOL_F32


@pytest.mark.parametrize("rows,lq,m", [(1, 1, 8), (6, 1, 8), (128, 1, 8), (3, 5, 8), (7, 1, 4)])
def test_sampling_fused_into_the_output_projection_equals_the_two_launch_path(rows, lq, m):
    """cape_msda_output_proj (MSDeformAttn sampling as the input stage of output_proj + residual + norm1, one launch,
    deformable_transformer.py:99-113 + deformable_transformer_v2.py:360-364) against cape::ms_deform_attn_decode followed by
    the skinny linear — the same device code in the same order, so bit-identical — and against plain torch."""
    from cape_b200 import decode_ops as K
    g = torch.Generator().manual_seed(rows * 10 + lq)
    shapes = torch.tensor(synthetic.CAPE_PYRAMID)
    starts = cape_b200.level_start_index_from_shapes(shapes)
    s_len = int((shapes[:, 0] * shapes[:, 1]).sum())
    n_out = 32 * m
    value = torch.randn(rows, s_len, m, 32, generator=g).cuda()
    ref = torch.rand(rows, lq, 4, 2, generator=g).cuda()
    off = (torch.randn(rows, lq, m, 4, 4, 2, generator=g) * 3).cuda()
    logits = torch.randn(rows, lq, m, 16, generator=g).cuda()
    lin = torch.nn.Linear(32 * m, n_out).cuda()
    ln = torch.nn.LayerNorm(n_out).cuda()
    with torch.no_grad():
        ln.weight.uniform_(0.5, 1.5)
        ln.bias.uniform_(-0.5, 0.5)
    res = torch.randn(rows * lq, n_out, generator=g).cuda()
    wt = lin.weight.detach().t().contiguous()
    shapes_d, starts_d = shapes.cuda(), starts.cuda()
    with torch.no_grad():
        sampled = torch.ops.cape.ms_deform_attn_decode(value, shapes_d, starts_d, ref, off, logits).view(rows * lq, -1)
        two = K.skinny_linear(sampled, wt, lin.bias, residual=res, gamma=ln.weight, beta=ln.bias, eps=ln.eps)
        before = cape_b200.launch_count()
        one = K.msda_output_proj(value, shapes_d, starts_d, ref, off, logits, wt, lin.bias, residual=res, gamma=ln.weight,
                                 beta=ln.bias, eps=ln.eps)
        assert cape_b200.launch_count() == before + 1
        want = ln(res + lin(sampled))
    torch.cuda.synchronize()
    assert torch.equal(one, two)
    assert rel_err(one.cpu(), want.cpu()) < FWD_TOL_F32


def test_fused_output_projection_rejects_what_it_does_not_cover():
    """cape_msda_output_proj: error codes instead of launches for dimensions outside D = 32 / P = 4 / L = 4, N > 256, missing
    LayerNorm parameters; zero rows is a no-op."""
    import ctypes
    from cape_b200 import _lib
    lib = _lib.load()
    dev = "cuda"
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    value = torch.zeros(2, 5440, 8, 32, device=dev)
    shapes = torch.tensor(synthetic.CAPE_PYRAMID, device=dev)
    starts = cape_b200.level_start_index_from_shapes(shapes.cpu()).to(dev)
    ref, off, lg = torch.zeros(2, 1, 4, 2, device=dev), torch.zeros(2, 1, 8, 4, 4, 2, device=dev), torch.zeros(2, 1, 8, 16, device=dev)
    wt, vec, y = torch.zeros(256, 256, device=dev), torch.zeros(512, device=dev), torch.zeros(2, 512, device=dev)

    def call(dims, n_out=256, gamma=vec):
        return lib.cape_msda_output_proj(p(value), p(shapes), p(starts), p(ref), p(off), p(lg), ctypes.byref(dims), p(wt), p(vec),
                                         p(y), 512, None if gamma is None else p(gamma), p(vec), 1e-5, p(y), 512, n_out, None)
    before = cape_b200.launch_count()
    assert call(_lib.Dims(2, 5440, 8, 32, 1, 4, 4)) == 0
    assert cape_b200.launch_count() == before + 1
    assert call(_lib.Dims(2, 5440, 8, 32, 1, 4, 2)) == -1            # P != 4
    assert call(_lib.Dims(2, 5440, 8, 16, 1, 4, 4)) == -1            # D != 32
    assert call(_lib.Dims(2, 5440, 8, 32, 1, 3, 4)) == -1            # L != 4
    assert call(_lib.Dims(2, 5440, 8, 32, 1, 4, 4), n_out=512) == -1  # LayerNorm epilogue: N <= 256
    assert call(_lib.Dims(2, 5440, 8, 32, 1, 4, 4), gamma=None) == -1
    assert call(_lib.Dims(0, 5440, 8, 32, 1, 4, 4)) == 0             # no rows: nothing launched
    assert cape_b200.launch_count() == before + 1
    torch.cuda.synchronize()


# ---- 3xTF32 tensor-core linear (csrc/linear_tf32x3.cu) ---------------------------------------------------------------
@pytest.mark.parametrize("m,n,k", [(128, 128, 32), (1, 128, 64), (300, 256, 256), (4099, 1024, 256), (1000, 256, 1024),
                                    (129, 384, 96)])
def test_linear_tf32x3_matches_fp64_at_fp32_level(m, n, k):
    """The hi/lo split keeps the tensor-core GEMM at fp32-level accuracy: both it and cuBLAS' fp32 kernel are compared
    with an fp64 product; plain TF32 (one pass) would be off by ~1e-3."""
    gen = torch.Generator().manual_seed(m + n + k)
    x = torch.randn(m, k, generator=gen).cuda()
    w = (torch.randn(n, k, generator=gen) / k ** 0.5).cuda()
    b = torch.randn(n, generator=gen).cuda()
    ref = x.double() @ w.double().t() + b.double()
    got = cape_b200.linear_tf32x3(x, w, b)
    assert got.shape == (m, n) and not torch.isnan(got).any()
    err = rel_err(got.cpu(), ref.cpu())
    err_fp32 = rel_err(torch.nn.functional.linear(x, w, b).cpu(), ref.cpu())
    assert err < 1e-5 and err < 12 * max(err_fp32, 1e-7), (err, err_fp32)
    relu = cape_b200.linear_tf32x3(x.view(1, m, k), w, None, relu=True)            # leading dims, no bias, fused ReLU
    assert relu.shape == (1, m, n)
    assert rel_err(relu.cpu(), (x.double() @ w.double().t()).clamp_min(0).cpu()[None]) < 1e-5


def test_linear_mode_routes_inference_only_and_tracks_weight_updates():
    lin = torch.nn.Linear(256, 256).cuda()
    x = torch.randn(4, 200, 256, device="cuda")
    old = cape_b200.set_linear_mode("tf32x3")
    try:
        from cape_b200 import gemm
        before = cape_b200.launch_count()
        xg = x.clone().requires_grad_(True)
        y_train = gemm.linear(lin, xg)                                   # autograd on: kernel forward + input gradient
        assert y_train.requires_grad and cape_b200.launch_count() > before
        gout = torch.randn_like(y_train)
        gx, gw, gb = torch.autograd.grad(y_train, (xg, lin.weight, lin.bias), gout)
        xr = x.clone().requires_grad_(True)
        rx, rw, rb = torch.autograd.grad(lin(xr), (xr, lin.weight, lin.bias), gout)
        assert rel_err(gx.cpu(), rx.cpu()) < 1e-5 and rel_err(gw.cpu(), rw.cpu()) < 1e-5 and rel_err(gb.cpu(), rb.cpu()) < 1e-5
        before = cape_b200.launch_count()
        with torch.no_grad():
            y = gemm.linear(lin, x)
            assert cape_b200.launch_count() > before
            assert rel_err(y.cpu(), y_train.detach().cpu()) < 1e-5
            lin.weight.mul_(2.0)                                         # in-place update: the cached lo part must follow
            y2 = gemm.linear(lin, x)
            assert rel_err(y2.cpu(), lin(x).cpu()) < 1e-5
            small = gemm.linear(lin, x[:1, :3])                          # tiny inputs stay on nn.Linear
            assert torch.equal(small, lin(x[:1, :3]))
            with torch.autocast("cuda", dtype=torch.float16):            # AMP keeps its own (half-precision) Linear
                assert gemm.linear(lin, x).dtype == torch.float16
    finally:
        cape_b200.set_linear_mode(old)
    assert cape_b200.linear_mode() == old


def test_generation_with_tensor_core_linears_produces_the_reference_tokens(fx):
    g, tr, spec = fx["g"], fx["tr"], fx["spec"]
    old = cape_b200.set_linear_mode("tf32x3")
    try:
        gen = cape_b200.AutoregressiveGenerator(tr, spec, max_batch_size=2, device="cuda")
        out = gen.generate(fx["feats"], fx["masks"], fx["pos"], fx["query_embed"], fx["sup"], fx["sup_mask"])
    finally:
        cape_b200.set_linear_mode(old)
    _check_generation(out, g, "gen")


@pytest.mark.parametrize("rows,n,k", [(4096, 256, 256), (2048, 1024, 256), (20 * 5440, 256, 1024), (1024, 128, 256)])
def test_linear_tf32x3_weight_gradient_split_over_the_rows(rows, n, k):
    """grad_w = g^T x on the 3xTF32 kernel (transposed operands, reduction split over the SMs, partial tiles added by the
    TMA) against an fp64 product; cuBLAS fp32 is measured on the same inputs for scale."""
    from cape_b200 import gemm
    gen = torch.Generator().manual_seed(rows + n + k)
    g = torch.randn(rows, n, generator=gen).cuda()
    x = torch.randn(rows, k, generator=gen).cuda()
    assert gemm.wgrad_supported(g, x)
    ref = g.double().t() @ x.double()
    got = gemm.linear_tf32x3_wgrad(g, x)
    again = gemm.linear_tf32x3_wgrad(g, x)
    assert got.shape == (n, k)
    err, err_fp32 = rel_err(got.cpu(), ref.cpu()), rel_err((g.t() @ x).cpu(), ref.cpu())
    assert err < 1e-5 and err < 12 * max(err_fp32, 1e-7), (err, err_fp32)
    rerun = rel_err(again.cpu(), got.cpu())
    assert rerun < 5e-6, rerun                                         # order of the TMA adds may differ run to run
    # end to end through autograd
    lin = torch.nn.Linear(k, n).cuda()
    old = cape_b200.set_linear_mode("tf32x3")
    try:
        xg = x.clone().requires_grad_(True)
        gx, gw, gb = torch.autograd.grad(gemm.linear(lin, xg), (xg, lin.weight, lin.bias), g)
    finally:
        cape_b200.set_linear_mode(old)
    xr = x.clone().requires_grad_(True)
    rx, rw, rb = torch.autograd.grad(lin(xr), (xr, lin.weight, lin.bias), g)
    errs = (rel_err(gx.cpu(), rx.cpu()), rel_err(gw.cpu(), rw.cpu()), rel_err(gb.cpu(), rb.cpu()))
    assert max(errs) < 1e-5, errs


def test_empty_and_degenerate_inputs_of_the_sequence_and_step_kernels():
    """Zero rows / tokens are no-ops (not launches with a zero grid), and the ABI rejects what the kernels cannot take."""
    import ctypes
    from cape_b200 import _lib, decode_ops as K
    lib = _lib.load()
    dev = "cuda"
    table = torch.randn(10, 256, device=dev)
    empty_idx = torch.zeros(2, 0, dtype=torch.int64, device=dev)
    empty_d = torch.zeros(2, 0, device=dev)
    out = cape_b200.seq_embed(table, empty_idx, empty_idx, empty_idx, empty_idx, empty_d, empty_d, empty_d, empty_d, -1)
    assert out.shape == (2, 0, 256)
    bad = torch.full((1, 1), 99, dtype=torch.int64, device=dev)                    # token id outside the table -> NaN row
    one = torch.ones(1, 1, device=dev)
    assert torch.isnan(cape_b200.seq_embed(table, bad, bad, bad, bad, one, one, one, one, -1)).all()
    w = torch.randn(256, 256, device=dev)
    assert K.skinny_linear(torch.zeros(0, 256, device=dev), w).shape == (0, 256)
    assert K.tiny_linear(torch.zeros(0, 256, device=dev), torch.randn(3, 256, device=dev)).shape == (0, 3)
    kc = torch.zeros(0, 8, 256, device=dev)
    assert K.decode_attention(torch.zeros(0, 256, device=dev), kc, kc).shape == (0, 256)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    x = torch.randn(4, 256, device=dev)
    y = torch.empty(4, 256, device=dev)
    # K not a multiple of 16 / unknown epilogue / LayerNorm without gamma / misaligned rows
    assert lib.cape_skinny_linear(p(x), 256, None, 0, p(w), None, None, 0, None, None, 1e-5, None, p(y), 256, 4, 250, 256, 0, None) == -1
    assert lib.cape_skinny_linear(p(x), 256, None, 0, p(w), None, None, 0, None, None, 1e-5, None, p(y), 256, 4, 256, 256, 7, None) == -1
    assert lib.cape_skinny_linear(p(x), 256, None, 0, p(w), None, None, 0, None, None, 1e-5, None, p(y), 256, 4, 256, 256, 2, None) == -1
    assert lib.cape_skinny_linear(p(x), 255, None, 0, p(w), None, None, 0, None, None, 1e-5, None, p(y), 256, 4, 256, 256, 0, None) == -4
    # head dim other than 32 / more than 1024 keys
    assert lib.cape_decode_attention(p(x), 256, None, None, 0, p(x), p(x), None, None, p(y), 1, 4, 4, 64, None) == -1
    assert lib.cape_decode_attention(p(x), 256, None, None, 0, p(x), p(x), None, None, p(y), 1, 2048, 8, 32, None) == -1
    # 3xTF32 linear: N not a multiple of 128, K not a multiple of 32; host pointers are rejected
    assert lib.cape_linear_tf32x3(p(x), p(w), p(w), None, p(y), 4, 200, 256, 0, None) == -1
    assert lib.cape_linear_tf32x3(p(x), p(w), p(w), None, p(y), 4, 256, 250, 0, None) == -1
    host = torch.randn(4, 256)
    assert lib.cape_linear_tf32x3(p(host), p(w), p(w), None, p(y), 4, 256, 256, 0, None) == -5
    assert b"host memory" in lib.cape_last_error()
    with pytest.raises(ValueError):
        cape_b200.linear_tf32x3(torch.zeros(0, 256, device=dev), w)


def test_guard_zones_around_outputs_of_the_step_and_gemm_kernels():
    """compute-sanitizer is closed on this pool: outputs (and inputs) of the new kernels are carved out of one NaN-filled
    arena and called through the C ABI directly; afterwards every guard word must be untouched — ragged row counts
    exercise the TMA store clipping, the split-K TMA reduce and the partial tiles of the skinny kernel."""
    import ctypes
    from cape_b200 import _lib
    lib = _lib.load()
    guard = 4096
    sizes = {"x": 300 * 256, "w": 256 * 256, "w_lo": 256 * 256, "bias": 256, "y": 300 * 256,             # 3xTF32 forward, M = 300
             "g": 1056 * 128, "xg": 1056 * 256, "gw": 128 * 256, "ws": (128 + 2 * 256) * 1056,             # weight gradient, 1056 rows
             "sx": 5 * 256, "swt": 256 * 256, "sy": 5 * 256, "gamma": 256, "beta": 256,                    # skinny linear + LayerNorm, 5 rows
             "q": 3 * 256, "knew": 3 * 256, "vnew": 3 * 256, "kc": 3 * 9 * 256, "vc": 3 * 9 * 256, "ao": 3 * 256}
    total = sum(v + 2 * guard for v in sizes.values())
    arena = torch.full((total,), float("nan"), device="cuda")
    views, off = {}, 0
    for k, numel in sizes.items():
        off += guard
        views[k] = arena[off:off + numel]
        off += numel + guard
    gen = torch.Generator(device="cuda").manual_seed(0)
    for k in ("x", "w", "bias", "g", "xg", "sx", "swt", "gamma", "beta", "q", "knew", "vnew", "kc", "vc"):
        views[k].copy_(torch.randn(views[k].numel(), device="cuda", generator=gen) * 0.1)
    for k in ("w_lo", "y", "gw", "ws", "sy", "ao"):
        views[k].zero_()
    snapshot = arena.clone()
    outputs = ("w_lo", "y", "gw", "ws", "sy", "ao", "kc", "vc")
    p = lambda k: ctypes.c_void_p(views[k].data_ptr())
    sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    pos = torch.full((1,), 4, dtype=torch.int64, device="cuda")
    _lib.check(lib.cape_tf32_split_lo(p("w"), p("w_lo"), 256 * 256, sp), "split")
    _lib.check(lib.cape_linear_tf32x3(p("x"), p("w"), p("w_lo"), p("bias"), p("y"), 300, 256, 256, 1, sp), "gemm")
    _lib.check(lib.cape_linear_tf32x3_wgrad(p("g"), p("xg"), p("gw"), p("ws"), 1056, 128, 256, sp), "wgrad")
    _lib.check(lib.cape_skinny_linear(p("sx"), 256, None, 0, p("swt"), p("bias"), p("sx"), 256, p("gamma"), p("beta"), 1e-5, None,
                                      p("sy"), 256, 5, 256, 256, 2, sp), "skinny")
    _lib.check(lib.cape_decode_attention(p("q"), 256, p("knew"), p("vnew"), 256, p("kc"), p("vc"), ctypes.c_void_p(pos.data_ptr()),
                                         None, p("ao"), 3, 9, 8, 32, sp), "attention")
    torch.cuda.synchronize()
    mask = torch.ones(total, dtype=torch.bool, device="cuda")           # everything except the output tensors must be unchanged
    off = 0
    for k, numel in sizes.items():
        off += guard
        if k in outputs:
            mask[off:off + numel] = False
        off += numel + guard
    same = (arena == snapshot) | (torch.isnan(arena) & torch.isnan(snapshot))
    assert bool(same[mask].all()), "a kernel wrote outside its output tensors"
    for k in ("y", "gw", "sy", "ao"):
        assert bool(torch.isfinite(views[k]).all()), k
    # and the results are right
    x, w, b = views["x"].view(300, 256), views["w"].view(256, 256), views["bias"]
    assert rel_err(views["y"].view(300, 256).cpu(), (x.double() @ w.double().t() + b.double()).clamp_min(0).cpu()) < 1e-5
    g, xg = views["g"].view(1056, 128), views["xg"].view(1056, 256)
    assert rel_err(views["gw"].view(128, 256).cpu(), (g.double().t() @ xg.double()).cpu()) < 1e-5
    kc = views["kc"].view(3, 9, 256)
    assert torch.equal(kc[:, 4], views["knew"].view(3, 256)) and torch.equal(kc[:, 5:], snapshot_view(snapshot, sizes, guard, "kc").view(3, 9, 256)[:, 5:])


def snapshot_view(snapshot, sizes, guard, name):
    off = 0
    for k, numel in sizes.items():
        off += guard
        if k == name:
            return snapshot[off:off + numel]
        off += numel + guard
    raise KeyError(name)
