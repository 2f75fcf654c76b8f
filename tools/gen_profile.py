"""A few steps of AutoregressiveGenerator (6 + 6 layers, N=128) run eagerly, for a kernel launch list under ncu:
    ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "step/" --csv --log-file out.csv \
        python tools/gen_profile.py
Development tool."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cape_b200

dev = torch.device("cuda", 0)
n = int(os.environ.get("GEN_N", "128"))
spec = cape_b200.TokenizerSpec(num_bins=44, seq_len=101)
torch.manual_seed(0)
tr = cape_b200.DeformableTransformer(
    d_model=256, nhead=8, num_encoder_layers=6, num_decoder_layers=6, dim_feedforward=1024, dropout=0.1, poly_refine=True,
    return_intermediate_dec=True, aux_loss=True, num_feature_levels=4, query_pos_type="sine", vocab_size=spec.vocab_size,
    seq_len=spec.seq_len, pad_idx=spec.pad)
tr.attach_heads(*cape_b200.build_prediction_heads(256, 3, 6, True))
with torch.no_grad():
    for head in tr.decoder.class_embed:
        head.bias.copy_(torch.tensor([50.0, 0.0, 0.0]))
tr = tr.to(dev).eval()
pyr = cape_b200.synthetic.CAPE_PYRAMID
memory = torch.randn(n, 5440, 256, device=dev)
shapes = torch.tensor(pyr, device=dev)
enc_cache = {"memory": memory, "spatial_shapes": shapes,
             "level_start_index": cape_b200.level_start_index_from_shapes(shapes),
             "valid_ratios": torch.ones(n, 4, 2, device=dev),
             "mask_flatten": torch.zeros(n, 5440, dtype=torch.bool, device=dev), "src_flatten": memory}
query_embed = torch.randn(spec.seq_len, 2, device=dev)
sup = torch.randn(n, 17, 256, device=dev)
sup_mask = torch.zeros(n, 17, dtype=torch.bool, device=dev)
gen = cape_b200.AutoregressiveGenerator(tr, spec, n, dev, fused=os.environ.get("GEN_FUSED", "1") == "1")
with torch.no_grad():
    import time
    out = gen.generate(None, None, None, query_embed, sup, sup_mask, enc_cache=enc_cache)      # captures the graph
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = gen.generate(None, None, None, query_embed, sup, sup_mask, enc_cache=enc_cache)
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    gen.state.reset()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(100):
        gen.graph.replay()
    torch.cuda.synchronize()
    t_rep = (time.perf_counter() - t0) / 100
    print(f"generate() {t_all * 1e3:.1f} ms for {out['steps']} steps; graph replay {t_rep * 1e6:.0f} us per step")

    def timed(fn):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) * 1e3, r
    t_reset, _ = timed(lambda: gen.reset(memory, shapes, enc_cache["level_start_index"], sup, sup_mask,
                                         padding_mask=enc_cache["mask_flatten"]))
    t_collect, _ = timed(gen._collect)
    print(f"reset {t_reset:.1f} ms, collect {t_collect:.1f} ms")
    gen.state.reset()
    for _ in range(3):
        gen._run()
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_push("step")
    gen._run()
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
    # timing of the same step, eager and as a graph
    if gen.fused and gen._fprep is None:
        pass
    import time
    t0 = time.perf_counter()
    for _ in range(20):
        gen._run()
    torch.cuda.synchronize()
    print(f"eager step {(time.perf_counter() - t0) / 20 * 1e6:.0f} us")
print("ok")
