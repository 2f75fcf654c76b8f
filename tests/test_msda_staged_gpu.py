"""The shared-memory staged kernels (csrc/msda_forward_staged.cu, msda_backward_staged.cu): coarse pyramid levels
copied into shared memory by the TMA, persistent grid.  In production they take over above a problem-size threshold; here
the threshold is lowered through cape_set_tuning so small oracle-sized cases run through them too.

Checked against the C / numpy oracles (fp32 1e-5 forward, 1e-4 gradients; bf16 / fp16 2e-2), against the L1 kernels
(same arithmetic, same order: bit-identical forward), and once at the full bench shape (N=20, Lq=S=5440)."""
import numpy as np
import pytest
import torch

import cape_b200
from cape_b200 import _lib, synthetic
from oracle import msda_c, msda_numpy
from tests.conftest import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture
def staged():
    """Force the staged kernels for any size; restore the defaults afterwards."""
    knobs = ("FWD_STAGED", "FWD_STAGED_MIN_QM", "FWD_STAGED_KB", "BWD_MODE", "BWD_STAGED_KB", "BWD_TC_MIN_QM")
    _lib.set_tuning("FWD_STAGED", 1)
    _lib.set_tuning("FWD_STAGED_MIN_QM", 1)
    _lib.set_tuning("BWD_TC_MIN_QM", 1)
    yield _lib.set_tuning
    for k in knobs:
        _lib.set_tuning(k, 0)


def _fwd(inp, dtype=torch.float32, aux=None):
    aux = aux or dtype
    out = cape_b200.ms_deform_attn(inp["value"].cuda().to(dtype), inp["spatial_shapes"].cuda(),
                                   inp["level_start_index"].cuda(), inp["sampling_locations"].cuda().to(aux),
                                   inp["attention_weights"].cuda().to(aux))
    torch.cuda.synchronize()
    return out


def _oracle_args(inp):
    return tuple(inp[k].numpy() for k in ("value", "spatial_shapes", "level_start_index", "sampling_locations",
                                          "attention_weights"))


@pytest.mark.parametrize("budget_kb", [8, 48, 200])           # level 3 only / levels 2-3 / levels 1-3 in shared memory
@pytest.mark.parametrize("dist", ["encoder", "uniform"])
@pytest.mark.parametrize("n,lq", [(2, 700), (3, 129), (1, 5440)])
def test_forward_staged_fp32_vs_c_oracle_and_l1_kernel(staged, budget_kb, dist, n, lq):
    inp = synthetic.make_inputs(n, lq, dist=dist, seed=n * 7 + lq)
    want = msda_c.msda_forward(*_oracle_args(inp), dtype=np.float32)
    staged("FWD_STAGED_KB", budget_kb)
    before = cape_b200.launch_count()
    got = _fwd(inp)
    assert cape_b200.launch_count() == before + 1
    assert rel_err(got.cpu().numpy(), want) < 1e-5
    staged("FWD_STAGED", 2)                                     # the L1 kernel: same arithmetic (FMA contraction may differ)
    assert torch.allclose(_fwd(inp), got, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("shapes,m", [(synthetic.CAPE_PYRAMID_512, 8), (((16, 12), (8, 6), (4, 3), (2, 2)), 3),
                                      (((40, 40), (3, 50), (7, 7), (1, 1)), 5)])
def test_forward_staged_other_pyramids_and_head_counts(staged, shapes, m):
    """1360-token pyramid (every level fits shared memory), non-square levels, M != 8."""
    inp = synthetic.make_inputs(2, 333, shapes, n_heads=m, dist="uniform", seed=m)
    want = msda_c.msda_forward(*_oracle_args(inp), dtype=np.float32)
    assert rel_err(_fwd(inp).cpu().numpy(), want) < 1e-5


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("aux_fp32", [True, False])
def test_forward_staged_half_precision(staged, dtype, aux_fp32):
    inp = synthetic.make_inputs(2, 600, dist="encoder", seed=21)
    aux = torch.float32 if aux_fp32 else dtype
    rounded = dict(inp)
    rounded["value"] = inp["value"].to(dtype).float()
    rounded["sampling_locations"] = inp["sampling_locations"].to(aux).float()
    rounded["attention_weights"] = inp["attention_weights"].to(aux).float()
    want = msda_c.msda_forward(*_oracle_args(rounded), dtype=np.float32)
    got = _fwd(inp, dtype, aux)
    assert rel_err(got.float().cpu().numpy(), want) < 2e-2
    staged("FWD_STAGED", 2)
    assert torch.allclose(_fwd(inp, dtype, aux).float(), got.float(), rtol=2e-2, atol=2e-2)


def test_fused_prologue_staged(staged):
    """cape::ms_deform_attn_decode (softmax + ref + off / (W, H) inside the kernel) through the staged kernel."""
    g = torch.Generator().manual_seed(4)
    n, lq, m, l, p = 2, 450, 8, 4, 4
    inp = synthetic.make_inputs(n, lq, dist="encoder", seed=5)
    ref = torch.rand(n, lq, l, 2, generator=g)
    off = torch.randn(n, lq, m, l, p, 2, generator=g) * 3
    logits = torch.randn(n, lq, m, l * p, generator=g)
    want = msda_numpy.msda_decode(inp["value"].numpy(), inp["spatial_shapes"].numpy(), inp["level_start_index"].numpy(),
                                  ref.numpy(), off.numpy(), logits.numpy())
    got = cape_b200.ms_deform_attn_decode(inp["value"].cuda(), inp["spatial_shapes"].cuda(), inp["level_start_index"].cuda(),
                                          ref.cuda(), off.cuda(), logits.cuda())
    torch.cuda.synchronize()
    assert rel_err(got.cpu().numpy(), want) < 1e-5


def test_inconsistent_pyramid_reads_nothing_out_of_bounds(staged):
    """level_start_index / spatial_shapes that do not fit S (the reference asserts, deformable_transformer.py:94): the
    offending level contributes zeros instead of gathering outside the value tensor."""
    inp = synthetic.make_inputs(1, 200, dist="uniform", seed=9)
    bad = dict(inp)
    bad["level_start_index"] = torch.tensor([0, 4096, 5120, 5400])           # 5400 + 64 > 5440
    good = dict(inp)
    good["attention_weights"] = inp["attention_weights"].clone()
    good["attention_weights"][:, :, :, 3] = 0                                # the same result with level 3 switched off
    assert torch.allclose(_fwd(bad), _fwd(good), atol=1e-6)


def test_full_bench_shape_against_the_c_oracle(staged):
    """N = 20, Lq = S = 5440 (the shape bench.py measures): forward + backward against the OpenMP C oracle."""
    for k in ("FWD_STAGED_MIN_QM",):
        staged(k, 0)                                                         # production thresholds
    inp = synthetic.make_inputs(20, 5440, dist="encoder", seed=11)
    a = _oracle_args(inp)
    want = msda_c.msda_forward(*a, dtype=np.float32)
    want_g = msda_c.msda_backward(inp["grad_output"].numpy(), *a, dtype=np.float32)
    v = inp["value"].cuda().requires_grad_(True)
    loc = inp["sampling_locations"].cuda().requires_grad_(True)
    attn = inp["attention_weights"].cuda().requires_grad_(True)
    out = cape_b200.ms_deform_attn(v, inp["spatial_shapes"].cuda(), inp["level_start_index"].cuda(), loc, attn)
    gv, gl, ga = torch.autograd.grad(out, (v, loc, attn), inp["grad_output"].cuda())
    torch.cuda.synchronize()
    assert rel_err(out.detach().cpu().numpy(), want) < 1e-5
    assert rel_err(gv.cpu().numpy(), want_g[0]) < 1e-4
    assert rel_err(ga.cpu().numpy(), want_g[2]) < 1e-4
    assert rel_err(gl.cpu().numpy(), want_g[1]) < 1e-4


# ---- backward: staged value rows (+ tensor-core scatter of the coarse levels) ---------------------------------------------
def _fwd_bwd(inp, dtype=torch.float32, aux=None):
    aux = aux or dtype
    v = inp["value"].cuda().to(dtype).requires_grad_(True)
    loc = inp["sampling_locations"].cuda().to(aux).requires_grad_(True)
    attn = inp["attention_weights"].cuda().to(aux).requires_grad_(True)
    out = cape_b200.ms_deform_attn(v, inp["spatial_shapes"].cuda(), inp["level_start_index"].cuda(), loc, attn)
    gv, gl, ga = torch.autograd.grad(out, (v, loc, attn), inp["grad_output"].cuda().to(dtype))
    torch.cuda.synchronize()
    return tuple(t.detach().float().cpu().numpy() for t in (gv, gl, ga))


# 2: staged rows + tensor-core scatter of the two coarse levels; 3: staged rows, all REDs; 5: small CTAs, tensor-core scatter of
# the coarsest level (csrc/msda_backward_tc.cu)
@pytest.mark.parametrize("mode", [2, 3, 5])
@pytest.mark.parametrize("dist", ["encoder", "uniform"])
@pytest.mark.parametrize("n,lq", [(2, 700), (3, 129), (1, 5440), (2, 31)])
def test_backward_staged_fp32_vs_c_oracle(staged, mode, dist, n, lq):
    inp = synthetic.make_inputs(n, lq, dist=dist, seed=n * 11 + lq)
    want = msda_c.msda_backward(inp["grad_output"].numpy(), *_oracle_args(inp), dtype=np.float32)
    staged("FWD_STAGED", 2)
    staged("BWD_MODE", mode)
    before = cape_b200.launch_count()
    gv, gl, ga = _fwd_bwd(inp)
    assert cape_b200.launch_count() == before + 2
    assert rel_err(gv, want[0]) < 1e-4
    assert rel_err(gl, want[1]) < 1e-4
    assert rel_err(ga, want[2]) < 1e-4


@pytest.mark.parametrize("shapes,m", [(synthetic.CAPE_PYRAMID_512, 8), (((16, 12), (8, 6), (4, 3), (2, 2)), 3),
                                      (((40, 40), (3, 50), (7, 7), (1, 1)), 5), (((30, 30), (20, 20), (25, 20), (4, 4)), 2)])
@pytest.mark.parametrize("mode", [2, 5])
def test_backward_tensor_core_scatter_other_pyramids(staged, shapes, m, mode):
    """Small pyramids (two covered levels in one 128-row tile), non-square levels, M != 8, and a pyramid whose level 2
    (500 pixels) does not fit the 384 covered rows (only the last level goes to the tensor cores)."""
    inp = synthetic.make_inputs(2, 333, shapes, n_heads=m, dist="uniform", seed=m)
    want = msda_c.msda_backward(inp["grad_output"].numpy(), *_oracle_args(inp), dtype=np.float32)
    staged("BWD_MODE", mode)
    gv, gl, ga = _fwd_bwd(inp)
    for got, w in zip((gv, gl, ga), want):
        assert rel_err(got, w) < 1e-4


@pytest.mark.parametrize("shapes", [((20, 24), (10, 12), (5, 6)), ((12, 12), (8, 8)), ((9, 9), (16, 16), (4, 4), (7, 9)),
                                    ((16, 16), (12, 12))])
def test_small_cta_tensor_core_scatter_with_fewer_levels(staged, shapes):
    """BWD_MODE=5 with 3 and 2 pyramid levels, a last level of exactly 64 pixels / of 63 pixels, and a last level too large
    for the 64-row tile (144 pixels: it keeps its REDs, the tiles stay unused)."""
    inp = synthetic.make_inputs(2, 301, shapes, n_heads=4, dist="uniform", seed=len(shapes))
    want = msda_c.msda_backward(inp["grad_output"].numpy(), *_oracle_args(inp), dtype=np.float32)
    staged("BWD_MODE", 5)
    before = cape_b200.launch_count()
    gv, gl, ga = _fwd_bwd(inp)
    assert cape_b200.launch_count() == before + 2
    for got, w in zip((gv, gl, ga), want):
        assert rel_err(got, w) < 1e-4


@pytest.mark.parametrize("mode", [2, 5])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_backward_tensor_core_scatter_half_precision(staged, dtype, mode):
    inp = synthetic.make_inputs(2, 600, dist="encoder", seed=21)
    rounded = dict(inp)
    rounded["value"] = inp["value"].to(dtype).float()
    rounded["grad_output"] = inp["grad_output"].to(dtype).float()
    want = msda_c.msda_backward(rounded["grad_output"].numpy(), *_oracle_args(rounded), dtype=np.float32)
    staged("BWD_MODE", mode)
    gv, gl, ga = _fwd_bwd(inp, dtype, torch.float32)
    for got, w in zip((gv, gl, ga), want):
        assert rel_err(got, w) < 2e-2


@pytest.mark.parametrize("mode", [2, 5])
def test_fused_backward_tensor_core_scatter(staged, mode):
    """cape::ms_deform_attn_fused_backward (softmax / location prologue inside the kernel) in modes 2 / 5 vs the L1 kernel."""
    g = torch.Generator().manual_seed(4)
    n, lq, m, l, p = 2, 450, 8, 4, 4
    inp = synthetic.make_inputs(n, lq, dist="encoder", seed=5)
    ref = torch.rand(n, lq, l, 2, generator=g).cuda()
    off = (torch.randn(n, lq, m, l, p, 2, generator=g) * 3).cuda()
    logits = torch.randn(n, lq, m, l * p, generator=g).cuda()
    gout = inp["grad_output"].cuda()

    def run():
        v = inp["value"].cuda().requires_grad_(True)
        o, lg, r = off.clone().requires_grad_(True), logits.clone().requires_grad_(True), ref.clone().requires_grad_(True)
        out = cape_b200.ms_deform_attn_decode(v, inp["spatial_shapes"].cuda(), inp["level_start_index"].cuda(), r, o, lg)
        res = torch.autograd.grad(out, (v, o, lg, r), gout)
        torch.cuda.synchronize()
        return res
    staged("BWD_MODE", 1)
    want = run()
    staged("BWD_MODE", mode)
    got = run()
    for a, b in zip(got, want):
        assert rel_err(a.cpu().numpy(), b.cpu().numpy()) < 1e-4
