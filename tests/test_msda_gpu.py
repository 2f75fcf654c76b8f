"""Parity of the CUDA path (through the C ABI / torch.library ops) against the oracle and the golden fixtures.

Tolerances are BASELINE.json's: forward 1e-5 relative (max|a-b| / max|b|) in fp32, 2e-2 in bf16, gradients 1e-4
relative in fp32.  All tests need a GPU (`-m gpu`); /root/reference is never read here.
"""
import ctypes
import glob
import os
import threading

import numpy as np
import pytest
import torch

import cape_b200
from cape_b200 import _lib, synthetic
from oracle import msda_c, msda_numpy
from tests.conftest import GOLDEN, rel_err

pytestmark = pytest.mark.gpu

FWD_TOL_F32 = 1e-5
GRAD_TOL_F32 = 1e-4
FWD_TOL_BF16 = 2e-2

CORE_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "core_*.npz")))


def _cuda(x, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(x)).cuda()
    return t.to(dtype) if dtype is not None else t


def _run_fwd_bwd(inp, dtype=torch.float32, aux_dtype=None):
    """inp: dict of CPU torch tensors / numpy arrays.  Returns numpy (out, gv, gl, ga)."""
    get = (lambda k: inp[k] if isinstance(inp[k], torch.Tensor) else torch.from_numpy(inp[k]))
    aux_dtype = aux_dtype or dtype
    v = get("value").cuda().to(dtype).requires_grad_(True)
    loc = get("sampling_locations").cuda().to(aux_dtype).requires_grad_(True)
    attn = get("attention_weights").cuda().to(aux_dtype).requires_grad_(True)
    shapes = get("spatial_shapes").cuda()
    starts = get("level_start_index").cuda()
    out = cape_b200.ms_deform_attn(v, shapes, starts, loc, attn)
    gv, gl, ga = torch.autograd.grad(out, (v, loc, attn), get("grad_output").cuda().to(dtype))
    torch.cuda.synchronize()
    return tuple(t.detach().float().cpu().numpy() for t in (out, gv, gl, ga))


def test_library_is_loaded_and_counts_launches():
    assert cape_b200.library_available()
    before = cape_b200.launch_count()
    inp = synthetic.make_inputs(1, 4, ((4, 4), (2, 2), (1, 1), (1, 1)), device="cuda")
    cape_b200.ms_deform_attn(inp["value"], inp["spatial_shapes"], inp["level_start_index"],
                             inp["sampling_locations"], inp["attention_weights"])
    torch.cuda.synchronize()
    assert cape_b200.launch_count() == before + 1
    maps = open("/proc/self/maps").read()
    assert "libcape_msda.so" in maps


@pytest.mark.parametrize("case", CORE_CASES)
def test_golden_fixtures_fp32(case):
    """CUDA fp32 vs the reference's own fp32 outputs (fixtures made by oracle/make_golden.py)."""
    g = np.load(os.path.join(GOLDEN, case + ".npz"))
    out, gv, gl, ga = _run_fwd_bwd(g)
    assert rel_err(out, g["out32"]) < FWD_TOL_F32
    assert rel_err(gv, g["grad_value32"]) < GRAD_TOL_F32
    assert rel_err(gl, g["grad_loc32"]) < GRAD_TOL_F32
    assert rel_err(ga, g["grad_attn32"]) < GRAD_TOL_F32


@pytest.mark.parametrize("dist", ["encoder", "uniform"])
@pytest.mark.parametrize("n,lq", [(2, 5440), (1, 1000), (3, 200), (5, 1)])
def test_cape_pyramid_fp32_vs_c_oracle(dist, n, lq):
    inp = synthetic.make_inputs(n, lq, dist=dist, seed=n * 100 + lq)
    a = tuple(inp[k].numpy() for k in ("value", "spatial_shapes", "level_start_index", "sampling_locations",
                                       "attention_weights"))
    want_out = msda_c.msda_forward(*a, dtype=np.float32)
    want_g = msda_c.msda_backward(inp["grad_output"].numpy(), *a, dtype=np.float32)
    out, gv, gl, ga = _run_fwd_bwd(inp)
    assert rel_err(out, want_out) < FWD_TOL_F32
    assert rel_err(gv, want_g[0]) < GRAD_TOL_F32
    assert rel_err(gl, want_g[1]) < GRAD_TOL_F32
    assert rel_err(ga, want_g[2]) < GRAD_TOL_F32


@pytest.mark.parametrize("shapes", [synthetic.CAPE_PYRAMID_512, ((16, 12), (8, 6), (4, 3)), ((20, 20),),
                                    ((9, 5), (4, 4))])
def test_other_pyramids_fp32(shapes):
    """1360-token pyramid, non-square levels, 3 / 1 / 2 levels (L < 4 fast path)."""
    inp = synthetic.make_inputs(2, 77, shapes, dist="uniform", seed=len(shapes))
    a = tuple(inp[k].numpy() for k in ("value", "spatial_shapes", "level_start_index", "sampling_locations",
                                       "attention_weights"))
    want_out = msda_c.msda_forward(*a, dtype=np.float32)
    want_g = msda_c.msda_backward(inp["grad_output"].numpy(), *a, dtype=np.float32)
    out, gv, gl, ga = _run_fwd_bwd(inp)
    assert rel_err(out, want_out) < FWD_TOL_F32
    for got, want in zip((gv, gl, ga), want_g):
        assert rel_err(got, want) < GRAD_TOL_F32


@pytest.mark.parametrize("m,d,l,p", [(3, 16, 2, 3), (2, 64, 3, 8), (1, 4, 1, 1), (4, 32, 5, 4), (8, 32, 4, 2),
                                     (2, 256, 1, 2)])
def test_generic_dims_fp32(m, d, l, p):
    shapes = ((7, 5), (4, 3), (3, 3), (2, 2), (2, 1))[:l]
    inp = synthetic.make_inputs(2, 19, shapes, n_heads=m, head_dim=d, n_points=p, dist="uniform", seed=m + d)
    a = tuple(inp[k].numpy() for k in ("value", "spatial_shapes", "level_start_index", "sampling_locations",
                                       "attention_weights"))
    want_out = msda_c.msda_forward(*a, dtype=np.float32)
    want_g = msda_c.msda_backward(inp["grad_output"].numpy(), *a, dtype=np.float32)
    out, gv, gl, ga = _run_fwd_bwd(inp)
    assert rel_err(out, want_out) < FWD_TOL_F32
    for got, want in zip((gv, gl, ga), want_g):
        assert rel_err(got, want) < GRAD_TOL_F32


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("aux_fp32", [True, False])
def test_half_precision_vs_oracle_on_rounded_inputs(dtype, aux_fp32):
    """bf16 / fp16 value: the kernel against the fp32 oracle evaluated on the SAME rounded inputs (what autocast
    feeds the reference), tolerance 2e-2."""
    inp = synthetic.make_inputs(2, 600, dist="encoder", seed=21)
    aux = torch.float32 if aux_fp32 else dtype
    rounded = dict(inp)
    rounded["value"] = inp["value"].to(dtype).float()
    rounded["grad_output"] = inp["grad_output"].to(dtype).float()
    rounded["sampling_locations"] = inp["sampling_locations"].to(aux).float()
    rounded["attention_weights"] = inp["attention_weights"].to(aux).float()
    a = tuple(rounded[k].numpy() for k in ("value", "spatial_shapes", "level_start_index", "sampling_locations",
                                           "attention_weights"))
    want_out = msda_c.msda_forward(*a, dtype=np.float32)
    want_g = msda_c.msda_backward(rounded["grad_output"].numpy(), *a, dtype=np.float32)
    out, gv, gl, ga = _run_fwd_bwd(inp, dtype=dtype, aux_dtype=aux)
    assert rel_err(out, want_out) < FWD_TOL_BF16
    for got, want in zip((gv, gl, ga), want_g):
        assert rel_err(got, want) < FWD_TOL_BF16


def test_core_pytorch_drop_in_signature_and_list_shapes():
    inp = synthetic.make_inputs(2, 33, ((8, 8), (4, 4), (2, 2), (1, 1)), seed=5)
    want = msda_c.msda_forward(*(inp[k].numpy() for k in ("value", "spatial_shapes", "level_start_index",
                                                           "sampling_locations", "attention_weights")))
    v, loc, attn = (inp[k].cuda() for k in ("value", "sampling_locations", "attention_weights"))
    out_t = cape_b200.ms_deform_attn_core_pytorch(v, inp["spatial_shapes"].cuda(), loc, attn)
    out_l = cape_b200.ms_deform_attn_core_pytorch(v, [(8, 8), (4, 4), (2, 2), (1, 1)], loc, attn)
    out_f = cape_b200.MSDeformAttnFunction.apply(v, inp["spatial_shapes"].cuda(), inp["level_start_index"].cuda(),
                                                 loc, attn, 64)
    assert out_t.shape == (2, 33, 256) and out_t.is_contiguous()
    for o in (out_t, out_l, out_f):
        assert rel_err(o.cpu().numpy(), want) < FWD_TOL_F32


def test_non_contiguous_and_cpu_resident_metadata():
    inp = synthetic.make_inputs(2, 21, ((8, 8), (4, 4), (2, 2), (1, 1)), seed=6)
    want = msda_c.msda_forward(*(inp[k].numpy() for k in ("value", "spatial_shapes", "level_start_index",
                                                           "sampling_locations", "attention_weights")))
    v = inp["value"].cuda().permute(0, 2, 1, 3).contiguous().permute(0, 2, 1, 3)     # same values, strided view
    assert not v.is_contiguous()
    out = cape_b200.ms_deform_attn(v, inp["spatial_shapes"], inp["level_start_index"],   # metadata left on the CPU
                                   inp["sampling_locations"].cuda(), inp["attention_weights"].cuda())
    assert rel_err(out.cpu().numpy(), want) < FWD_TOL_F32


def test_empty_and_degenerate_inputs():
    shapes = torch.tensor([[4, 4], [2, 2], [1, 1], [1, 1]], device="cuda")
    starts = torch.tensor([0, 16, 20, 21], device="cuda")
    v = torch.randn(2, 22, 8, 32, device="cuda")
    out = cape_b200.ms_deform_attn(v, shapes, starts, torch.zeros(2, 0, 8, 4, 4, 2, device="cuda"),
                                   torch.zeros(2, 0, 8, 4, 4, device="cuda"))
    assert out.shape == (2, 0, 256)
    out = cape_b200.ms_deform_attn(v[:0], shapes, starts, torch.zeros(0, 3, 8, 4, 4, 2, device="cuda"),
                                   torch.zeros(0, 3, 8, 4, 4, device="cuda"))
    assert out.shape == (0, 3, 256)
    # every sample far outside the map, NaN / inf locations: zeros, not a fault
    loc = torch.full((2, 3, 8, 4, 4, 2), 7.5, device="cuda")
    loc[0, 0] = float("inf")
    loc[0, 1] = float("nan")
    loc[1, 2] = -1e30
    attn = torch.full((2, 3, 8, 4, 4), 1 / 16, device="cuda")
    out = cape_b200.ms_deform_attn(v, shapes, starts, loc, attn)
    torch.cuda.synchronize()
    assert float(out.abs().max()) == 0.0


@pytest.mark.parametrize("n,lq,shapes", [(2, 37, synthetic.CAPE_PYRAMID), (1, 1, synthetic.CAPE_PYRAMID),
                                         (3, 130, ((9, 5), (4, 4), (2, 3))), (2, 19, ((7, 5), (4, 3)))])
def test_guard_zones_stay_intact(n, lq, shapes):
    """compute-sanitizer is closed on this pool, so out-of-bounds accesses are hunted with guard zones: every tensor the
    kernels touch is carved out of one arena with NaN-filled guards on both sides.  After forward + backward (direct
    and fused ops) the guards must be bit-identical (no stray write) and every result finite (no stray read that
    mattered — a guard value reaching an output would be NaN)."""
    generic = len(shapes) == 2
    kw = dict(n_heads=3, head_dim=16, n_points=3) if generic else {}
    inp = synthetic.make_inputs(n, lq, shapes, dist="uniform", seed=lq, **kw)
    guard = 4096                                      # floats on each side
    tensors = {k: inp[k] for k in ("value", "sampling_locations", "attention_weights", "grad_output")}
    m, l, p = inp["sampling_locations"].shape[2:5]
    tensors["ref"] = torch.rand(n, lq, l, 2)
    tensors["logits"] = torch.randn(n, lq, m, l * p)
    total = sum(t.numel() + 2 * guard for t in tensors.values())
    arena = torch.full((total,), float("nan"), device="cuda")
    views, off = {}, 0
    for k, t in tensors.items():
        off += guard
        views[k] = arena[off:off + t.numel()].view(t.shape)
        views[k].copy_(t)
        off += t.numel() + guard
    snapshot = arena.clone()
    shapes_t, starts_t = inp["spatial_shapes"].cuda(), inp["level_start_index"].cuda()
    v = views["value"].requires_grad_(True)
    loc = views["sampling_locations"].requires_grad_(True)
    attn = views["attention_weights"].requires_grad_(True)
    out = cape_b200.ms_deform_attn(v, shapes_t, starts_t, loc, attn)
    grads = torch.autograd.grad(out, (v, loc, attn), views["grad_output"])
    ref = views["ref"].requires_grad_(True)
    logits = views["logits"].requires_grad_(True)
    out2 = cape_b200.ms_deform_attn_fused(v, shapes_t, starts_t, ref, loc, logits)   # loc reused as raw offsets
    grads2 = torch.autograd.grad(out2, (v, ref, loc, logits), views["grad_output"])
    torch.cuda.synchronize()
    assert torch.equal(torch.isnan(arena), torch.isnan(snapshot))
    same = (arena == snapshot) | torch.isnan(snapshot)
    assert bool(same.all())
    for t in (out, out2) + tuple(grads) + tuple(grads2):
        assert bool(torch.isfinite(t).all())


def test_seeded_fuzz_of_shapes_vs_c_oracle():
    """40 seeded random configurations — fast-path and generic dimensions, ragged query counts, tiny and lopsided maps,
    both location distributions, fp32 and bf16 value — against the C oracle."""
    rng = np.random.default_rng(20240607)
    for case in range(40):
        fast = case % 2 == 0
        l = int(rng.integers(1, 5 if fast else 7))
        m = int(rng.integers(1, 9))
        d, p = (32, 4) if fast else (int(rng.choice([4, 8, 16, 24, 32, 48, 64])), int(rng.integers(1, 9)))
        shapes = tuple((int(rng.integers(1, 20)), int(rng.integers(1, 20))) for _ in range(l))
        n, lq = int(rng.integers(1, 4)), int(rng.integers(1, 150))
        dist = "uniform" if case % 3 else "encoder"
        inp = synthetic.make_inputs(n, lq, shapes, n_heads=m, head_dim=d, n_points=p, dist=dist, seed=case)
        half = case % 5 == 0
        if half:
            inp["value"] = inp["value"].bfloat16().float()
            inp["grad_output"] = inp["grad_output"].bfloat16().float()
        a = tuple(inp[k].numpy() for k in ("value", "spatial_shapes", "level_start_index", "sampling_locations",
                                           "attention_weights"))
        want_out = msda_c.msda_forward(*a, dtype=np.float32)
        want_g = msda_c.msda_backward(inp["grad_output"].numpy(), *a, dtype=np.float32)
        got = _run_fwd_bwd(inp, dtype=torch.bfloat16 if half else torch.float32, aux_dtype=torch.float32)
        tol_f, tol_g = (FWD_TOL_BF16, FWD_TOL_BF16) if half else (FWD_TOL_F32, GRAD_TOL_F32)
        ctx = f"case {case}: n={n} lq={lq} m={m} d={d} l={l} p={p} shapes={shapes} {dist} half={half}"
        assert rel_err(got[0], want_out) < tol_f, ctx
        for g_, w_, name in zip(got[1:], want_g, ("grad_value", "grad_loc", "grad_attn")):
            assert rel_err(g_, w_) < tol_g, ctx + " " + name


def test_forward_is_bitwise_repeatable_and_backward_within_tolerance_across_runs():
    """The forward has no atomics: two runs are bit-identical.  grad_value is accumulated with fp32 atomics: run-to-run
    differences (summation order over up to ~1400 hits per row) stay an order of magnitude inside the 1e-4 tolerance."""
    inp = synthetic.make_inputs(4, 5440, dist="encoder", seed=99)
    a = _run_fwd_bwd(inp)
    b = _run_fwd_bwd(inp)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3])
    assert rel_err(a[1], b[1]) < 1e-5


def test_error_behaviour():
    v = torch.randn(1, 4, 1, 4, device="cuda")
    shapes = torch.tensor([[2, 2]], device="cuda")
    starts = torch.tensor([0], device="cuda")
    with pytest.raises((ValueError, RuntimeError)):
        cape_b200.ms_deform_attn(v, shapes, starts, torch.zeros(1, 2, 1, 1, 1, 3, device="cuda"),
                                 torch.zeros(1, 2, 1, 1, 1, device="cuda"))
    with pytest.raises((ValueError, RuntimeError)):
        cape_b200.ms_deform_attn(v, shapes, starts, torch.zeros(1, 2, 1, 1, 1, 2, device="cuda"),
                                 torch.zeros(1, 2, 1, 1, 2, device="cuda"))
    # raw ABI: host pointer where a device pointer is expected -> CAPE_ERR_NOT_DEVICE_PTR, with a message
    lib = _lib.load()
    host = np.zeros(64, dtype=np.float32)
    dims = _lib.Dims(1, 4, 1, 4, 1, 1, 1)
    hp = ctypes.c_void_p(host.ctypes.data)
    rc = lib.cape_msda_forward(hp, ctypes.c_void_p(shapes.data_ptr()), ctypes.c_void_p(starts.data_ptr()), hp, hp, hp,
                               ctypes.byref(dims), 0, 0, None)
    assert rc == -5 and b"host memory" in lib.cape_last_error() or rc == -5
    # misaligned device pointer
    base = torch.zeros(64, device="cuda")
    rc = lib.cape_msda_forward(ctypes.c_void_p(base.data_ptr() + 4), ctypes.c_void_p(shapes.data_ptr()),
                               ctypes.c_void_p(starts.data_ptr()), ctypes.c_void_p(base.data_ptr()),
                               ctypes.c_void_p(base.data_ptr()), ctypes.c_void_p(base.data_ptr()),
                               ctypes.byref(dims), 0, 0, None)
    assert rc == -4


def test_decode_variant_matches_prologue_plus_core():
    """cape::ms_deform_attn_decode == softmax + (ref + off/(W,H)) + core on the cached value
    (deformable_transformer.py:99-105,112), for Lq = 1 (one new token) and Lq = 3."""
    for b, k, seed in ((4, 1, 0), (2, 3, 1), (64, 1, 2)):
        g = torch.Generator().manual_seed(seed)
        shapes = synthetic.CAPE_PYRAMID
        s = sum(h * w for h, w in shapes)
        value = torch.randn(b, s, 8, 32, generator=g)
        ref = torch.rand(b, k, 4, 2, generator=g)
        off = torch.randn(b, k, 8, 4, 4, 2, generator=g) * 3
        logits = torch.randn(b, k, 8, 16, generator=g)
        want = msda_numpy.msda_decode(value.numpy(), np.array(shapes), synthetic.level_start_index(shapes),
                                      ref.numpy(), off.numpy(), logits.numpy(), dtype=np.float64)
        got = cape_b200.ms_deform_attn_decode(value.cuda(), torch.tensor(shapes).cuda(), None, ref.cuda(), off.cuda(),
                                              logits.cuda())
        assert got.shape == (b, k, 256)
        assert rel_err(got.cpu().numpy(), want) < FWD_TOL_F32
        got16 = cape_b200.ms_deform_attn_decode(value.cuda().bfloat16(), torch.tensor(shapes).cuda(), None, ref.cuda(),
                                                off.cuda(), logits.cuda())
        want16 = msda_numpy.msda_decode(value.bfloat16().float().numpy(), np.array(shapes),
                                        synthetic.level_start_index(shapes), ref.numpy(), off.numpy(), logits.numpy())
        assert rel_err(got16.float().cpu().numpy(), want16) < FWD_TOL_BF16


@pytest.mark.parametrize("dims", [(8, 32, 4, 4), (2, 32, 3, 4), (3, 16, 2, 3)])
def test_fused_op_gradients_match_the_composed_reference_form(dims):
    """cape::ms_deform_attn_decode is differentiable (fused softmax / location prologue, forward AND backward): its
    gradients w.r.t. value, reference_points, sampling_offsets and attention_logits must equal autograd through the
    materialised form softmax -> ref + off/(W,H) -> cape::ms_deform_attn (deformable_transformer.py:100-112).  The
    last case has dimensions outside the fused kernels and exercises the composed fallback."""
    m, d, l, p = dims
    shapes = ((12, 16), (6, 8), (3, 4), (2, 2))[:l]
    s = sum(h * w for h, w in shapes)
    g = torch.Generator().manual_seed(m * 10 + l)
    n, lq = 2, 37
    value = torch.randn(n, s, m, d, generator=g).cuda().requires_grad_(True)
    ref = torch.rand(n, lq, l, 2, generator=g).cuda().requires_grad_(True)
    off = (torch.randn(n, lq, m, l, p, 2, generator=g) * 2).cuda().requires_grad_(True)
    logits = torch.randn(n, lq, m, l * p, generator=g).cuda().requires_grad_(True)
    gout = torch.randn(n, lq, m * d, generator=g).cuda()
    shapes_t = torch.tensor(shapes).cuda()
    starts_t = cape_b200.level_start_index_from_shapes(shapes_t)
    out = cape_b200.ms_deform_attn_fused(value, shapes_t, starts_t, ref, off, logits)
    got = torch.autograd.grad(out, (value, ref, off, logits), gout)
    attn = torch.softmax(logits, -1).view(n, lq, m, l, p)
    wh = torch.stack([shapes_t[:, 1], shapes_t[:, 0]], -1).float()
    loc = ref[:, :, None, :, None, :] + off / wh[None, None, None, :, None, :]
    out2 = cape_b200.ms_deform_attn(value, shapes_t, starts_t, loc, attn)
    want = torch.autograd.grad(out2, (value, ref, off, logits), gout)
    assert rel_err(out.detach().cpu().numpy(), out2.detach().cpu().numpy()) < FWD_TOL_F32
    for a, b, name in zip(got, want, ("value", "reference_points", "sampling_offsets", "attention_logits")):
        assert rel_err(a.cpu().numpy(), b.cpu().numpy()) < GRAD_TOL_F32, name


def test_module_unfused_form_matches_fused():
    g = np.load(os.path.join(GOLDEN, "module_forward.npz"))
    mod = cape_b200.MSDeformAttn(int(g["d_model"]), int(g["n_levels"]), int(g["n_heads"]), int(g["n_points"])).cuda()
    mod.load_state_dict({k[len("param."):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("param.")})
    args = (_cuda(g["query"]), _cuda(g["reference_points"]), _cuda(g["input_flatten"]), _cuda(g["spatial_shapes"]),
            _cuda(g["level_start_index"]), _cuda(g["padding_mask"]))
    fused = mod(*args)
    mod.fuse_prologue = False
    plain = mod(*args)
    assert rel_err(plain.detach().cpu().numpy(), g["out"]) < FWD_TOL_F32
    assert rel_err(fused.detach().cpu().numpy(), plain.detach().cpu().numpy()) < FWD_TOL_F32


def test_decode_generic_dims():
    g = torch.Generator().manual_seed(3)
    shapes = ((5, 7), (3, 2))
    value = torch.randn(2, 41, 3, 16, generator=g)
    ref = torch.rand(2, 2, 2, 2, generator=g)
    off = torch.randn(2, 2, 3, 2, 3, 2, generator=g)
    logits = torch.randn(2, 2, 3, 6, generator=g)
    want = msda_numpy.msda_decode(value.numpy(), np.array(shapes), [0, 35], ref.numpy(), off.numpy(), logits.numpy())
    got = cape_b200.ms_deform_attn_decode(value.cuda(), shapes, None, ref.cuda(), off.cuda(), logits.cuda())
    assert rel_err(got.cpu().numpy(), want) < FWD_TOL_F32


def test_module_mirror_matches_reference_module_fixture():
    """MSDeformAttn mirror with the reference's weights: forward + every gradient vs the reference module's own
    outputs (deformable_transformer.py:76-114), 2-d and 4-d reference points, padding mask."""
    g = np.load(os.path.join(GOLDEN, "module_forward.npz"))
    mod = cape_b200.MSDeformAttn(int(g["d_model"]), int(g["n_levels"]), int(g["n_heads"]), int(g["n_points"])).cuda()
    mod.load_state_dict({k[len("param."):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("param.")})
    q = _cuda(g["query"]).requires_grad_(True)
    x = _cuda(g["input_flatten"]).requires_grad_(True)
    shapes, starts = _cuda(g["spatial_shapes"]), _cuda(g["level_start_index"])
    out = mod(q, _cuda(g["reference_points"]), x, shapes, starts, _cuda(g["padding_mask"]))
    assert rel_err(out.detach().cpu().numpy(), g["out"]) < FWD_TOL_F32
    params = dict(mod.named_parameters())
    grads = torch.autograd.grad(out, [q, x] + list(params.values()), _cuda(g["grad_output"]))
    assert rel_err(grads[0].cpu().numpy(), g["grad_query"]) < GRAD_TOL_F32
    assert rel_err(grads[1].cpu().numpy(), g["grad_input_flatten"]) < GRAD_TOL_F32
    for (name, _), gr in zip(params.items(), grads[2:]):
        assert rel_err(gr.cpu().numpy(), g["grad_param." + name]) < GRAD_TOL_F32, name
    out4 = mod(q, _cuda(g["reference_points4"]), x, shapes, starts, None)
    assert rel_err(out4.detach().cpu().numpy(), g["out_ref4"]) < FWD_TOL_F32
    # inference path: fused decode prologue, and use_cache served from an attached holder — same numbers
    mod.eval()
    mod.cache = cape_b200.ValueCache()
    with torch.no_grad():
        first = mod(q, _cuda(g["reference_points"]), x, shapes, starts, _cuda(g["padding_mask"]), use_cache=False)
        launches = cape_b200.launch_count()
        again = mod(q, _cuda(g["reference_points"]), torch.zeros_like(x), shapes, starts, _cuda(g["padding_mask"]),
                    use_cache=True)      # memory argument ignored: the cached projection is used
    assert cape_b200.launch_count() == launches + 1
    assert rel_err(first.cpu().numpy(), g["out"]) < FWD_TOL_F32
    assert torch.equal(first, again)


def test_autocast_runs_the_core_in_fp32_like_the_reference():
    inp = synthetic.make_inputs(1, 16, ((8, 8), (4, 4), (2, 2), (1, 1)), device="cuda", seed=9)
    with torch.autocast("cuda", dtype=torch.float16):
        out = cape_b200.ms_deform_attn(inp["value"].half(), inp["spatial_shapes"], inp["level_start_index"],
                                       inp["sampling_locations"], inp["attention_weights"])
    assert out.dtype == torch.float32


def test_fused_op_under_autocast_keeps_the_value_dtype_and_matches_fp32_math():
    torch.manual_seed(5)
    mod = cape_b200.MSDeformAttn(256, 4, 8, 4).cuda()
    with torch.no_grad():
        for prm in mod.parameters():
            prm.add_(torch.randn_like(prm) * 0.05)
    shapes = torch.tensor(((8, 8), (4, 4), (2, 2), (1, 1)), device="cuda")
    starts = cape_b200.level_start_index_from_shapes(shapes)
    q, src = torch.randn(2, 9, 256, device="cuda"), torch.randn(2, 85, 256, device="cuda")
    ref = torch.rand(2, 9, 4, 2, device="cuda")
    want = mod(q, ref, src, shapes, starts)
    with torch.autocast("cuda", dtype=torch.float16):
        got = mod(q, ref, src, shapes, starts)
        v16 = torch.randn(2, 85, 8, 32, device="cuda", dtype=torch.float16)
        raw = cape_b200.ms_deform_attn_fused(v16, shapes, starts, ref, torch.randn(2, 9, 8, 4, 4, 2, device="cuda"),
                                             torch.randn(2, 9, 8, 16, device="cuda"))
    assert got.dtype == torch.float16 and raw.dtype == torch.float16          # no fp32 copy of the value was made
    assert rel_err(got.detach().float().cpu().numpy(), want.detach().cpu().numpy()) < FWD_TOL_BF16


def test_backward_from_autograd_thread_and_side_stream():
    inp = synthetic.make_inputs(2, 64, ((8, 8), (4, 4), (2, 2), (1, 1)), seed=10)
    a = tuple(inp[k].numpy() for k in ("value", "spatial_shapes", "level_start_index", "sampling_locations",
                                       "attention_weights"))
    want = msda_c.msda_backward(inp["grad_output"].numpy(), *a, dtype=np.float32)
    results = {}

    def work():
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            v = inp["value"].cuda().requires_grad_(True)
            loc = inp["sampling_locations"].cuda().requires_grad_(True)
            attn = inp["attention_weights"].cuda().requires_grad_(True)
            out = cape_b200.ms_deform_attn(v, inp["spatial_shapes"].cuda(), inp["level_start_index"].cuda(), loc, attn)
            (out * inp["grad_output"].cuda()).sum().backward()
            s.synchronize()
            results["g"] = (v.grad.cpu().numpy(), loc.grad.cpu().numpy(), attn.grad.cpu().numpy())

    t = threading.Thread(target=work)
    t.start()
    t.join()
    for got, w in zip(results["g"], want):
        assert rel_err(got, w) < GRAD_TOL_F32


def test_full_size_properties_training_shape():
    """N=20, Lq=5440 (BASELINE.json's named shape): size-independent identities instead of an oracle run.
    out is linear in value and in attention_weights, so
        <out(v), g> == <v, grad_value(g)>          (adjoint)
        <out(v), g> == <attn, grad_attn(g)>        (Euler, degree 1 in attn)
        out(a*v1 + b*v2) == a*out(v1) + b*out(v2)  (linearity)
    and a slice is checked against the C oracle."""
    n, lq = 20, 5440
    inp = synthetic.make_inputs(n, lq, dist="encoder", seed=1234, device="cuda")
    v = inp["value"].requires_grad_(True)
    loc = inp["sampling_locations"].requires_grad_(True)
    attn = inp["attention_weights"].requires_grad_(True)
    g = inp["grad_output"]
    out = cape_b200.ms_deform_attn(v, inp["spatial_shapes"], inp["level_start_index"], loc, attn)
    gv, gl, ga = torch.autograd.grad(out, (v, loc, attn), g)
    lhs = (out.double() * g.double()).sum().item()
    assert abs(lhs - (v.detach().double() * gv.double()).sum().item()) < 1e-6 * abs(lhs) + 1e-3
    assert abs(lhs - (attn.detach().double() * ga.double()).sum().item()) < 1e-6 * abs(lhs) + 1e-3
    v2 = torch.randn_like(v)
    with torch.no_grad():
        o2 = cape_b200.ms_deform_attn(v2, inp["spatial_shapes"], inp["level_start_index"], loc, attn)
        o12 = cape_b200.ms_deform_attn(0.5 * v + 2.0 * v2, inp["spatial_shapes"], inp["level_start_index"], loc, attn)
    assert rel_err((0.5 * out.detach() + 2.0 * o2).cpu().numpy(), o12.cpu().numpy()) < 1e-5
    # last batch element against the oracle
    sl = slice(n - 1, n)
    a = (v.detach()[sl].cpu().numpy(), inp["spatial_shapes"].cpu().numpy(), inp["level_start_index"].cpu().numpy(),
         loc.detach()[sl].cpu().numpy(), attn.detach()[sl].cpu().numpy())
    want_out = msda_c.msda_forward(*a, dtype=np.float32)
    want_g = msda_c.msda_backward(g[sl].cpu().numpy(), *a, dtype=np.float32)
    assert rel_err(out.detach()[sl].cpu().numpy(), want_out) < FWD_TOL_F32
    assert rel_err(gv[sl].cpu().numpy(), want_g[0]) < GRAD_TOL_F32
    assert rel_err(gl[sl].cpu().numpy(), want_g[1]) < GRAD_TOL_F32
    assert rel_err(ga[sl].cpu().numpy(), want_g[2]) < GRAD_TOL_F32


def test_host_buffer_round_trip_abi():
    """cape_msda_forward_backward_host: pinned host buffers in, results back on the host."""
    lib = _lib.load()
    inp = synthetic.make_inputs(2, 300, dist="encoder", seed=77)
    names = ("value", "sampling_locations", "attention_weights", "grad_output")
    pinned = {k: inp[k].contiguous().pin_memory() for k in names}
    dims = _lib.Dims(2, 5440, 8, 32, 300, 4, 4)
    ws_bytes = lib.cape_msda_host_workspace_bytes(ctypes.byref(dims), 1)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
    out = torch.empty(2, 300, 256).pin_memory()
    gv = torch.empty(2, 5440, 8, 32).pin_memory()
    gl = torch.empty(2, 300, 8, 4, 4, 2).pin_memory()
    ga = torch.empty(2, 300, 8, 4, 4).pin_memory()
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    stream = torch.cuda.current_stream()
    rc = lib.cape_msda_forward_backward_host(p(pinned["value"]), p(inp["spatial_shapes"]), p(inp["level_start_index"]),
                                             p(pinned["sampling_locations"]), p(pinned["attention_weights"]),
                                             p(pinned["grad_output"]), p(out), p(gv), p(gl), p(ga), ctypes.byref(dims),
                                             p(ws), ws_bytes, ctypes.c_void_p(stream.cuda_stream))
    _lib.check(rc, "cape_msda_forward_backward_host")
    stream.synchronize()
    a = tuple(inp[k].numpy() for k in ("value", "spatial_shapes", "level_start_index", "sampling_locations",
                                       "attention_weights"))
    assert rel_err(out.numpy(), msda_c.msda_forward(*a, dtype=np.float32)) < FWD_TOL_F32
    want_g = msda_c.msda_backward(inp["grad_output"].numpy(), *a, dtype=np.float32)
    for got, want in zip((gv, gl, ga), want_g):
        assert rel_err(got.numpy(), want) < GRAD_TOL_F32
    # too-small workspace is refused
    rc = lib.cape_msda_forward_backward_host(p(pinned["value"]), p(inp["spatial_shapes"]), p(inp["level_start_index"]),
                                             p(pinned["sampling_locations"]), p(pinned["attention_weights"]),
                                             p(pinned["grad_output"]), p(out), p(gv), p(gl), p(ga), ctypes.byref(dims),
                                             p(ws), 1024, ctypes.c_void_p(stream.cuda_stream))
    assert rc == -6


def test_host_buffer_round_trip_pipelined_chunks():
    """N=6 at the CAPE shape exceeds the 32 MB threshold: the call splits the batch into ragged chunks (2,2,2) on helper
    streams; results must equal the device-resident op's."""
    lib = _lib.load()
    n, lq = 6, 5440
    inp = synthetic.make_inputs(n, lq, dist="encoder", seed=78)
    pinned = {k: inp[k].contiguous().pin_memory() for k in ("value", "sampling_locations", "attention_weights",
                                                             "grad_output")}
    dims = _lib.Dims(n, 5440, 8, 32, lq, 4, 4)
    ws_bytes = lib.cape_msda_host_workspace_bytes(ctypes.byref(dims), 1)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
    res = [torch.empty(s).pin_memory() for s in ((n, lq, 256), (n, 5440, 8, 32), (n, lq, 8, 4, 4, 2), (n, lq, 8, 4, 4))]
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    stream = torch.cuda.Stream()
    for _ in range(2):            # second call reuses the helper streams
        for r in res:
            r.fill_(float("nan"))
        rc = lib.cape_msda_forward_backward_host(
            p(pinned["value"]), p(inp["spatial_shapes"]), p(inp["level_start_index"]), p(pinned["sampling_locations"]),
            p(pinned["attention_weights"]), p(pinned["grad_output"]), p(res[0]), p(res[1]), p(res[2]), p(res[3]),
            ctypes.byref(dims), p(ws), ws_bytes, ctypes.c_void_p(stream.cuda_stream))
        _lib.check(rc, "cape_msda_forward_backward_host")
        stream.synchronize()      # only the caller's stream is synchronised
        want = _run_fwd_bwd(inp)
        assert rel_err(res[0].numpy(), want[0]) == 0.0
        assert rel_err(res[1].numpy(), want[1]) < 1e-5      # atomics: summation order differs run to run
        assert rel_err(res[2].numpy(), want[2]) == 0.0
        assert rel_err(res[3].numpy(), want[3]) == 0.0


def _layer_kwargs(g):
    return dict(d_model=int(g["d_model"]), d_ffn=int(g["d_ffn"]), dropout=0.0, n_levels=int(g["n_levels"]),
                n_heads=int(g["n_heads"]), n_points=int(g["n_points"]))


def _load_params(module, g):
    module.load_state_dict({k[len("param."):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("param.")})
    return module.cuda()


def test_encoder_stack_mirror_matches_reference():
    """2-layer DeformableTransformerEncoder (deformable_transformer.py:155-291) with the reference's weights."""
    g = np.load(os.path.join(GOLDEN, "encoder_stack.npz"))
    enc = _load_params(cape_b200.DeformableTransformerEncoder(
        cape_b200.DeformableTransformerEncoderLayer(**_layer_kwargs(g)), 2), g)
    x = _cuda(g["src"]).requires_grad_(True)
    out = enc(x, _cuda(g["spatial_shapes"]), _cuda(g["level_start_index"]), _cuda(g["valid_ratios"]), _cuda(g["pos"]))
    assert rel_err(out.detach().cpu().numpy(), g["out"]) < FWD_TOL_F32
    gx, = torch.autograd.grad(out, x, _cuda(g["grad_output"]))
    assert rel_err(gx.cpu().numpy(), g["grad_src"]) < GRAD_TOL_F32


def test_encoder_stack_with_tensor_core_linears_stays_at_fp32_level():
    """CAPE-width encoder (d_model 256, FFN 1024, 2 layers) at inference through the opt-in 3xTF32 linears vs the same
    stack on nn.Linear: the sampling op is unchanged, the 12 projections each carry ~2e-6 instead of ~5e-7."""
    enc = cape_b200.DeformableTransformerEncoder(
        cape_b200.DeformableTransformerEncoderLayer(256, 1024, 0.0, "relu", 4, 8, 4), 2)
    synthetic.fill_parameters_(enc, seed=5)
    enc = enc.cuda().eval()
    shapes = ((16, 12), (8, 6), (4, 3), (2, 2))
    s = sum(h * w for h, w in shapes)
    n = 3
    src = torch.from_numpy(synthetic.seeded_array("src", (n, s, 256), 5)).cuda()
    pos = torch.from_numpy(synthetic.seeded_array("pos", (n, s, 256), 5)).cuda() * 0.5
    args = (src, torch.tensor(shapes, device="cuda"), torch.tensor(synthetic.level_start_index(shapes), device="cuda"),
            torch.ones(n, 4, 2, device="cuda"), pos)
    with torch.no_grad():
        want = enc(*args)
        old = cape_b200.set_linear_mode("tf32x3")
        try:
            enc(*args)                                  # first call also splits the 12 weights (cached afterwards)
            before = cape_b200.launch_count()
            out = enc(*args)
            launches = cape_b200.launch_count() - before
        finally:
            cape_b200.set_linear_mode(old)
    assert launches == 2 * (1 + 6)                      # per layer: the sampling kernel + 6 tensor-core linears
    assert rel_err(out.cpu().numpy(), want.cpu().numpy()) < 2e-5


def test_encoder_layer_training_step_with_tensor_core_linears_matches_fp32_gradients():
    """One CAPE-width encoder layer, forward + backward with autograd, opt-in 3xTF32 linears (forward, input gradient and
    weight gradient on the tensor cores) vs nn.Linear: outputs and every parameter gradient at fp32 level."""
    layer = cape_b200.DeformableTransformerEncoderLayer(256, 1024, 0.0, "relu", 4, 8, 4)
    synthetic.fill_parameters_(layer, seed=9)
    layer = layer.cuda()
    shapes = ((32, 24), (16, 12), (8, 6), (4, 3))
    s = sum(h * w for h, w in shapes)                       # 1020 tokens x 2 images = 2040 rows (>= 1024, % 32 != 0 -> pad check)
    n = 2
    src = torch.from_numpy(synthetic.seeded_array("src", (n, s, 256), 9)).cuda()
    pos = torch.from_numpy(synthetic.seeded_array("pos", (n, s, 256), 9)).cuda() * 0.5
    shapes_t = torch.tensor(shapes, device="cuda")
    starts = torch.tensor(synthetic.level_start_index(shapes), device="cuda")
    ref = cape_b200.DeformableTransformerEncoder.get_reference_points(shapes_t, torch.ones(n, 4, 2, device="cuda"), "cuda")
    gout = torch.from_numpy(synthetic.seeded_array("gout", (n, s, 256), 9)).cuda()

    def run():
        x = src.clone().requires_grad_(True)
        out = layer(x, pos, ref, shapes_t, starts, None)
        grads = torch.autograd.grad(out, [x] + list(layer.parameters()), gout)
        return out.detach(), grads

    want_out, want = run()
    old = cape_b200.set_linear_mode("tf32x3")
    try:
        before = cape_b200.launch_count()
        got_out, got = run()
        assert cape_b200.launch_count() - before >= 2 + 6 * 2          # the sampling pair + tensor-core forward / dgrad calls
    finally:
        cape_b200.set_linear_mode(old)
    assert rel_err(got_out.cpu().numpy(), want_out.cpu().numpy()) < 2e-5
    names = ["src"] + [k for k, _ in layer.named_parameters()]
    errs = {k: rel_err(a.cpu().numpy(), b.cpu().numpy()) for k, a, b in zip(names, got, want)}
    assert max(errs.values()) < 1e-4, errs


def test_decoder_layer_mirror_teacher_forced_and_incremental():
    """TransformerDecoderLayer v1 (deformable_transformer_v2.py:262-370): teacher-forced forward/backward, then the same
    sequence token by token with KV cache + projected-value cache, then one decode step replayed from a CUDA graph."""
    g = np.load(os.path.join(GOLDEN, "decoder_layer.npz"))
    layer = _load_params(cape_b200.TransformerDecoderLayer(**_layer_kwargs(g)), g).eval()
    shapes, starts = _cuda(g["spatial_shapes"]), _cuda(g["level_start_index"])
    sup, sup_mask = _cuda(g["support_features"]), _cuda(g["support_mask"])
    qpos, refp = _cuda(g["query_pos"]), _cuda(g["reference_points"])
    tgt = _cuda(g["tgt"]).requires_grad_(True)
    mem = _cuda(g["memory"]).requires_grad_(True)
    out, _ = layer(tgt, qpos, refp, mem, shapes, starts, None, _cuda(g["causal_mask"]), support_features=sup,
                   support_mask=sup_mask)
    assert rel_err(out.detach().cpu().numpy(), g["out_teacher_forced"]) < FWD_TOL_F32
    g_t, g_m = torch.autograd.grad(out, (tgt, mem), _cuda(g["grad_output"]))
    assert rel_err(g_t.cpu().numpy(), g["grad_tgt"]) < GRAD_TOL_F32
    assert rel_err(g_m.cpu().numpy(), g["grad_memory"]) < GRAD_TOL_F32

    n, t_len = g["tgt"].shape[:2]
    layer.setup_caches(n, t_len, device="cuda")
    tgt, mem = tgt.detach(), mem.detach()
    steps = []
    with torch.no_grad():
        for i in range(t_len):
            launches = cape_b200.launch_count()
            o, _ = layer(tgt[:, i:i + 1], qpos[:, i:i + 1], refp[:, i:i + 1], mem, shapes, starts, None,
                         torch.zeros(1, i + 1, device="cuda"), input_pos=i, support_features=sup, support_mask=sup_mask)
            assert cape_b200.launch_count() == launches + 1          # one fused sampling kernel per step
            steps.append(o)
    inc = torch.cat(steps, 1)
    assert rel_err(inc.cpu().numpy(), g["out_incremental"]) < FWD_TOL_F32
    assert layer.cross_attn.cache.get().shape == (n, g["memory"].shape[1], int(g["n_heads"]), 32)

    # CUDA graph: capture the last decode step (static shapes, cached value) and replay it on new token features
    i = t_len - 1
    static_tgt = tgt[:, i:i + 1].clone()
    mask = torch.zeros(1, i + 1, device="cuda")
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side), torch.no_grad():
        for _ in range(2):          # warm-up outside capture
            layer(static_tgt, qpos[:, i:i + 1], refp[:, i:i + 1], mem, shapes, starts, None, mask, input_pos=i,
                  support_features=sup, support_mask=sup_mask)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph), torch.no_grad():
        static_out, _ = layer(static_tgt, qpos[:, i:i + 1], refp[:, i:i + 1], mem, shapes, starts, None, mask,
                              input_pos=i, support_features=sup, support_mask=sup_mask)
    static_tgt.copy_(tgt[:, i:i + 1] * 0.5)
    graph.replay()
    with torch.no_grad():
        eager, _ = layer(tgt[:, i:i + 1] * 0.5, qpos[:, i:i + 1], refp[:, i:i + 1], mem, shapes, starts, None, mask,
                         input_pos=i, support_features=sup, support_mask=sup_mask)
    torch.cuda.synchronize()
    assert torch.allclose(static_out, eager, atol=1e-6, rtol=1e-5)


def test_incremental_decoder_one_graph_per_step_matches_reference_incremental_outputs():
    """IncrementalDecoder (static shapes, whole step in one CUDA graph) against the reference layer's own token-by-token
    outputs (tests/golden/decoder_layer.npz, produced with the reference's KVCache), eager and graph-replayed, and a
    2-layer stack against the eager mirror."""
    g = np.load(os.path.join(GOLDEN, "decoder_layer.npz"))
    layer = _load_params(cape_b200.TransformerDecoderLayer(**_layer_kwargs(g)), g).eval()
    shapes, starts = _cuda(g["spatial_shapes"]), _cuda(g["level_start_index"])
    sup, sup_mask = _cuda(g["support_features"]), _cuda(g["support_mask"])
    qpos, refp, tgt, mem = _cuda(g["query_pos"]), _cuda(g["reference_points"]), _cuda(g["tgt"]), _cuda(g["memory"])
    n, t_len = tgt.shape[:2]
    for use_graph in (False, True):
        dec = cape_b200.IncrementalDecoder([layer], n, t_len + 3, "cuda")     # max_len > sequence: masked tail
        dec.reset(mem, shapes, starts, sup, sup_mask)
        steps = []
        for i in range(t_len):
            before = cape_b200.launch_count()
            o = dec.step(i, tgt[:, i:i + 1], qpos[:, i:i + 1], refp[:, i:i + 1], use_graph=use_graph)
            steps.append(o.clone())
            if not use_graph:
                assert cape_b200.launch_count() == before + 1
        inc = torch.cat(steps, 1)
        assert rel_err(inc.cpu().numpy(), g["out_incremental"]) < FWD_TOL_F32, use_graph
    # two layers: graph replay == eager mirror layers driven with python-int positions
    layer2 = _load_params(cape_b200.TransformerDecoderLayer(**_layer_kwargs(g)), g).eval()
    with torch.no_grad():
        for prm in layer2.parameters():
            prm.mul_(1.1)
    stack = [layer, layer2]
    dec = cape_b200.IncrementalDecoder(stack, n, t_len, "cuda")
    dec.reset(mem, shapes, starts, sup, sup_mask)
    for l in stack:
        l.setup_caches(n, t_len, device="cuda")
    with torch.no_grad():
        for i in range(t_len):
            got = dec.step(i, tgt[:, i:i + 1], qpos[:, i:i + 1], refp[:, i:i + 1])
            x = tgt[:, i:i + 1]
            for l in stack:
                x, _ = l(x, qpos[:, i:i + 1], refp[:, i:i + 1], mem, shapes, starts, None,
                         torch.zeros(1, i + 1, device="cuda"), input_pos=i, support_features=sup, support_mask=sup_mask)
            assert torch.allclose(got, x, atol=2e-5, rtol=1e-4), i


def test_nccl_flat_grad_allreduce_two_gpus():
    """The one collective a data-parallel CAPE step needs, over NCCL (needs >= 2 GPUs; the CPU suite covers gloo)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import subprocess
    import sys
    code = (
        "import os, torch, torch.distributed as dist, sys\n"
        "sys.path.insert(0, os.environ['CAPE_REPO'])\n"
        "from cape_b200 import dist as cdist\n"
        "r, w, lr = cdist.init_from_env('nccl')\n"
        "dev = torch.device('cuda', lr)\n"
        "torch.manual_seed(0)\n"
        "m = torch.nn.Linear(8, 4).to(dev)\n"
        "unused = torch.nn.Parameter(torch.ones(3, device=dev))\n"
        "m(torch.full((2, 8), float(r + 1), device=dev)).sum().backward()\n"
        "local = m.weight.grad.clone()\n"
        "cdist.FlatGradAllreduce(list(m.parameters()) + [unused])()\n"
        "want = local.clone(); dist.all_reduce(want); want /= w\n"
        "assert torch.allclose(m.weight.grad, want) and unused.grad is None\n"
        "assert cdist.max_over_ranks(float(r), dev) == w - 1\n"
        "cdist.barrier(dev); dist.destroy_process_group(); print('rank', r, 'ok')\n")
    import tempfile
    env = dict(os.environ, CAPE_REPO=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    with tempfile.NamedTemporaryFile("w", suffix="_nccl_worker.py", delete=False) as f:
        f.write(code)
        script = f.name
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", script], env=env,
                         capture_output=True, text=True, timeout=300)
    os.unlink(script)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("ok") == 2


def test_nccl_grad_buckets_overlapped_two_gpus():
    """GradBuckets over NCCL: per-bucket all-reduces launched from inside the last backward of an accumulation window
    (async on NCCL's stream, joined before the optimizer), through the reference-shaped loop of tests/test_dist_gloo.py.
    Both ranks must end with the weights of one process that saw both ranks' batches."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import subprocess
    import sys
    code = (
        "import os, torch, torch.distributed as dist, sys\n"
        "sys.path.insert(0, os.environ['CAPE_REPO'])\n"
        "from cape_b200 import dist as cdist\n"
        "from tests.test_dist_gloo import _toy_model, _toy_batches, _reference_shaped_loop\n"
        "r, w, lr = cdist.init_from_env('nccl')\n"
        "dev = torch.device('cuda', lr)\n"
        "model = _toy_model().to(dev)\n"
        "opt = torch.optim.AdamW(model.parameters(), lr=0.05, weight_decay=0.1)\n"
        "buckets = cdist.GradBuckets(model.parameters(), bucket_bytes=64)\n"
        "move = lambda b: {'x': b['x'].to(dev), 'nested': {'y': b['nested']['y'].to(dev)}, 'tag': b['tag']}\n"
        "loader = [move(b) for b in _toy_batches(7)]\n"
        "def engine(model, criterion, loader, optimizer, device, epoch, max_norm=0, accumulation_steps=1, scaler=None):\n"
        "    _reference_shaped_loop(model, loader, optimizer, accumulation_steps, max_norm)\n"
        "cdist.train_one_epoch_data_parallel(engine, model, None, loader, opt, dev, 0, buckets, accumulation_steps=3,\n"
        "                                    max_norm=0.5, queries_per_episode=2)\n"
        "torch.cuda.synchronize()\n"
        "ref = _toy_model(); ropt = torch.optim.AdamW(ref.parameters(), lr=0.05, weight_decay=0.1)\n"
        "_reference_shaped_loop(ref, _toy_batches(7), ropt, 3, 0.5)\n"
        "for (k, a), (_, b) in zip(model.state_dict().items(), ref.state_dict().items()):\n"
        "    assert torch.allclose(a.cpu(), b, atol=1e-5), k\n"
        "assert buckets.known and len(buckets.buckets) > 1 and model.never_used.grad is None\n"
        "cdist.barrier(dev); dist.destroy_process_group(); print('rank', r, 'ok')\n")
    import tempfile
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, CAPE_REPO=repo, PYTHONPATH=repo + os.pathsep + os.environ.get("PYTHONPATH", ""))
    with tempfile.NamedTemporaryFile("w", suffix="_nccl_buckets_worker.py", delete=False) as f:
        f.write(code)
        script = f.name
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29534", script], env=env,
                         capture_output=True, text=True, timeout=300)
    os.unlink(script)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("ok") == 2


def test_opcheck_registration():
    inp = synthetic.make_inputs(1, 5, ((4, 4), (2, 2), (1, 1), (1, 1)), device="cuda", seed=3)
    args = (inp["value"].requires_grad_(True), inp["spatial_shapes"], inp["level_start_index"],
            inp["sampling_locations"].requires_grad_(True), inp["attention_weights"].requires_grad_(True))
    torch.library.opcheck(torch.ops.cape.ms_deform_attn.default, args,
                          test_utils=("test_schema", "test_faketensor", "test_autograd_registration"))


def test_masked_fill_rows_matches_masked_fill_forward_and_backward():
    """deformable_transformer.py:96-97 done in place with the mask inspected on the device: all-False mask (CAPE's only
    case) leaves the tensor untouched; masks with True rows give masked_fill's result and gradient."""
    from cape_b200.ops import masked_fill_rows_
    g = torch.Generator().manual_seed(0)
    x = torch.randn(3, 700, 256, generator=g).cuda()
    w = torch.randn(256, 256, generator=g).cuda().requires_grad_(True)
    for frac in (0.0, 0.02, 1.0):
        mask = (torch.rand(3, 700, generator=g) < frac).cuda()
        before = cape_b200.launch_count()
        got = masked_fill_rows_(x @ w, mask)
        assert cape_b200.launch_count() == before + 1
        want = (x @ w).masked_fill(mask[..., None], 0.0)
        assert torch.equal(got, want)
        gg, = torch.autograd.grad((got * got).sum(), w)
        gw, = torch.autograd.grad((want * want).sum(), w)
        assert torch.allclose(gg, gw, rtol=1e-5, atol=1e-4)
    half = torch.randn(2, 300, 256, generator=g).cuda().half()
    mask = (torch.rand(2, 300, generator=g) < 0.3).cuda()
    assert torch.equal(masked_fill_rows_(half.clone(), mask), half.masked_fill(mask[..., None], 0.0))
    leaf = torch.randn(2, 5, 32, device="cuda", requires_grad=True)          # a leaf is never modified in place
    out = masked_fill_rows_(leaf, torch.ones(2, 5, dtype=torch.bool, device="cuda"))
    assert out is not leaf and float(out.abs().sum()) == 0.0 and float(leaf.abs().sum()) > 0


@pytest.mark.parametrize("generic", [False, True])
def test_inconsistent_pyramid_is_skipped_not_read_out_of_bounds(generic):
    """level_start_index / spatial_shapes that do not fit S.  The reference asserts (deformable_transformer.py:94);
    the kernels cannot raise, so the offending level contributes nothing — forward and backward — instead of gathering /
    scattering outside value / grad_value (guard zones around the tensors stay intact)."""
    p = 3 if generic else 4                                              # P = 3 takes the generic kernels
    inp = synthetic.make_inputs(1, 300, n_points=p, dist="uniform", seed=9)
    bad_starts = torch.tensor([0, 4096, 5120, 5400])                     # 5400 + 8*8 > S = 5440

    def run(starts, attn):
        pad = 4096
        vbuf = torch.full((inp["value"].numel() + 2 * pad,), float("nan"), device="cuda")
        v = vbuf[pad:-pad].view(inp["value"].shape)
        v.copy_(inp["value"].cuda())
        v = v.detach().requires_grad_(True)
        loc = inp["sampling_locations"].cuda().requires_grad_(True)
        a = attn.cuda().requires_grad_(True)
        out = cape_b200.ms_deform_attn(v, inp["spatial_shapes"].cuda(), starts.cuda(), loc, a)
        gv, gl = torch.autograd.grad(out, (v, loc), inp["grad_output"].cuda())
        torch.cuda.synchronize()
        assert torch.isnan(vbuf[:pad]).all() and torch.isnan(vbuf[-pad:]).all()
        return out.detach(), gv, gl

    off = inp["attention_weights"].clone()
    off[:, :, :, 3] = 0                                                   # the same result with level 3 switched off
    out_bad, gv_bad, gl_bad = run(bad_starts, inp["attention_weights"])
    out_ok, gv_ok, gl_ok = run(inp["level_start_index"], off)
    assert torch.isfinite(out_bad).all() and torch.allclose(out_bad, out_ok, atol=1e-6)
    assert torch.allclose(gv_bad, gv_ok, atol=1e-6)
    assert torch.allclose(gl_bad[:, :, :, :3], gl_ok[:, :, :, :3], atol=1e-5)
    assert float(gl_bad[:, :, :, 3].abs().max()) == 0.0
