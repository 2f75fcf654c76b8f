"""Host logic around the unmodified reference model that needs no GPU: staging recipe, synthetic episodes vs the
reference's own tokenisation, binding the mirror transformer to a live model's tensors, and the feature-pyramid
prologue of the ``forward_inference`` drop-in vs what the reference hands its transformer."""
import os
import sys
import types

import numpy as np
import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "tools"))
import stage_reference  # noqa: E402

pytestmark = pytest.mark.skipif(not stage_reference.available(), reason="reference not available on this machine")


@pytest.fixture(scope="module")
def model():
    m, _, _, _ = stage_reference.build_cape_model("cpu", seed=0)
    return m.eval()


def test_staged_tree_is_ignored_by_git_and_complete():
    root = stage_reference.root()
    for rel in ("models/deformable_transformer.py", "models/cape_model.py", "models/engine_cape.py", "util/misc.py",
                "datasets/discrete_tokenizer.py"):
        assert os.path.exists(os.path.join(root, rel)), rel
    with open(os.path.join(REPO, ".gitignore")) as f:
        assert "baseline/_ref/" in f.read()
    ignore = os.path.join(REPO, ".gpurunignore")
    if os.path.exists(ignore):
        with open(ignore) as f:
            assert "baseline" not in f.read()              # the staged reference must travel to the GPU box


def test_synthetic_targets_equal_the_reference_tokenisation():
    import cape_b200
    stage_reference.activate()
    from datasets.discrete_tokenizer import DiscreteTokenizerV2
    from datasets.mp100_cape import MP100CAPE
    fake = types.SimpleNamespace(tokenizer=DiscreteTokenizerV2(44, 200, add_cls=False))
    rng = np.random.RandomState(0)
    for k in (1, 9, 17, 100):
        kp = rng.uniform(0, 1, size=(k, 2))
        kp[0] = [1.0, 0.0]                                   # clamp edge: ceil(43) stays inside the vocabulary
        want = MP100CAPE._tokenize_keypoints(fake, [[x * 512, y * 512] for x, y in kp], 512, 512)
        got = cape_b200.synthetic.tokenize_keypoints(kp, 44, 200)
        assert set(got) == set(want)
        for key in want:
            assert got[key].dtype == want[key].dtype and got[key].shape == want[key].shape, key
            assert torch.allclose(got[key].float(), want[key].float(), atol=1e-6), (k, key)
    batch = cape_b200.synthetic.make_episode_batch(3, 2, num_keypoints=9, shots=5, seed=1, image_size=64)
    assert batch["query_images"].shape == (6, 3, 64, 64) and batch["support_coords"].shape == (6, 9, 2)
    assert torch.equal(batch["support_coords"][0], batch["support_coords"][1])       # support repeated per query
    assert not batch["support_masks"].any() and len(batch["support_skeletons"]) == 6
    assert batch["query_targets"]["seq11"].shape == (6, 200)


def test_mirror_binds_to_the_live_tensors(model):
    import cape_b200
    from cape_b200.transformer import mirror_from_reference, mirror_is_current
    ref = model.base_model.transformer
    mirror = mirror_from_reference(ref)
    assert mirror_is_current(mirror, ref)
    ref_params = dict(ref.named_parameters())
    for name, p in mirror.named_parameters():
        assert not p.is_meta and p.data_ptr() == ref_params[name].data_ptr(), name
    assert mirror.decoder.class_embed is ref.decoder.class_embed
    with torch.no_grad():                                    # an optimizer-style in-place update is seen, no copy
        ref.level_embed.add_(1.0)
    assert torch.equal(mirror.level_embed, ref.level_embed)
    keys = list(model.state_dict().keys())
    model.base_model.transformer.float()                     # no-op cast keeps storage
    assert mirror_is_current(mirror, ref)
    ref.level_embed.data = ref.level_embed.data.clone()      # a re-allocation (model.to(...), assign-load) is detected
    assert not mirror_is_current(mirror, ref)
    assert list(model.state_dict().keys()) == keys


def test_feature_pyramid_prologue_equals_what_the_reference_hands_its_transformer(model):
    import cape_b200
    from cape_b200 import patch
    stage_reference.activate()
    rf = sys.modules["models.roomformer_v2"]
    from datasets.discrete_tokenizer import DiscreteTokenizerV2
    base = model.base_model
    saved_tok = base.tokenizer
    base.tokenizer = DiscreteTokenizerV2(44, 2, add_cls=False)      # two decode steps are enough to reach the call
    seen = {}

    def grab(module, args, kwargs):
        if "srcs" not in seen:
            seen["srcs"], seen["masks"], seen["pos"] = args[0], args[1], args[2]
    handle = base.transformer.register_forward_pre_hook(grab, with_kwargs=True)
    images = torch.rand(1, 3, 512, 512, generator=torch.Generator().manual_seed(0))
    try:
        with torch.no_grad():
            base.forward_inference(images, use_cache=True)
            srcs, masks, pos = patch._feature_pyramid(base, images, rf)
    finally:
        handle.remove()
        base.tokenizer = saved_tok
        for layer in base.transformer.decoder.layers:        # drop the cache modules the reference registered (A.2)
            layer.kv_cache = None
            if hasattr(layer.cross_attn, "cache"):
                del layer.cross_attn.cache
    assert [tuple(s.shape[-2:]) for s in srcs] == [(64, 64), (32, 32), (16, 16), (8, 8)]
    for a, b in zip(srcs + masks + pos, seen["srcs"] + seen["masks"] + seen["pos"]):
        assert torch.equal(a, b)


def test_patch_and_unpatch_forward_inference_swap():
    import cape_b200
    stage_reference.activate()
    import models.deformable_transformer as dt
    import models.roomformer_v2 as rf
    original = rf.RoomFormerV2.forward_inference
    core = dt.ms_deform_attn_core_pytorch
    cape_b200.patch_reference(dt, swap_forward_inference=True)
    try:
        assert rf.RoomFormerV2.forward_inference is not original
        assert rf.RoomFormerV2.forward_inference.__wrapped__ is original
        assert dt.ms_deform_attn_core_pytorch is cape_b200.ms_deform_attn_core_pytorch
        cape_b200.patch_reference(dt, swap_forward_inference=True)          # idempotent
        assert rf.RoomFormerV2.forward_inference.__wrapped__ is original
    finally:
        cape_b200.unpatch_reference()
    assert rf.RoomFormerV2.forward_inference is original and dt.ms_deform_attn_core_pytorch is core


# ---- the data-parallel wrapper around the REAL train_one_epoch_episodic (gloo, world size 2, CPU) ----------------------
# hidden_dim stays 256: the reference hard-codes 128 sine features per coordinate for the query positions
# (deformable_transformer_v2.py:1005-1018), other widths do not run
_SMALL_MODEL = ["--enc_layers", "1", "--dec_layers", "1", "--dim_feedforward", "128", "--dropout", "0.0", "--lr", "1e-3"]


def _small_reference_model():
    model, criterion, args, _ = stage_reference.build_cape_model("cpu", extra_args=_SMALL_MODEL, seed=3)
    model.train()
    criterion.train()
    for m in model.modules():                                    # no stochastic layers: both runs must be deterministic
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
        if isinstance(m, torch.nn.MultiheadAttention):
            m.dropout = 0.0
    return model, criterion, args


def _episode_batches(n_batches):
    import cape_b200
    return [cape_b200.synthetic.make_episode_batch(2, 2, num_keypoints=9, shots=1, image_size=64, seed=50 + i)
            for i in range(n_batches)]


def _dp_engine_worker(rank, world, port, out):
    import contextlib
    import io
    sys.path.insert(0, REPO)
    sys.path.insert(0, os.path.join(REPO, "tools"))
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port), TQDM_DISABLE="1")
    import torch.distributed as dist
    from cape_b200 import dist as cdist
    cdist.init_from_env("gloo")
    model, criterion, args = _small_reference_model()
    import util.misc
    from models.engine_cape import train_one_epoch_episodic
    opt = stage_reference.build_optimizer(model, args)
    buckets = cdist.GradBuckets(model.parameters(), bucket_bytes=1 << 20)
    with contextlib.redirect_stdout(io.StringIO()):
        cdist.train_one_epoch_data_parallel(train_one_epoch_episodic, model, criterion, _episode_batches(3), opt, "cpu", 0,
                                            buckets, accumulation_steps=2, max_norm=args.clip_max_norm,
                                            queries_per_episode=2, misc_module=util.misc)
    cdist.barrier()
    names = ["base_model.transformer.encoder.layers.0.linear1.weight", "base_model.query_embed.weight",
             "support_encoder.coord_mlp.0.weight" if hasattr(model.support_encoder, "coord_mlp") else "base_model.transformer.level_embed"]
    state = model.state_dict()
    out.put((rank, {k: state[k].double().sum().item() for k in names if k in state}, len(buckets.buckets), buckets.known))
    dist.destroy_process_group()


@pytest.mark.timeout(400)
def test_real_training_loop_data_parallel_equals_single_process():
    """train_one_epoch_data_parallel drives the reference's UNEDITED train_one_epoch_episodic (engine_cape.py:48-301) on two
    gloo ranks, each on its episodes rank::2 of every batch: both ranks must end with the weights a single process gets from
    the whole batches (3 batches, accumulation 2 -> one full and one ragged optimizer step, clipping on)."""
    import contextlib
    import io
    import torch.multiprocessing as mp
    from tests.test_dist_gloo import _free_port
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dp_engine_worker, args=(r, 2, port, q)) for r in range(2)]
    saved_path = list(sys.path)
    sys.path.insert(0, REPO)      # spawn hands sys.path to the children: `tests` must resolve to THIS repo's package even
    try:                          # if an earlier test left the reference checkout (which has its own `tests/`) in front
        for p in procs:
            p.start()
    finally:
        sys.path[:] = saved_path
    results = sorted((q.get(timeout=240) for _ in procs), key=lambda r: r[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    os.environ["TQDM_DISABLE"] = "1"
    model, criterion, args = _small_reference_model()
    stage_reference.activate()
    from models.engine_cape import train_one_epoch_episodic
    opt = stage_reference.build_optimizer(model, args)
    with contextlib.redirect_stdout(io.StringIO()):
        train_one_epoch_episodic(model, criterion, _episode_batches(3), opt, "cpu", 0, max_norm=args.clip_max_norm,
                                 accumulation_steps=2)
    state = model.state_dict()
    for rank, sums, n_buckets, known in results:
        assert known and n_buckets >= 1 and sums
        for k, v in sums.items():
            want = state[k].double().sum().item()
            assert abs(v - want) <= 1e-4 * max(1.0, abs(want)), (rank, k, v, want)
