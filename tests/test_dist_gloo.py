"""world_size-2 gloo runs of the data-parallel host logic (episode sharding, max-over-ranks timing, flat-bucket
gradient all-reduce with never-used parameters)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from cape_b200 import dist as cdist
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    r, w, _ = cdist.init_from_env("gloo")
    assert (r, w) == (rank, world)
    episodes = cdist.shard_episodes(7, r, w)
    slow = cdist.max_over_ranks(10.0 + r)
    total = cdist.sum_over_ranks(len(episodes))
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2))
    unused = torch.nn.Parameter(torch.ones(5))               # trainable, never receives a gradient
    params = list(model.parameters()) + [unused]
    x = torch.full((2, 4), float(r + 1))
    model(x).sum().backward()
    local = [p.grad.clone() for p in model.parameters()]
    cdist.FlatGradAllreduce(params)()
    gathered = [None] * w
    dist.all_gather_object(gathered, [g.tolist() for g in local])
    ok = unused.grad is None
    for i, p in enumerate(model.parameters()):
        mean = sum(torch.tensor(gathered[k][i]) for k in range(w)) / w
        ok &= torch.allclose(p.grad, mean, atol=1e-6)
    cdist.barrier()
    out.put((rank, episodes, slow, total, bool(ok)))
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_gloo_sharding_and_grad_allreduce():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=150) for _ in procs)
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    (r0, e0, slow0, tot0, ok0), (r1, e1, slow1, tot1, ok1) = results
    assert e0 == [0, 2, 4, 6] and e1 == [1, 3, 5]
    assert slow0 == slow1 == 11.0 and tot0 == tot1 == 7.0
    assert ok0 and ok1


def test_single_process_helpers():
    from cape_b200 import dist as cdist
    assert cdist.shard_episodes(5, 0, 1) == [0, 1, 2, 3, 4]
    assert list(cdist.shard_sizes(10, 4)) == [3, 3, 2, 2]
    assert cdist.max_over_ranks(3.5) == 3.5
    with pytest.raises(ValueError):
        cdist.shard_episodes(5, 2, 2)


def _ckpt_worker(rank, world, port, path, out):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from cape_b200 import dist as cdist
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    cdist.init_from_env("gloo")
    torch.manual_seed(3)                                     # replicas start identical
    model = torch.nn.Linear(4, 2)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3)
    sched = torch.optim.lr_scheduler.StepLR(opt, 2)
    model(torch.ones(1, 4) * (rank + 1)).sum().backward()
    cdist.FlatGradAllreduce(model.parameters())()
    opt.step()
    sched.step()
    wrote = cdist.save_checkpoint(path, model, opt, sched, epoch=4, best_pck=0.25, epochs_without_improvement=1)
    exists_after_barrier = os.path.exists(path)             # the barrier inside save_checkpoint orders this read
    fresh = torch.nn.Linear(4, 2)
    info = cdist.load_checkpoint(path, fresh, restore_rng=False)
    same = all(torch.equal(a, b) for a, b in zip(model.state_dict().values(), fresh.state_dict().values()))
    out.put((rank, wrote, exists_after_barrier, same, info["start_epoch"]))
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_checkpoint_is_written_by_rank_zero_only_and_loads_on_every_rank(tmp_path):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    path = str(tmp_path / "ckpt.pth")
    procs = [ctx.Process(target=_ckpt_worker, args=(r, 2, port, path, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=150) for _ in procs)
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    assert results == [(0, True, True, True, 5), (1, False, True, True, 5)]
    assert [f for f in os.listdir(tmp_path) if ".tmp." in f] == []


def test_checkpoint_layout_is_the_reference_dict_and_resume_restores_everything(tmp_path):
    """Keys of train_cape_episodic.py:863-888; resume semantics of :633-696 (non-strict model load, RNG streams)."""
    import random

    import numpy as np
    from cape_b200 import dist as cdist
    assert cdist.checkpoint_name(10, 1e-4, 2, 4, 2) == "checkpoint_e010_lr1e-04_bs2_acc4_qpe2.pth"
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(3, 3), torch.nn.Linear(3, 1))
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=1e-4)
    sched = torch.optim.lr_scheduler.StepLR(opt, 3)
    for _ in range(2):
        opt.zero_grad()
        model(torch.randn(5, 3)).sum().backward()
        opt.step()
        sched.step()
    path = str(tmp_path / cdist.checkpoint_name(1, 1e-4, 2, 4, 2))
    assert cdist.save_checkpoint(path, model, opt, sched, epoch=1, args={"lr": 1e-4}, train_stats={"loss": 1.0},
                                 val_stats={"pck": 0.5}, best_pck=0.5, epochs_without_improvement=2)
    expect = (torch.rand(3), np.random.rand(3), random.random())            # what the RNG streams produce next
    raw = torch.load(path, weights_only=False)
    assert set(cdist.CHECKPOINT_KEYS) <= set(raw) and raw["scaler"] is None and raw["epoch"] == 1
    # a reference-made checkpoint carries leaked cache buffers (Appendix A.2): they must not break the load
    raw["model"]["0.kv_cache.k_cache"] = torch.zeros(1)
    torch.save(raw, path)
    model2 = torch.nn.Sequential(torch.nn.Linear(3, 3), torch.nn.Linear(3, 1))
    opt2 = torch.optim.AdamW(model2.parameters(), lr=1.0)
    sched2 = torch.optim.lr_scheduler.StepLR(opt2, 3)
    torch.rand(10), np.random.rand(10), random.random()                      # disturb the streams
    info = cdist.load_checkpoint(path, model2, opt2, sched2)
    assert info["start_epoch"] == 2 and info["best_pck"] == 0.5 and info["epochs_without_improvement"] == 2
    assert info["unexpected_keys"] == ["0.kv_cache.k_cache"] and info["missing_keys"] == []
    assert all(torch.equal(a, b) for a, b in zip(model.state_dict().values(), model2.state_dict().values()))
    assert opt2.param_groups[0]["lr"] == opt.param_groups[0]["lr"] and sched2.last_epoch == sched.last_epoch
    got = (torch.rand(3), np.random.rand(3), random.random())
    assert torch.equal(got[0], expect[0]) and np.array_equal(got[1], expect[1]) and got[2] == expect[2]


# ---- GradBuckets: gradients as views, all-reduce launched from inside the last backward of an accumulation window ------
def _toy_batches(n_batches, episodes=4, k=2):
    g = torch.Generator().manual_seed(7)
    return [{"x": torch.randn(episodes * k, 4, generator=g), "tag": [f"b{i}r{r}" for r in range(episodes * k)],
             "nested": {"y": torch.randn(episodes * k, 2, generator=g)}} for i in range(n_batches)]


def _toy_model():
    torch.manual_seed(11)
    model = torch.nn.Sequential(torch.nn.Linear(4, 8), torch.nn.Tanh(), torch.nn.Linear(8, 2))
    model.register_parameter("never_used", torch.nn.Parameter(torch.ones(5)))
    return model


def _reference_shaped_loop(model, loader, optimizer, accumulation_steps, max_norm):
    """The control flow of engine_cape.py:88-290: zero_grad, backward per micro-batch with the loss divided by the
    accumulation steps, clip + step + zero_grad every `accumulation_steps` batches and once more for a ragged tail."""
    optimizer.zero_grad()
    idx = -1
    for idx, batch in enumerate(loader):
        loss = ((model(batch["x"]) - batch["nested"]["y"]) ** 2).mean()
        (loss / accumulation_steps).backward()
        if (idx + 1) % accumulation_steps == 0:
            torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm)
            optimizer.step()
            optimizer.zero_grad()
    if (idx + 1) % accumulation_steps != 0:
        torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm)
        optimizer.step()
        optimizer.zero_grad()


def _bucket_worker(rank, world, port, out):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from cape_b200 import dist as cdist
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    cdist.init_from_env("gloo")
    model = _toy_model()
    opt = torch.optim.AdamW(model.parameters(), lr=0.05, weight_decay=0.1)
    buckets = cdist.GradBuckets(model.parameters(), bucket_bytes=64)          # several tiny buckets
    n_buckets_first = len(buckets.buckets)
    loader = _toy_batches(7)                                                   # 7 batches, accumulation 3: ragged tail

    def engine(model, criterion, loader, optimizer, device, epoch, max_norm=0, accumulation_steps=1, scaler=None):
        _reference_shaped_loop(model, loader, optimizer, accumulation_steps, max_norm)

    cdist.train_one_epoch_data_parallel(engine, model, None, loader, opt, "cpu", 0, buckets, accumulation_steps=3,
                                        max_norm=0.5, queries_per_episode=2)
    cdist.barrier()
    out.put((rank, {k: v.tolist() for k, v in model.state_dict().items()}, n_buckets_first, buckets.known,
             model.never_used.grad is None))
    dist.destroy_process_group()


@pytest.mark.timeout(240)
def test_grad_buckets_overlapped_allreduce_matches_single_process_training():
    """Two ranks, each on its episodes rank::2 of every global batch, must end with the weights of ONE process training
    on the whole batches — through gradient accumulation, clipping, a ragged last window and a never-used parameter."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_bucket_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted((q.get(timeout=200) for _ in procs), key=lambda r: r[0])
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    # single process on the full batches: mean over 8 rows == mean of the two ranks' 4-row means
    model = _toy_model()
    opt = torch.optim.AdamW(model.parameters(), lr=0.05, weight_decay=0.1)
    _reference_shaped_loop(model, _toy_batches(7), opt, 3, 0.5)
    want = model.state_dict()
    for rank, state, n_buckets, known, unused_none in results:
        assert n_buckets > 1 and known and unused_none
        for k in want:
            assert torch.allclose(torch.tensor(state[k]), want[k], atol=1e-6), (rank, k)
    assert torch.equal(want["never_used"], torch.ones(5))                      # untouched: no decay on a None gradient


def test_shard_batch_and_sync_schedule():
    from cape_b200 import dist as cdist
    batch = _toy_batches(1, episodes=3, k=2)[0]
    s1 = cdist.shard_batch(batch, 1, 2, 2)
    assert s1["tag"] == ["b0r2", "b0r3"] and torch.equal(s1["x"], batch["x"][2:4])
    assert torch.equal(s1["nested"]["y"], batch["nested"]["y"][2:4])
    s0 = cdist.shard_batch(batch, 0, 2, 2)
    assert s0["tag"] == ["b0r0", "b0r1", "b0r4", "b0r5"]

    class Spy:
        def __init__(self):
            self.flags = []

        def begin_micro_batch(self, sync):
            self.flags.append(sync)
    spy = Spy()
    assert len(list(cdist.ShardedEpisodeLoader(_toy_batches(5), spy, accumulation_steps=2))) == 5
    assert spy.flags == [False, True, False, True, True]                       # step after 2, 4 and the ragged 5th
    assert cdist.rank_seed(42, 3) == 45
