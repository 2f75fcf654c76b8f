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
