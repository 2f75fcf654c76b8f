"""Episode-level data parallelism (one process per GPU, ``torch.distributed``; NCCL on B200, gloo in CPU tests).

The MSDeformAttn path has no cross-sample term: episodes are sharded across ranks and the op itself
needs no collective.  The only exchange in training is one gradient all-reduce per optimizer step, which the
reference never had (its ``util/misc.py:341-377`` bootstrap is never called).  ``FlatGradAllreduce`` does it over a
single flat bucket with fixed slots, so parameters that never receive a gradient (38 tensors in ``CAPEModel``,
SURVEY.md §5) contribute zeros instead of breaking the collective the way stock DDP does without
``find_unused_parameters``.
"""
from __future__ import annotations

import os
from typing import Iterable, List, Sequence

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None):
    """Initialise the default process group from RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* (torchrun's contract).
    Returns (rank, world_size, local_rank).  Single-process runs (no WORLD_SIZE) return (0, 1, 0) without a group."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", str(rank)))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, rank=rank, world_size=world,
                                    device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local_rank


def world_size() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def shard_episodes(n_episodes: int, rank: int, world: int) -> List[int]:
    """Episodes owned by ``rank``: ``rank, rank + world, ...`` (round robin, so ragged tails spread evenly)."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    return list(range(rank, n_episodes, world))


def max_over_ranks(value: float, device="cpu") -> float:
    """The slowest rank's time — how every multi-GPU number here is reported."""
    if world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device="cpu") -> float:
    if world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def barrier(device=None):
    if world_size() > 1:
        if device is not None and torch.device(device).type == "cuda":
            dist.barrier(device_ids=[torch.device(device).index])
        else:
            dist.barrier()


class GradBuckets:
    """Gradients of every trainable parameter live as VIEWS of one flat fp32 buffer, cut into buckets that are
    all-reduced (sum, then / world) asynchronously from inside the backward pass as soon as every gradient of a bucket has
    been accumulated — so the collective of the tail buckets overlaps the rest of backward, there is no pack / unpack
    launch per parameter and no host synchronisation.

    Slots are fixed, so every rank reduces the same layout whether or not a given parameter produced a gradient
    (``CAPEModel`` has 38 trainable tensors that never do): such parameters contribute zeros and end the
    step with ``grad = None`` exactly as in a single-process run (AdamW then skips them, weight decay included).

    Gradient accumulation: call :meth:`begin_micro_batch` with ``sync=False`` for all but the last micro-batch of an
    optimizer step (local accumulation only, the ``no_sync`` of stock DDP), ``sync=True`` for the last.
    ``zero_grad()`` (also installed on a wrapped optimizer by :meth:`wrap_optimizer`) zeroes the flat buffer with one
    memset and keeps the views in place.

    Which parameters a bucket has to wait for is learned from the first optimizer step (every bucket is launched at the
    end of backward until then), after which the flat layout is re-sorted into the observed gradient order so buckets
    complete front to back.  The set of used parameters must not change between steps (``static_graph`` semantics);
    a violation raises instead of reducing a half-written bucket.
    """

    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_bytes: int = 32 << 20, overlap: bool = True):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if any(p.dtype != torch.float32 for p in self.params):
            raise TypeError("GradBuckets expects fp32 parameters (AMP keeps fp32 master weights under autocast)")
        self.index = {id(p): i for i, p in enumerate(self.params)}
        self.bucket_bytes = int(bucket_bytes)
        self.overlap = overlap
        self.sync = True
        self.known = False                       # which parameters produce gradients: learned during the first step
        self.ever_fired = [False] * len(self.params)
        self.fire_order: List[int] = []
        self._relayout_pending = False
        self._handles = []
        self._layout(list(range(len(self.params)))[::-1])     # backward visits parameters roughly in reverse order
        self._reset_step_state()

    # -- layout ------------------------------------------------------------------------------------------------------
    def _layout(self, order: Sequence[int]) -> None:
        device = self.params[0].device if self.params else torch.device("cpu")
        total = sum(self.params[i].numel() for i in order)
        self.flat = torch.zeros(total, dtype=torch.float32, device=device)
        self.views: List[torch.Tensor] = [None] * len(self.params)
        self.bucket_of = [0] * len(self.params)
        self.buckets: List[List[int]] = [[]]
        self.bucket_range: List[List[int]] = [[0, 0]]
        off, limit = 0, max(1, self.bucket_bytes // 4)
        for i in order:
            n = self.params[i].numel()
            if self.buckets[-1] and off + n - self.bucket_range[-1][0] > limit:
                self.buckets.append([])
                self.bucket_range.append([off, off])
            self.views[i] = self.flat[off:off + n].view_as(self.params[i])
            self.bucket_of[i] = len(self.buckets) - 1
            self.buckets[-1].append(i)
            off += n
            self.bucket_range[-1][1] = off
        self.numel = total

    def _reset_step_state(self) -> None:
        self.fired_now = [False] * len(self.params)
        self.pending = [sum(1 for i in b if self.ever_fired[i]) for b in self.buckets]
        self.launched = [False] * len(self.buckets)
        self.works = []
        self._callback_queued = False

    # -- installation ------------------------------------------------------------------------------------------------
    def install(self) -> "GradBuckets":
        """Point every ``p.grad`` at its view and register the post-accumulate hooks."""
        for i, p in enumerate(self.params):
            if p.grad is not None:
                self.views[i].copy_(p.grad)
            p.grad = self.views[i]
        if not self._handles:
            self._handles = [p.register_post_accumulate_grad_hook(self._hook) for p in self.params]
        return self

    def remove(self) -> None:
        for h in self._handles:
            h.remove()
        self._handles = []

    def wrap_optimizer(self, optimizer):
        """Make ``optimizer.zero_grad()`` — which the reference's loop calls (``engine_cape.py:88,258``) and which would
        otherwise drop the views (``set_to_none=True``) — zero the flat buffer instead."""
        optimizer.zero_grad = lambda set_to_none=True: self.zero_grad()
        return optimizer

    def zero_grad(self) -> None:
        if self._relayout_pending:               # gradients are being discarded anyway: free to move the slots
            order = self.fire_order + [i for i in range(len(self.params)) if not self.ever_fired[i]]
            self._layout(order)
            self._relayout_pending = False
            self._reset_step_state()
        else:
            self.flat.zero_()
        for i, p in enumerate(self.params):
            if self.ever_fired[i] or not self.known:
                p.grad = self.views[i]

    def begin_micro_batch(self, sync: bool) -> None:
        """Call before every forward/backward: ``sync`` says whether this backward ends an optimizer step."""
        self.sync = bool(sync)
        self.fired_now = [False] * len(self.params)

    # -- backward-time machinery ---------------------------------------------------------------------------------------
    def _hook(self, p: torch.nn.Parameter) -> None:
        i = self.index[id(p)]
        v = self.views[i]
        g = p.grad
        if g is None:
            return
        if g.data_ptr() != v.data_ptr():         # someone dropped the view (an external zero_grad(set_to_none=True))
            v.copy_(g)
            p.grad = v
        if not self.ever_fired[i]:
            if self.known:
                raise RuntimeError("GradBuckets: a parameter that produced no gradient during the first optimizer step "
                                   "produced one now; the set of used parameters must be static")
            self.ever_fired[i] = True
            self.fire_order.append(i)
        if self.fired_now[i]:
            return
        self.fired_now[i] = True
        if not self.sync:
            return
        if not self._callback_queued:
            torch.autograd.Variable._execution_engine.queue_callback(self._finalize)
            self._callback_queued = True
        if self.known and self.overlap and world_size() > 1:
            b = self.bucket_of[i]
            self.pending[b] -= 1
            if self.pending[b] == 0 and not self.launched[b]:
                self._launch(b)

    def _launch(self, b: int) -> None:
        lo, hi = self.bucket_range[b]
        self.launched[b] = True
        if hi > lo:
            self.works.append(dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM, async_op=True))

    def _finalize(self) -> None:
        """End of the synchronising backward: launch what is left, join, average."""
        world = world_size()
        if world > 1:
            for b in range(len(self.buckets)):
                if not self.launched[b]:
                    self._launch(b)
            for w in self.works:
                w.wait()
            self.flat.div_(world)
        if not self.known:
            self.known = True
            self._relayout_pending = True
        for i, p in enumerate(self.params):
            if not self.ever_fired[i]:
                p.grad = None                    # what a single-process run leaves behind for a never-used parameter
        self._reset_step_state()

    def reduce_now(self) -> None:
        """Manual mode (no hooks, or after a backward that ran with ``sync=False``): reduce everything in one go."""
        for i, p in enumerate(self.params):
            if p.grad is not None and p.grad.data_ptr() != self.views[i].data_ptr():
                self.views[i].copy_(p.grad)
                p.grad = self.views[i]
                if not self.ever_fired[i]:
                    self.ever_fired[i] = True
                    self.fire_order.append(i)
        self._finalize()


class FlatGradAllreduce:
    """One all-reduce (sum, then / world) of every trainable parameter's gradient, call-style interface kept from round 1
    (``allreduce()`` after the last backward of an optimizer step).  Now a thin front of :class:`GradBuckets`: gradients
    are views of the flat buffer, so the call is the collective itself — no per-parameter pack / unpack launches and no
    host read of per-parameter flags."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.buckets = GradBuckets(params, overlap=False)
        self.params = self.buckets.params
        self.numel = self.buckets.numel
        self._seen = [False] * len(self.params)

    def __call__(self) -> None:
        b = self.buckets
        for i, p in enumerate(self.params):
            if p.grad is not None:
                b.ever_fired[i] = True
                if not self._seen[i]:
                    self._seen[i] = True
                    b.fire_order.append(i)
            elif b.ever_fired[i]:
                b.views[i].zero_()               # had gradients before, none this step: contributes zeros
        b.known = False                          # manual mode: usage is re-read from p.grad every call
        b.reduce_now()
        b._relayout_pending = False
        for i, p in enumerate(self.params):      # leave plain .grad tensors behind (callers may set_to_none afterwards)
            if b.ever_fired[i] and p.grad is None:
                p.grad = b.views[i]


class ShardedEpisodeLoader:
    """Per-rank view of an episodic loader for data-parallel training around the reference's unmodified
    ``train_one_epoch_episodic`` (``engine_cape.py:48-301``).

    * sharding: every rank iterates the same global loader (same seed) and keeps episodes ``rank::world`` of each collated
      batch (rows ``e*K .. (e+1)*K`` of every ``(B*K, ...)`` tensor and list, ``episodic_collate_fn``'s layout) — the
      global batch equals the single-process one.  With ``shard=False`` the loader is assumed to be per-rank already
      (built with :func:`rank_seed`);
    * it tells the :class:`GradBuckets` which micro-batches end an optimizer step — ``(batch_idx + 1) % accumulation_steps
      == 0`` or the last batch of the epoch, the two places the reference calls ``optimizer.step()`` (:230-258, :276-290)."""

    def __init__(self, loader, buckets: "GradBuckets | None", accumulation_steps: int = 1, rank: int = 0, world: int = 1,
                 queries_per_episode: int = 2, shard: bool = True):
        self.loader, self.buckets = loader, buckets
        self.accumulation_steps = max(1, int(accumulation_steps))
        self.rank, self.world, self.k, self.shard = rank, world, queries_per_episode, shard

    def __len__(self):
        return len(self.loader)

    def __iter__(self):
        it = iter(self.loader)
        try:
            nxt = next(it)
        except StopIteration:
            return
        idx = 0
        while True:
            cur = nxt
            try:
                nxt = next(it)
                last = False
            except StopIteration:
                last = True
            if self.buckets is not None:
                self.buckets.begin_micro_batch(last or (idx + 1) % self.accumulation_steps == 0)
            yield shard_batch(cur, self.rank, self.world, self.k) if self.shard and self.world > 1 else cur
            idx += 1
            if last:
                return


def shard_batch(batch: dict, rank: int, world: int, queries_per_episode: int) -> dict:
    """Rows of a collated episodic batch that belong to episodes ``rank::world`` (``episodic_collate_fn`` puts the K
    queries of episode e — and its repeated support — at rows ``e*K .. (e+1)*K``, episodic_sampler.py:432-470)."""
    k = queries_per_episode
    n = None
    for v in batch.values():
        if isinstance(v, torch.Tensor):
            n = v.shape[0]
            break
    if n is None or n % k:
        raise ValueError("batch has no (B*K, ...) tensor or B*K is not a multiple of queries_per_episode")
    rows = [e * k + j for e in shard_episodes(n // k, rank, world) for j in range(k)]
    index = torch.tensor(rows, dtype=torch.long)

    def take(v):
        if isinstance(v, torch.Tensor) and v.dim() > 0 and v.shape[0] == n:
            return v.index_select(0, index.to(v.device))
        if isinstance(v, (list, tuple)) and len(v) == n:
            return type(v)(v[r] for r in rows)
        if isinstance(v, dict):
            return {kk: take(vv) for kk, vv in v.items()}
        return v
    return {key: take(v) for key, v in batch.items()}


def rank_seed(seed: int, rank: int) -> int:
    """Seed of rank ``rank``'s own episodic sampler (``EpisodicSampler(seed=...)``, episodic_sampler.py:17-36) when every
    rank draws its own episodes instead of slicing a global batch."""
    return int(seed) + int(rank)


def train_one_epoch_data_parallel(train_one_epoch, model, criterion, loader, optimizer, device, epoch: int,
                                  buckets: GradBuckets, accumulation_steps: int = 1, max_norm: float = 0.0, scaler=None,
                                  queries_per_episode: int = 2, shard: bool = True, misc_module=None, **kwargs):
    """Run the reference's own ``train_one_epoch_episodic`` (passed in as ``train_one_epoch``; engine_cape.py:48) as one
    rank of a data-parallel job: the loader is sharded per rank, gradients are averaged over ranks from inside the last
    backward of every accumulation window (before the loop's ``clip_grad_norm_`` and ``optimizer.step()``), and
    ``optimizer.zero_grad()`` keeps the flat gradient views.  The loop itself is not edited."""
    rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    buckets.install()
    buckets.wrap_optimizer(optimizer)
    sharded = ShardedEpisodeLoader(loader, buckets, accumulation_steps, rank, world_size(), queries_per_episode, shard)
    # The reference's logging helper cannot run with more than one process: reduce_dict (util/misc.py:128-153) does
    # torch.stack() over a loss dict that holds Python floats next to tensors and raises TypeError as soon as a process
    # group exists (the reference never initialises one).  With `misc_module` (the imported util.misc) the
    # loop's logging is kept per-rank for the duration of the call; gradients are still averaged by `buckets`.
    saved = {}
    if misc_module is not None and world_size() > 1:
        for name, fn in (("get_world_size", lambda: 1), ("is_dist_avail_and_initialized", lambda: False)):
            saved[name] = getattr(misc_module, name)
            setattr(misc_module, name, fn)
    try:
        return train_one_epoch(model, criterion, sharded, optimizer, device, epoch, max_norm=max_norm,
                               accumulation_steps=accumulation_steps, scaler=scaler, **kwargs)
    finally:
        for name, fn in saved.items():
            setattr(misc_module, name, fn)


def shard_sizes(total: int, world: int) -> Sequence[int]:
    """Per-rank item counts of a round-robin shard of ``total`` items."""
    return [len(range(r, total, world)) for r in range(world)]


def bind_to_gpu_numa_node(device_index: int) -> int | None:
    """Pin this process (and therefore its first-touch pinned host buffers) to the CPU cores of the NUMA node the GPU hangs
    off, so that with one process per GPU the host<->device copies of different ranks do not all cross the socket
    interconnect.  Best effort: returns the node number, or None when the topology cannot be read."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(device_index)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:            # nvml reports an 8-digit PCI domain, sysfs uses 4
            bus = bus[4:]
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            # single-node topology (the B200 boxes of this pool report numa_node = -1 for every GPU): there is no socket
            # to stay on; give every rank its own slice of the allowed cores instead so their copy / launch threads do
            # not migrate onto each other
            world = world_size()
            cores = sorted(os.sched_getaffinity(0))
            if world > 1 and len(cores) >= world:
                rank = dist.get_rank() if dist.is_initialized() else device_index
                os.sched_setaffinity(0, set(cores[rank::world]))
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


# ---- checkpoints in the reference's layout ----------------------------------------------------
CHECKPOINT_KEYS = ("model", "optimizer", "lr_scheduler", "scaler", "epoch", "args", "train_stats", "val_stats", "best_pck",
                   "epochs_without_improvement", "rng_state", "np_rng_state", "py_rng_state")


def checkpoint_name(epoch: int, lr: float, batch_size: int, accumulation_steps: int, queries_per_episode: int) -> str:
    """File name of ``/root/reference/models/train_cape_episodic.py:853-859``."""
    return f"checkpoint_e{epoch:03d}_lr{lr:.0e}_bs{batch_size}_acc{accumulation_steps}_qpe{queries_per_episode}.pth"


def save_checkpoint(path, model, optimizer, lr_scheduler, epoch: int, args=None, scaler=None, train_stats=None,
                    val_stats=None, best_pck: float = 0.0, epochs_without_improvement: int = 0) -> bool:
    """Write the dict of ``train_cape_episodic.py:863-888`` (same keys, same meaning) so either code base resumes from the
    other's file.  Data-parallel runs hold identical replicas: rank 0 alone writes (atomically: temp file + rename), then
    every rank meets at a barrier.  Returns True on the rank that wrote."""
    import random

    import numpy as np
    wrote = False
    rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    if rank == 0:
        state = {
            "model": model.state_dict(), "optimizer": optimizer.state_dict(), "lr_scheduler": lr_scheduler.state_dict(),
            "scaler": scaler.state_dict() if scaler is not None else None, "epoch": epoch, "args": args,
            "train_stats": train_stats, "val_stats": val_stats, "best_pck": best_pck,
            "epochs_without_improvement": epochs_without_improvement, "rng_state": torch.get_rng_state(),
            "np_rng_state": np.random.get_state(), "py_rng_state": random.getstate(),
        }
        if torch.cuda.is_available():
            state["cuda_rng_state"] = torch.cuda.get_rng_state_all()
        tmp = f"{path}.tmp.{os.getpid()}"
        torch.save(state, tmp)
        os.replace(tmp, path)
        wrote = True
    barrier()
    return wrote


def load_checkpoint(path, model, optimizer=None, lr_scheduler=None, scaler=None, restore_rng: bool = True) -> dict:
    """Resume as ``train_cape_episodic.py:633-696`` does: non-strict model load (old checkpoints carry the KV-cache buffers
    and duplicated support-layer keys the reference leaks into ``state_dict()``), optimizer /
    scheduler / scaler state, best-model tracking and the four RNG streams.  Every rank reads the same file.  Returns
    ``{"start_epoch", "best_pck", "epochs_without_improvement", "missing_keys", "unexpected_keys"}``."""
    import random

    import numpy as np
    ckpt = torch.load(path, map_location="cpu", weights_only=False)
    missing, unexpected = model.load_state_dict(ckpt["model"], strict=False)
    if optimizer is not None and ckpt.get("optimizer") is not None:
        optimizer.load_state_dict(ckpt["optimizer"])
    if lr_scheduler is not None and ckpt.get("lr_scheduler") is not None:
        lr_scheduler.load_state_dict(ckpt["lr_scheduler"])
    if scaler is not None and ckpt.get("scaler") is not None:
        scaler.load_state_dict(ckpt["scaler"])
    if restore_rng:
        if "rng_state" in ckpt:
            torch.set_rng_state(ckpt["rng_state"].cpu())
        if "cuda_rng_state" in ckpt and torch.cuda.is_available() \
                and len(ckpt["cuda_rng_state"]) == torch.cuda.device_count():
            torch.cuda.set_rng_state_all(ckpt["cuda_rng_state"])
        if "np_rng_state" in ckpt:
            np.random.set_state(ckpt["np_rng_state"])
        if "py_rng_state" in ckpt:
            random.setstate(ckpt["py_rng_state"])
    return {"start_epoch": int(ckpt["epoch"]) + 1, "best_pck": ckpt.get("best_pck", 0.0),
            "epochs_without_improvement": ckpt.get("epochs_without_improvement", 0),
            "missing_keys": list(missing), "unexpected_keys": list(unexpected)}
