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


class FlatGradAllreduce:
    """One all-reduce (sum, then / world) of every trainable parameter's gradient through a flat fp32 bucket.

    Slots are fixed at construction from ``params`` order, so every rank reduces the same layout whether or not a
    given parameter produced a gradient this step (``grad is None`` -> zeros in, and the averaged slot is written
    back only if some rank had a gradient, mirroring what a single-process run would leave as ``None``).
    """

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.offsets: List[int] = []
        total = 0
        for p in self.params:
            self.offsets.append(total)
            total += p.numel()
        self.numel = total
        device = self.params[0].device if self.params else torch.device("cpu")
        self.flat = torch.zeros(total + len(self.params), dtype=torch.float32, device=device)

    def __call__(self) -> None:
        world = world_size()
        n = self.numel
        flags = self.flat[n:]
        self.flat.zero_()
        for i, (p, off) in enumerate(zip(self.params, self.offsets)):
            if p.grad is not None:
                self.flat[off:off + p.numel()].copy_(p.grad.reshape(-1))
                flags[i] = 1.0
        if world > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat[:n].div_(world)
        has_grad = flags.tolist()
        for i, (p, off) in enumerate(zip(self.params, self.offsets)):
            if has_grad[i] > 0:
                g = self.flat[off:off + p.numel()].view_as(p).to(p.dtype)
                if p.grad is None:
                    p.grad = g.clone()
                else:
                    p.grad.copy_(g)


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
