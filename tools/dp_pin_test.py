"""Does the CPU-core assignment of the ranks explain the fp32 training arm's slower steps at N > 1?  (development tool)
    torchrun --nproc-per-node 2 tools/dp_pin_test.py   with PIN=none|interleave|block"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from cape_b200 import dist as cdist

rank, world, local = cdist.init_from_env("nccl")
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
mode = os.environ.get("PIN", "none")
cores = sorted(os.sched_getaffinity(0))
if mode == "interleave":
    os.sched_setaffinity(0, set(cores[rank::world]))
elif mode == "block":
    k = len(cores) // world
    os.sched_setaffinity(0, set(cores[rank * k:(rank + 1) * k]))
probe = os.environ.get("PROBE", "0")
if probe == "1":
    bench.copy_probe(dev, rank, world)
elif probe == "2":                                   # only the allocations of the probe, no copies
    bufs = [torch.empty(256 << 20, dtype=torch.uint8).pin_memory() for _ in range(2)]
    devs = [torch.empty(256 << 20, dtype=torch.uint8, device=dev) for _ in range(2)]
    del bufs, devs
elif probe == "3":                                   # device buffers only
    devs = [torch.empty(256 << 20, dtype=torch.uint8, device=dev) for _ in range(2)]
    del devs
r = bench.cape_train_step(dev, rank, world)
if rank == 0:
    print(f"PROBE={probe} PIN={mode:10s} cores/rank {len(os.sched_getaffinity(0)):3d} of {len(cores)}  fp32 arm {r['ms_per_optimizer_step']:.1f} ms/step  "
          f"{r['episodes_per_s']:.1f} episodes/s", flush=True)
cdist.barrier(dev)
torch.distributed.destroy_process_group()
