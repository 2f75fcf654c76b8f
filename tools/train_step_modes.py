"""bench.train_step in both linear modes, fresh process, with allocator statistics.  Development tool."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
steps = int(os.environ.get("STEPS", "2"))
for mode in sys.argv[1:] or ["fp32", "tf32x3", "fp32", "tf32x3"]:
    s0 = torch.cuda.memory_stats(dev)
    r = bench.train_step(dev, 0, 1, linear_mode=mode, steps=steps)
    s1 = torch.cuda.memory_stats(dev)
    d = {k: s1[k] - s0[k] for k in ("num_device_alloc", "num_device_free", "num_alloc_retries")}
    print(mode, r["ms_per_optimizer_step"], "ms", r["episodes_per_s"], "episodes/s", d,
          "peak GB", round(s1["allocated_bytes.all.peak"] / 2 ** 30, 1), "reserved GB", round(s1["reserved_bytes.all.peak"] / 2 ** 30, 1), flush=True)
