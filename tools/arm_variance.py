"""Run-to-run spread of bench.py's model-level arms inside ONE process (development tool): each arm N times."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
n = int(os.environ.get("REPS", 3))
arms = {"train_amp": lambda: bench.cape_train_step(dev, 0, 1, amp=True)["ms_per_optimizer_step"],
        "train": lambda: bench.cape_train_step(dev, 0, 1)["ms_per_optimizer_step"],
        "infer": lambda: bench.cape_inference(dev, 0, 1)["s_per_batch"],
        "generation": lambda: bench.generation(dev)["episodes_per_s"]}
for name in (sys.argv[1:] or list(arms)):
    vals = []
    for _ in range(n):
        try:
            vals.append(arms[name]())
        except Exception as exc:  # noqa: BLE001
            vals.append(f"{type(exc).__name__}: {exc}"[:120])
    print(name, json.dumps(vals), flush=True)
