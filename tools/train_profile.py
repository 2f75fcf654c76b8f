"""Where the time of one CAPE training micro-batch goes on the GPU (torch.profiler kernel table, patched model).
Development tool: prints the top kernels by device time and the share of the MSDeformAttn sampling kernels.
    python tools/train_profile.py [tf32x3]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import cape_b200
import stage_reference as sr

tc = len(sys.argv) > 1 and sys.argv[1] == "tf32x3"
dev = torch.device("cuda:0")
sr.activate()
import models.deformable_transformer as dt
if tc:
    cape_b200.patch_reference(dt, swap_layer_classes=True)
model, criterion, margs, _ = sr.build_cape_model(dev, seed=1234)
cape_b200.unpatch_reference()
model.train()
criterion.train()
batch = cape_b200.synthetic.make_episode_batch(10, 2, 17, 5, seed=1)
cape_b200.patch_reference(dt)
cape_b200.set_linear_mode("tf32x3" if tc else "fp32")


def step():
    b = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in batch.items()}
    targets = {k: v.to(dev) if torch.is_tensor(v) else v for k, v in b["query_targets"].items()} \
        if isinstance(b["query_targets"], dict) else b["query_targets"]
    out = model(samples=b["query_images"], support_coords=b["support_coords"], support_mask=b["support_masks"],
                targets=targets, skeleton_edges=b["support_skeletons"])
    losses = criterion(out, targets)
    loss = sum(losses[k] * criterion.weight_dict[k] for k in losses if k in criterion.weight_dict)
    loss.backward()
    model.zero_grad(set_to_none=True)


for _ in range(2):
    step()
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total, e.count) for e in prof.key_averages()]
rows = [r for r in rows if r[1] > 0]
total = sum(r[1] for r in rows)
msda = sum(r[1] for r in rows if "msda_" in r[0])
print(f"linear mode {'tf32x3' if tc else 'fp32 (cuBLAS)'}: device time of one micro-batch (10 episodes x 2 queries) fwd + bwd = {total / 1e3:.1f} ms, "
      f"{sum(r[2] for r in rows)} kernel launches; MSDeformAttn sampling kernels {msda / 1e3:.1f} ms = {100 * msda / total:.1f} %")
for name, t, c in sorted(rows, key=lambda r: -r[1])[:22]:
    print(f"{t / 1e3:9.2f} ms {100 * t / total:5.1f} %  x{c:<5d} {name[:110]}")
