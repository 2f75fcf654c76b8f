"""Encoder (6 layers, N = 128) time with nn.Linear vs the opt-in 3xTF32 linears.  Development tool."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cape_b200

dev = torch.device("cuda", 0)
n = int(os.environ.get("GEN_N", "128"))
torch.manual_seed(0)
enc = cape_b200.DeformableTransformerEncoder(
    cape_b200.DeformableTransformerEncoderLayer(256, 1024, 0.1, "relu", 4, 8, 4), 6).to(dev).eval()
pyr = cape_b200.synthetic.CAPE_PYRAMID
src = torch.randn(n, 5440, 256, device=dev)
pos = torch.randn(n, 5440, 256, device=dev)
shapes = torch.tensor(pyr, device=dev)
starts = cape_b200.level_start_index_from_shapes(shapes)
valid = torch.ones(n, 4, 2, device=dev)
with torch.no_grad():
    for mode in ("fp32", "tf32x3", "fp32", "tf32x3"):
        cape_b200.set_linear_mode(mode)
        enc(src, shapes, starts, valid, pos, None)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            enc(src, shapes, starts, valid, pos, None)
        torch.cuda.synchronize()
        print(f"{mode}: {(time.perf_counter() - t0) / 3 * 1e3:.1f} ms per encoder pass")
